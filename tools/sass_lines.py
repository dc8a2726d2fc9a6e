#!/usr/bin/env python
"""Static SASS instruction count per source-line range of one kernel (nvdisasm -g line info).
    python tools/sass_lines.py <so> <mangled-substring> [name:lo-hi,...]"""
import collections, glob, os, re, subprocess, sys, tempfile

so, pat = sys.argv[1], sys.argv[2]
ranges = sys.argv[3] if len(sys.argv) > 3 else ""
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    text = "".join(subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
                   for c in glob.glob(os.path.join(td, "*.cubin")))
cur, line, per, ops = None, 0, collections.Counter(), collections.defaultdict(collections.Counter)
for ln in text.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        cur = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = int(m.group(2)); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur and pat in cur:
        per[line] += 1; ops[line][m.group(1)] += 1
print("total", sum(per.values()))
if ranges:
    for spec in ranges.split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        c = collections.Counter()
        for l in per:
            if lo <= l <= hi: c.update(ops[l])
        print(f"{name:12s} {sum(c.values()):6d}  {c.most_common(8)}")
else:
    for l, c in sorted(per.items()): print(l, c, ops[l].most_common(4))
