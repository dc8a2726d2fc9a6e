#!/usr/bin/env python
"""Per-source-line / per-range summary of an `ncu --set full --import-source on` capture.

    python tools/ncu_lines.py gpurun_out/X.ncu-rep [--so multigrid_b200/_lib/libmultigrid_b200.so]
                              [--kernel step_obs_kernel] [--ranges name:lo-hi,...] [--top 40]

ncu's CSV source page is SASS-only; this joins it with `nvdisasm -g` line info of the cubin in
the built .so (same build!) so stall samples and executed instructions can be read per CUDA
source line, and per named line range (the kernel's phases).
"""
from __future__ import annotations

import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import tempfile


def sass_rows(rep, kernel_id=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    kernels, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(dict(zip(cur["hdr"], r)))
    return kernels


def line_map(so, mangled_pat):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
        cubins = glob.glob(os.path.join(td, "*.cubin"))
        text = "".join(subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
                       for c in cubins)
    maps, cur, line, fname = {}, None, None, None
    for ln in text.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = maps.setdefault(m.group(1), {})
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            fname, line = os.path.basename(m.group(1)), int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m and cur is not None:
            cur[int(m.group(1), 16)] = (fname, line)
    return {k: v for k, v in maps.items() if re.search(mangled_pat, k)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--so", default="multigrid_b200/_lib/libmultigrid_b200.so")
    ap.add_argument("--kernel", default="step_obs_kernel")
    ap.add_argument("--mangled", default=None, help="regex on the mangled name (default: derived)")
    ap.add_argument("--ranges", default="")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--index", type=int, default=0, help="which matching launch in the report")
    args = ap.parse_args()

    ks = [k for k in sass_rows(args.rep) if args.kernel in k["name"]]
    k = ks[args.index]
    print("kernel:", k["name"], " SASS rows:", len(k["rows"]))
    m = re.findall(r"\((int|bool)\)(-?\d+)", k["name"])
    pat = args.mangled or (f"{args.kernel}I" + "".join(f"L{'i' if t == 'int' else 'b'}{v.replace('-', 'n')}E" for t, v in m) + "E" if m else args.kernel)
    lm = line_map(args.so, pat)
    assert len(lm) == 1, list(lm)
    lm = next(iter(lm.values()))
    base = int(k["rows"][0]["Address"], 16)
    per = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    stall_cols = [c for c in k["hdr"] if c.startswith("stall_") and "Not Issued" not in c]
    for r in k["rows"]:
        off = int(r["Address"], 16) - base
        key = lm.get(off, ("?", 0))
        c = per[key]
        vals = {"samples": int(r["# Samples"] or 0), "inst": int(r["Instructions Executed"] or 0),
                "thread_inst": int(r["Thread Instructions Executed"] or 0),
                "smem_wf": int(r["L1 Wavefronts Shared"] or 0),
                "smem_wf_ideal": int(r["L1 Wavefronts Shared Ideal"] or 0),
                "gsect": int(r["L2 Theoretical Sectors Global"] or 0),
                "gsect_ideal": int(r["L2 Theoretical Sectors Global Ideal"] or 0)}
        for sc in stall_cols:
            vals[sc] = int(r[sc] or 0)
        c.update(vals)
        tot.update(vals)
    print("total:", {k_: v for k_, v in tot.items() if not k_.startswith("stall_")})
    print("stalls:", {k_[6:]: v for k_, v in sorted(tot.items(), key=lambda kv: -kv[1]) if k_.startswith("stall_") and v})

    def show(label, c):
        st = sorted(((v, s[6:]) for s, v in c.items() if s.startswith("stall_") and v), reverse=True)[:3]
        print(f"{label:28s} samples {c['samples']:7d} ({100*c['samples']/max(1,tot['samples']):5.1f}%)  inst {c['inst']:9d} "
              f"({100*c['inst']/max(1,tot['inst']):5.1f}%, {c['thread_inst']/max(1,c['inst']):4.1f} thr)  smem_wf {c['smem_wf']:8d}/{c['smem_wf_ideal']:8d}  "
              f"gsect {c['gsect']:8d}/{c['gsect_ideal']:8d}  " + " ".join(f"{s}:{v}" for v, s in st))

    if args.ranges:
        print("\n-- ranges --")
        for spec in args.ranges.split(","):
            name, rng = spec.split(":")
            lo, hi = map(int, rng.split("-"))
            c = collections.Counter()
            for (f, l), v in per.items():
                if lo <= l <= hi:
                    c.update(v)
            show(name, c)
    print("\n-- top lines by samples --")
    for (f, l), c in sorted(per.items(), key=lambda kv: -kv[1]["samples"])[:args.top]:
        show(f"{f}:{l}", c)


if __name__ == "__main__":
    main()
