#!/usr/bin/env python
"""Per-source-line dynamic warp-instructions per warp (and stall samples) from an ncu capture.
    python tools/ncu_perwarp.py <rep> <mangled-pattern> <warps> [min_per_warp]"""
import sys, collections
sys.path.insert(0, 'tools')
import ncu_lines as N
rep, pat, warps = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 6
k = [k for k in N.sass_rows(rep) if 'step_obs' in k['name']][0]
lm = next(iter(N.line_map('multigrid_b200/_lib/libmultigrid_b200.so', pat).values()))
base = int(k['rows'][0]['Address'], 16)
per = collections.defaultdict(lambda: [0, 0, 0])
for r in k['rows']:
    l = lm.get(int(r['Address'], 16) - base, ('?', 0))
    per[l][0] += int(r['Instructions Executed']); per[l][1] += int(r['Thread Instructions Executed']); per[l][2] += int(r['# Samples'])
src = open('multigrid_b200/csrc/mg_kernels.cuh').read().splitlines()
for (f, l), (a, b, s) in sorted(per.items()):
    if a / warps >= thr or s >= 15:
        text = src[l - 1].strip()[:84] if f == 'mg_kernels.cuh' and l <= len(src) else ''
        print(f"{f[:14]:14s}:{l:4d} {a/warps:7.1f}/warp thr {b/max(a,1):5.1f} samp {s:4d} | {text}")
