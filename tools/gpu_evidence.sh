#!/bin/bash
# Round-2 evidence pass on one B200 (under gpurun): GPU tests, smoke, bench lines of the three single-GPU BASELINE
# configurations + the reference arm, ncu launch list of the default bench, one `ncu --set full` capture each of the
# static-grid kernel (Empty-8x8, the headline) and of the general kernel (BlockedUnlockPickup).
# Usage: bash tools/gpu_evidence.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --config bup --no-cpu-baseline > $OUT/${TAG}_bench_bup.json 2> $OUT/${TAG}_bench_bup.err; echo "bench bup rc=$?"
timeout 300 python bench.py --config empty16 --no-cpu-baseline > $OUT/${TAG}_bench_e16.json 2> $OUT/${TAG}_bench_e16.err; echo "bench e16 rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"
for f in bench bench_bup bench_e16 bench_ref; do python - "$OUT/${TAG}_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "value %.3g" % d["value"], "us/launch", round(d["ms_per_step"] * 1e3, 2), "frac", r.get("frac"), "e2e %.3g" % d["e2e"]["value"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
# launch list (short bench under ncu: per-launch times are cold and serialised; the kernel's SHARE is what counts)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-verify > $OUT/${TAG}_ncu_bench.log 2>&1
# full captures in steady state (agents spread over the grids)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:static_ -s 600 -c 1 -f -o $OUT/${TAG}_static_prof \
    python bench.py --steps 640 --warmup 4 --no-cpu-baseline --no-verify > $OUT/${TAG}_ncu_static.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_obs_kernel -s 600 -c 1 -f -o $OUT/${TAG}_bup_prof \
    python bench.py --config bup --steps 640 --warmup 4 --no-cpu-baseline --no-verify > $OUT/${TAG}_ncu_bup.log 2>&1
# step + one-hot (fused vs separate pass), general kernels, reset / fresh-layout timings, host cost of a step
timeout 300 python tools/kbench.py --configs empty8,bup,empty16 --extra ONEHOT=1,ONEHOT=2 > $OUT/${TAG}_kbench.jsonl 2> $OUT/${TAG}_kbench.err
MG_NO_STATIC=1 timeout 300 python tools/kbench.py --configs empty8,bup,empty16 --extra ONEHOT=1,ONEHOT=2,CHAINED=1 > $OUT/${TAG}_kbench_general.jsonl 2>> $OUT/${TAG}_kbench.err
timeout 300 python tools/reset_bench.py > $OUT/${TAG}_reset.jsonl 2> $OUT/${TAG}_reset.err
timeout 200 python tools/api_overhead.py > $OUT/${TAG}_api.jsonl 2>&1
# memcheck over the kernels added in round 2 (palette / wire packers, fused one-hot, layout refresh)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_packed_obs.py tests/test_fresh_layouts.py tests/test_fused_one_hot.py -m gpu -x -q -k "palette or wire or fresh or ragged or static_grid_random or wrapper" > $OUT/${TAG}_memcheck_new.log 2>&1; tail -3 $OUT/${TAG}_memcheck_new.log
ls -la $OUT/${TAG}_*
