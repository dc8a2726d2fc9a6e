#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list + one full capture of the fused kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 64 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_obs_kernel -s 80 -c 1 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 16 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
