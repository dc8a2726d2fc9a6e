#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list + one full capture of
# the fused kernel in steady state (the ~1000th launch), per-warp timeline.
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
bash tools/gpu_prof.sh $TAG
timeout 300 python tools/trace_timeline.py > $OUT/${TAG}_timeline.txt 2>&1
ls -la $OUT | tail -20
