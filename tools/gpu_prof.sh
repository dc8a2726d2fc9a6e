#!/bin/bash
# ncu launch list + one full capture of the fused kernel, under the bench command.
# The captured launch is the ~1000th: by then agents have spread over the grids (steady state).
# Usage (under gpurun): bash tools/gpu_prof.sh <tag> [extra env assignments...]
TAG=${1:-p}; shift
for kv in "$@"; do export "$kv"; done
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 100 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1024 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_obs_kernel -s 1000 -c 1 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 1024 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
