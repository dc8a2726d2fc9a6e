#!/bin/bash
# ncu launch list + one full capture of the fused kernel, under the bench command.
# Usage (under gpurun): bash tools/gpu_prof.sh <tag> [extra env assignments...]
TAG=${1:-p}; shift
for kv in "$@"; do export "$kv"; done
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 100 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 64 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_obs_kernel -s 80 -c 1 \
    -f -o $OUT/${TAG}_prof python bench.py --steps 16 --warmup 4 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
timeout 300 python bench.py --no-cpu-baseline --steps 256 > $OUT/${TAG}_bench.json 2>$OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
