#!/bin/bash
# compute-sanitizer passes over the fused / chained / rollout / static kernels:
# racecheck (shared-memory hazards: the stage aliased onto the gathered cells, the transposed scratch, shadow
# lanes), synccheck (barrier / __syncwarp misuse) and memcheck on a selection of the GPU parity tests.
# Usage (under gpurun): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out; mkdir -p $OUT
SEL_STATIC="tests/test_static_path.py -k gpu_static_random or gpu_static_falls"
SEL_GENERAL='tests/test_gpu_parity.py -k "random_soup_vs_c_oracle and (0- or 1- or 5- or 7-) or back_to_back or chained_launches_rotating or ragged_batch or single_layout_dedup or (rollout_equals_single_steps and 0-)"'
for tool in racecheck synccheck memcheck; do
  for sel in static general; do
    if [ $sel = static ]; then ARGS="tests/test_static_path.py -k gpu_static_random"; else ARGS="tests/test_gpu_parity.py -k random_soup_vs_c_oracle or back_to_back or chained_launches_rotating or ragged_batch or single_layout_dedup"; fi
    LOG=$OUT/${TAG}_${tool}_${sel}.log
    if [ $sel = static ]; then
      timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_static_path.py -m gpu -x -q -k "gpu_static_random" > $LOG 2>&1
    else
      timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_soup_vs_c_oracle or back_to_back or chained_launches_rotating or ragged_batch or single_layout_dedup" > $LOG 2>&1
    fi
    echo "== $tool $sel rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" $LOG | tail -4
  done
done
