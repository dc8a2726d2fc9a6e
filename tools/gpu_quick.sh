#!/bin/bash
# Quick GPU pass while developing a kernel: smoke (bounded), parity tests, launch-knob sweep.
# Usage (under gpurun): bash tools/gpu_quick.sh <tag> [kbench args...]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -3 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 600 python tools/kbench.py "$@" > $OUT/${TAG}_kbench.jsonl 2> $OUT/${TAG}_kbench.err; echo "kbench rc=$?"
cat $OUT/${TAG}_kbench.jsonl; tail -3 $OUT/${TAG}_kbench.err
