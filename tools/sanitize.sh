#!/bin/bash
# compute-sanitizer passes over the loss kernels (and the small render fixtures): racecheck, initcheck, synccheck, memcheck.
# usage: gpurun -- 'bash tools/sanitize.sh gpurun_out/<tag>'
OUT=${1:-gpurun_out/sanitize}
mkdir -p $OUT
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer --tool $tool (loss kernels)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_loss_gpu.py -q -x -k "smoothness or oracle_autograd" > $OUT/loss_$tool.log 2>&1
  tail -4 $OUT/loss_$tool.log
done
