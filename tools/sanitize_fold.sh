#!/bin/bash
# compute-sanitizer passes over the fold kernels (two tile pipelines + bulk-store epilogue), the sparse-fold kernels and a
# band-kernel fixture.  usage: gpurun -- 'bash tools/sanitize_fold.sh gpurun_out/<tag>'
OUT=${1:-gpurun_out/sanitize_fold}
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool (fold / sparse fold)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_fold_gpu.py -q -x -k "sparse or matches_fp64" > $OUT/fold_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/fold_$tool.log | tail -3
done
echo "== compute-sanitizer --tool memcheck (band kernel fixture)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -q -x -k "fixture_kernel_dispatch or hands_steep" > $OUT/band_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" $OUT/band_memcheck.log | tail -3
