#!/bin/bash
# One gpurun batch: tcgen05 probe, GPU parity tests, smoke, bench, ncu launch list + one full capture of the render kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host_cores.txt
if [ -x tools/bin/umma_probe ]; then echo "== umma probe"; timeout 60 tools/bin/umma_probe 2>&1 | tee $OUT/umma_probe.log; fi
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -q -m gpu -s 2>&1 | tail -60 | tee $OUT/pytest_gpu.log
PYRC=${PIPESTATUS[0]}
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee $OUT/bench.log
if [ -n "$RUN_REF" ]; then echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_ref.log; fi
if [ -z "$SKIP_NCU" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full capture (render kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -f -o $OUT/render_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > $OUT/ncu_full_bench.log 2>&1
fi
ls -la $OUT
