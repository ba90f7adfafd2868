#!/bin/bash
# round 2, trip 1: first run of the batch-aware fused decode step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "=== pytest batch tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "batch or tiny" --timeout 300 > gpurun_out/pytest_batch.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_batch.log
echo "=== batch roofline 0.6B"; timeout 600 python tools/batch_decode_roofline.py 8 16 > gpurun_out/batch_decode.json 2> gpurun_out/batch_decode.err; echo "exit $?"; cat gpurun_out/batch_decode.json | cut -c1-1500; tail -5 gpurun_out/batch_decode.err
echo "=== pytest gpu (all)"; timeout 1200 python -m pytest tests/ -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -8 gpurun_out/pytest_gpu.log
