#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu (all, planes 3)"; timeout 2400 python -m pytest tests/ -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -12 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_report.json gpurun_out/parity_all.json 2>/dev/null
echo "=== pytest gpu (planes 2)"; ASRB_PLANES=2 timeout 2400 python -m pytest tests/ -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu_planes2.log 2>&1; echo "exit $?"; tail -12 gpurun_out/pytest_gpu_planes2.log
cp gpurun_out/parity_report.json gpurun_out/parity_planes2.json 2>/dev/null
echo "=== encoder roofline b8"; timeout 600 python tools/encoder_roofline.py 8 2>&1 | tail -25
echo "=== encoder roofline b64"; timeout 600 python tools/encoder_roofline.py 64 2>&1 | tail -25
