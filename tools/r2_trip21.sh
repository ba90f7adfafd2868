#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu quick"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 -k "not full_size" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_gpu_quick.log
echo "=== timeline + event time, fc1"; ASRB_GEMM_DEBUG=3120,3584,896 PLANES=3 timeout 600 python tools/encoder_roofline.py 8 2>&1 | grep "gemm_tc t\|cta0" | tail -7
echo "=== per-GEMM event times, B=8"; ASRB_GEMM_TIME=1 PLANES=3 timeout 600 python tools/encoder_roofline.py 8 2> gpurun_out/gemm_times_b8.txt | grep "encoder_ms\|prefill_ms"; python tools/gemm_times.py < gpurun_out/gemm_times_b8.txt
echo "=== per-GEMM event times, B=1"; ASRB_GEMM_TIME=1 PLANES=3 timeout 600 python tools/encoder_roofline.py 1 2> gpurun_out/gemm_times_b1.txt | grep "encoder_ms\|prefill_ms"; python tools/gemm_times.py < gpurun_out/gemm_times_b1.txt
echo "=== encoder roofline B=8 / B=1 (no instrumentation)"; PLANES=3 timeout 600 python tools/encoder_roofline.py 8 2>&1 | grep "encoder_ms\|prefill_ms"; PLANES=3 timeout 600 python tools/encoder_roofline.py 1 2>&1 | grep "encoder_ms\|prefill_ms"
