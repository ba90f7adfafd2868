#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== timeline + event time, qkv"; ASRB_GEMM_DEBUG=3120,2688,896 PLANES=3 timeout 600 python tools/encoder_roofline.py 8 2>&1 | grep "gemm_tc t\|cta0" | tail -8
echo "=== per-GEMM event times, B=8"; ASRB_GEMM_TIME=1 PLANES=3 timeout 600 python tools/encoder_roofline.py 8 2> gpurun_out/gemm_times_b8.txt | grep "encoder_ms\|prefill_ms"; python tools/gemm_times.py < gpurun_out/gemm_times_b8.txt
echo "=== per-GEMM event times, B=1"; ASRB_GEMM_TIME=1 PLANES=3 timeout 600 python tools/encoder_roofline.py 1 2> gpurun_out/gemm_times_b1.txt | grep "encoder_ms\|prefill_ms"; python tools/gemm_times.py < gpurun_out/gemm_times_b1.txt
