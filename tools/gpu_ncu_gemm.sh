#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# encoder at batch 8 (north_star "batch 8 x 30 s"): conv2/conv3 implicit GEMMs + first layers' linears
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 12 -o gpurun_out/prof_gemm_b8 -f \
    python tools/encoder_roofline.py 8 > gpurun_out/ncu_gemm_b8.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/prof_gemm_b8.ncu-rep
