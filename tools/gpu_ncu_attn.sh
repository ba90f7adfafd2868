#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_f32_kernel -s 40 -c 2 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/prof_attn.ncu-rep
