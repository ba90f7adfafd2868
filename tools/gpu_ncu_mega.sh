#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 5 -c 1 -o gpurun_out/prof_mega -f \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
ls -la gpurun_out/*.ncu-rep
