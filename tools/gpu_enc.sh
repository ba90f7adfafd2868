#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest encoder/prefill (swiglu epilogue change)"; timeout 900 python -m pytest tests/ -q -m gpu --timeout 600 -k "encoder or prefill or full_size or ids" > gpurun_out/pytest_quick.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_quick.log
for b in 8 64; do echo "=== encoder roofline B=$b"; timeout 600 python tools/encoder_roofline.py $b > gpurun_out/encoder_roofline_b$b.json 2> gpurun_out/enc.err; echo "exit $?"; cat gpurun_out/encoder_roofline_b$b.json; tail -2 gpurun_out/enc.err; done
