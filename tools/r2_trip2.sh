#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== debug batch"; timeout 600 python tools/debug_batch.py 2>&1 | tail -20
echo "=== batch roofline 0.6B"; timeout 900 python tools/batch_decode_roofline.py 8 16 > gpurun_out/batch_decode.json 2> gpurun_out/batch_decode.err; echo "exit $?"; cat gpurun_out/batch_decode.json | cut -c1-1800; tail -5 gpurun_out/batch_decode.err
