#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== debug batch"; timeout 240 python tools/debug_batch.py 2>&1 | grep -v "\[\]$" | tail -8
echo "=== timeline B=8"; timeout 300 python tools/batch_timeline.py 8 2>&1 | tail -26
echo "=== timeline B=16"; timeout 300 python tools/batch_timeline.py 16 2>&1 | grep -E "us/step|total cycles|layer mean"
