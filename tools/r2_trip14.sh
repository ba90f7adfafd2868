#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== debug batch"; timeout 200 python tools/debug_batch.py 2>&1 | tail -24
echo "=== timeline B=8"; timeout 200 python tools/batch_timeline.py 8 2>&1 | tail -64 | head -40
