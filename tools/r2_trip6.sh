#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== timeline B=8 flags=1"; ASRB_BATCH_FLAGS=1 timeout 300 python tools/batch_timeline.py 8 2>&1 | tail -26
echo "=== timeline B=8 flags=0"; ASRB_BATCH_FLAGS=0 timeout 300 python tools/batch_timeline.py 8 2>&1 | tail -26
