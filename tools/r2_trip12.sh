#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== timeline B=8"; timeout 300 python tools/batch_timeline.py 8 2>&1 | tail -70
