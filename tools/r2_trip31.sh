#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu quick"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 -k "not full_size" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "exit $?"; tail -4 gpurun_out/pytest_gpu_quick.log
echo "=== mega timeline"; timeout 300 python tools/mega_timeline.py 2>&1 | head -22 | cut -c1-420
echo "=== bench b1"; timeout 900 python bench.py --workload b1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1300
