#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== debug batch"; timeout 300 python tools/debug_batch.py 2>&1 | grep -v "\[\]$" | tail -8
echo "=== pytest gpu (all)"; timeout 2400 python -m pytest tests/ -q -m gpu --timeout 1200 -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_report.json gpurun_out/parity_all.json 2>/dev/null
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
