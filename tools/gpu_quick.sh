#!/bin/bash
# quick visit: mega tests + timeline + bench (no ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest ids"; timeout 600 python -m pytest tests/ -q -m gpu --timeout 600 -k "${PYTEST_K:-ids or eos}" > gpurun_out/pytest_quick.log 2>&1; echo "exit $?"; tail -4 gpurun_out/pytest_quick.log
echo "=== mega timeline"; timeout 300 python tools/mega_timeline.py > gpurun_out/mega_timeline.txt 2>&1; echo "exit $?"; tail -36 gpurun_out/mega_timeline.txt
if [ "${BENCH:-1}" = "1" ]; then
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
fi
