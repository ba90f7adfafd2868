#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest ids"; timeout 600 python -m pytest tests/ -q -m gpu --timeout 600 -k "ids or eos or full_size" > gpurun_out/pytest_quick.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_quick.log
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 5 -c 1 -o gpurun_out/prof_mega -f \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
