#!/bin/bash
# full validation + evidence for the shipped configuration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_gpu.log
echo "=== planes=2 experiment (full-size parity numbers only)"
ASRB_PLANES=2 timeout 600 python -m pytest tests/ -q -m gpu --timeout 600 -k "full_size" > gpurun_out/pytest_planes2.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_planes2.log; cp gpurun_out/parity_report.json gpurun_out/parity_planes2.json; grep -E "full_" gpurun_out/parity_planes2.json
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
echo "=== ncu full: decode step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 5 -c 1 -o gpurun_out/prof_mega -f \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
