#!/bin/bash
# One GPU-box visit: smoke (+memcheck), parity tests, bench, ncu launch list.  Logs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
if [ "${SANITIZE:-1}" = "1" ]; then
  echo "=== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|smoke ok" gpurun_out/memcheck.log | head -20
fi
echo "=== pytest gpu"; timeout 2400 python -m pytest tests/ -q -m gpu --timeout 900 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
cat gpurun_out/parity_report.json 2>/dev/null | head -80
echo "=== bench"; timeout 1200 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "=== ncu launch list"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-1500} --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; wc -l gpurun_out/launches.csv
fi
