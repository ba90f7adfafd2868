#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest (tc + mega, all)"; timeout 1200 python -m pytest tests/ -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -8 gpurun_out/pytest_gpu.log
cat gpurun_out/parity_report.json | grep -E "full_|prefill_logits"
echo "=== mega timeline"; timeout 600 python tools/mega_timeline.py > gpurun_out/mega_timeline.txt 2>&1; echo "exit $?"; cat gpurun_out/mega_timeline.txt | tail -40
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "=== ncu full: decode_step + gemm_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 3 -c 2 -o gpurun_out/prof_mega -f \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 6 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
ls -la gpurun_out/*.ncu-rep
