#!/bin/bash
# final validation of the committed build (selective re-poll + L2 prefetch): full GPU suite, smoke, default bench, ncu of the decode step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -6 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_report.json gpurun_out/parity_all.json 2>/dev/null
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
t0=$(date +%s); echo "=== bench (default flags)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $? wall $(( $(date +%s) - t0 )) s"; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
echo "=== ncu full: fused decode step (batch 1)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 8 -c 1 -o gpurun_out/r02_prof_decode_b1_final -f \
    python tools/run_batch.py 1 16 1 > gpurun_out/ncu_b1.log 2>&1; echo "exit $?"
python tools/ncu_summary.py gpurun_out/r02_prof_decode_b1_final.ncu-rep gpurun_out/r02_decode_b1_final_ncu.txt | cut -c1-500
