#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu quick"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 -k "not full_size" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "exit $?"; tail -4 gpurun_out/pytest_gpu_quick.log
echo "=== mega timeline"; timeout 300 python tools/mega_timeline.py 2>&1 | head -22 | cut -c1-420
echo "=== bench b1"; timeout 900 python bench.py --workload b1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1300
echo "=== ncu launch list (bench b1, 8 new tokens)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_b1_final.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline --workload b1 > gpurun_out/ncu_bench.log 2>&1
echo "exit $?"; wc -l gpurun_out/r02_launches_b1_final.csv
echo "=== ncu full: tcgen05 GEMMs + attention at batch 8 (after the epilogue rework)"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|attn_f32_kernel" -c 40 -o gpurun_out/r02_prof_gemm_b8 -f \
    python tools/run_batch.py 8 2 1 > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?"
python tools/ncu_summary.py gpurun_out/r02_prof_gemm_b8.ncu-rep gpurun_out/r02_gemm_attn_b8_final_ncu.txt | cut -c1-330
rm -f gpurun_out/r02_prof_gemm_b8.ncu-rep
