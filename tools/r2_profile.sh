#!/bin/bash
# round-2 ncu evidence: launch list of the default bench + --set full captures of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== ncu launch list (bench b1, 8 new tokens)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_b1.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline --workload b1 > gpurun_out/ncu_bench.log 2>&1
echo "exit $?"; wc -l gpurun_out/r02_launches_b1.csv
echo "=== ncu full: fused decode step (batch 1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 8 -c 1 -o gpurun_out/r02_prof_decode_b1 -f \
    python tools/run_batch.py 1 16 1 > gpurun_out/ncu_b1.log 2>&1; echo "exit $?"
echo "=== ncu full: batched decode step (batch 8)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_batch_kernel -s 8 -c 1 -o gpurun_out/r02_prof_decode_b8 -f \
    python tools/run_batch.py 8 16 1 > gpurun_out/ncu_b8.log 2>&1; echo "exit $?"
echo "=== ncu full: tcgen05 GEMMs + attention at batch 8 (first encoder layers + conv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_f32_kernel" -c 24 -o gpurun_out/r02_prof_gemm_b8 -f \
    python tools/run_batch.py 8 2 1 > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?"
echo "=== ncu full: bandwidth kernels at batch 8"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mel_|conv1_gelu|layernorm_s3|rmsnorm_s3|qk_norm_rope|embed_inject|splitk_reduce" -c 14 -o gpurun_out/r02_prof_bw_b8 -f \
    python tools/run_batch.py 8 2 1 > gpurun_out/ncu_bw.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
