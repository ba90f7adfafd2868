#!/bin/bash
# round-2 ncu evidence: launch list of the default bench + --set full captures of the dominant kernels.
# gpurun_out/ may carry at most 64 MiB back: the big reports are summarised on the box (tools/ncu_summary.py) and removed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== ncu launch list (bench b1, 8 new tokens)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_b1.csv \
    python bench.py --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline --workload b1 > gpurun_out/ncu_bench.log 2>&1
echo "exit $?"; wc -l gpurun_out/r02_launches_b1.csv
echo "=== ncu full: fused decode step (batch 1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 8 -c 1 -o gpurun_out/r02_prof_decode_b1 -f \
    python tools/run_batch.py 1 16 1 > gpurun_out/ncu_b1.log 2>&1; echo "exit $?"
python tools/ncu_summary.py gpurun_out/r02_prof_decode_b1.ncu-rep gpurun_out/r02_decode_b1_ncu.txt | cut -c1-400
echo "=== ncu full: batched decode step (batch 8)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_batch_kernel -s 8 -c 1 -o gpurun_out/r02_prof_decode_b8 -f \
    python tools/run_batch.py 8 16 1 > gpurun_out/ncu_b8.log 2>&1; echo "exit $?"
python tools/ncu_summary.py gpurun_out/r02_prof_decode_b8.ncu-rep gpurun_out/r02_decode_b8_ncu.txt | cut -c1-400
echo "=== ncu full: tcgen05 GEMMs + attention at batch 8"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|attn_f32_kernel" -c 40 -o gpurun_out/r02_prof_gemm_b8 -f \
    python tools/run_batch.py 8 2 1 > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?"
python tools/ncu_summary.py gpurun_out/r02_prof_gemm_b8.ncu-rep gpurun_out/r02_gemm_attn_b8_ncu.txt | cut -c1-300
rm -f gpurun_out/r02_prof_gemm_b8.ncu-rep
echo "=== ncu full: bandwidth kernels at batch 8"
timeout 900 ncu --set full --clock-control none -k regex:"mel_|conv1_gelu|layernorm_s3|rmsnorm_s3|qk_norm_rope|embed_inject|splitk_reduce|ingest" -c 14 -o gpurun_out/r02_prof_bw_b8 -f \
    python tools/run_batch.py 8 2 1 > gpurun_out/ncu_bw.log 2>&1; echo "exit $?"
python tools/ncu_summary.py gpurun_out/r02_prof_bw_b8.ncu-rep gpurun_out/r02_bw_kernels_b8_ncu.txt | cut -c1-300
rm -f gpurun_out/r02_prof_bw_b8.ncu-rep
du -sh gpurun_out
