#!/bin/bash
# round-2 final validation of the committed build: full GPU suite, smoke, both bench arms, per-batch decode roofline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 2400 python -m pytest tests/ -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -6 gpurun_out/pytest_gpu.log
cp gpurun_out/parity_report.json gpurun_out/parity_all.json 2>/dev/null
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== decode step roofline at batch 1 / 8 / 16"; timeout 900 python tools/batch_decode_roofline.py 1 8 16 > gpurun_out/batch_decode.json 2> gpurun_out/batch_decode.err; echo "exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/batch_decode.json'):
    r = json.loads(l)
    print(r['batch'], {k: (round(v['us_per_step'],1), round(v['frac_of_hbm_peak'],3), round(v['rtf'])) for k, v in r.items() if isinstance(v, dict) and 'us_per_step' in v}, r.get('ids_batch_equal_per_seq'), r.get('first_mismatch'))
PY
echo "=== encoder roofline at batch 8 / 64"; timeout 600 python tools/encoder_roofline.py 8 > gpurun_out/enc_b8.json 2> gpurun_out/enc_b8.err; echo "exit $?"; grep "encoder_ms\|prefill_ms\|mma_tflops" gpurun_out/enc_b8.json
PLANES=3 timeout 600 python tools/encoder_roofline.py 64 > gpurun_out/enc_b64.json 2> gpurun_out/enc_b64.err; echo "exit $?"; grep "encoder_ms\|prefill_ms\|mma_tflops" gpurun_out/enc_b64.json
