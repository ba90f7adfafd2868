#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export ASRB_BATCH_FLAGS=4
echo "=== batch tests (tiny), flags=$ASRB_BATCH_FLAGS"; timeout 80 python -m pytest tests/ -q -m gpu --timeout 60 -k "batch_step_tiny" 2>&1 | tail -2
echo "=== batch roofline 8, flags=$ASRB_BATCH_FLAGS"; timeout 100 python tools/batch_decode_roofline.py 8 > gpurun_out/batch_decode8.json 2> gpurun_out/batch_decode8.err; echo "exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/batch_decode8.json'):
    r = json.loads(l)
    print(r['batch'], {k: (round(v['us_per_step'],1), round(v['frac_of_hbm_peak'],3), round(v['rtf'])) for k, v in r.items() if isinstance(v, dict) and 'us_per_step' in v}, r.get('ids_batch_equal_per_seq'), r.get('first_mismatch'))
PY
