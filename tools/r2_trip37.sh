#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== batch tests (tiny)"; timeout 90 python -m pytest tests/ -q -m gpu --timeout 60 -k "batch_step_tiny or batch_larger_than_8" 2>&1 | tail -2
echo "=== batch roofline 8"; timeout 110 python tools/batch_decode_roofline.py 8 > gpurun_out/batch_decode8.json 2> gpurun_out/batch_decode8.err; echo "exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/batch_decode8.json'):
    r = json.loads(l)
    print(r['batch'], {k: (round(v['us_per_step'],1), round(v['frac_of_hbm_peak'],3), round(v['rtf'])) for k, v in r.items() if isinstance(v, dict) and 'us_per_step' in v}, r.get('ids_batch_equal_per_seq'), r.get('first_mismatch'))
PY
