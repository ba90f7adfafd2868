#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== debug batch"; timeout 240 python tools/debug_batch.py 2>&1 | tail -20
echo "=== timeline B=8 flags=1"; ASRB_BATCH_FLAGS=1 timeout 300 python tools/batch_timeline.py 8 2>&1 | tail -26
echo "=== timeline B=8 flags=0"; ASRB_BATCH_FLAGS=0 timeout 300 python tools/batch_timeline.py 8 2>&1 | grep -E "us/step|total cycles|layer mean"
echo "=== timeline B=16 flags=1"; ASRB_BATCH_FLAGS=1 timeout 300 python tools/batch_timeline.py 16 2>&1 | grep -E "us/step|total cycles|layer mean"
echo "=== batch roofline 0.6B"; timeout 900 python tools/batch_decode_roofline.py 8 16 > gpurun_out/batch_decode.json 2> gpurun_out/batch_decode.err; echo "exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/batch_decode.json'):
    r = json.loads(l)
    print(r['batch'], {k: (round(v['us_per_step'],1), round(v['frac_of_hbm_peak'],3), round(v['rtf'])) for k, v in r.items() if isinstance(v, dict) and 'us_per_step' in v}, r.get('ids_batch_equal_per_seq'), r.get('first_mismatch'))
PY
