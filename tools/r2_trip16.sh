#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest gpu (without the oracle-heavy full-size tests)"; timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 -k "not full_size" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_gpu_quick.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== batch roofline 0.6B"; timeout 900 python tools/batch_decode_roofline.py 8 16 > gpurun_out/batch_decode.json 2> gpurun_out/batch_decode.err; echo "exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/batch_decode.json'):
    r = json.loads(l)
    print(r['batch'], {k: (round(v['us_per_step'],1), round(v['frac_of_hbm_peak'],3), round(v['rtf'])) for k, v in r.items() if isinstance(v, dict) and 'us_per_step' in v}, r.get('ids_batch_equal_per_seq'), r.get('first_mismatch'))
PY
