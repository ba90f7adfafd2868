#!/bin/bash
# experiment: fused decode step under different ASRB_MEGA_FLAGS (see decode_mega.cu Params::flags)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in ${FLAGS:-0 1 2 3}; do
  echo "=== flags $f"
  ASRB_MEGA_FLAGS=$f timeout 300 python tools/mega_timeline.py > gpurun_out/mega_flags_$f.txt 2>&1; echo "exit $?"
  grep -E "us/step|p[1-5]_|detail" gpurun_out/mega_flags_$f.txt | head -14
  ASRB_MEGA_FLAGS=$f timeout 300 python -m pytest tests/ -q -m gpu --timeout 300 -k "eos or ids_exact" 2>&1 | tail -1
done
