#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest ids"; timeout 600 python -m pytest tests/ -q -m gpu --timeout 600 -k "ids or eos" > gpurun_out/pytest_quick.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_quick.log
for v in ${PFS:-16 4 0}; do
  echo "=== ASRB_MEGA_PF=$v"; ASRB_MEGA_PF=$v timeout 300 python tools/mega_timeline.py > gpurun_out/mega_timeline_pf$v.txt 2>&1; echo "exit $?"
  grep -E "us/step|cta0|flags_wait|gather_plain|layer total|detail" gpurun_out/mega_timeline_pf$v.txt | head -8
  grep -A8 "cross-CTA" gpurun_out/mega_timeline_pf$v.txt
done
