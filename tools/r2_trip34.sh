#!/bin/bash
cd "$(dirname "$0")/.."
cp experiments/_exp_libasr_b200.so qwen3_asr_rs_b200/libasr_b200.so   # experimental build (masked re-poll gather + runtime L2 prefetch distance)
mkdir -p gpurun_out
for pf in 0 2 4 8 16; do
  echo "=== ASRB_MEGA_PF=$pf"; ASRB_MEGA_PF=$pf timeout 300 python bench.py --workload b1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rtf', round(d['value'],1), 'decode us/step', round(d['decode']['us_per_step'],1), 'stage', d['stage_ms'])"
done
echo "=== timeline PF=4"; ASRB_MEGA_PF=4 timeout 300 python tools/mega_timeline.py 2>&1 | grep "us/step\|layer-5 detail\|P4 consume" | cut -c1-400
