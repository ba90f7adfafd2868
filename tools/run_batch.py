"""Minimal driver for profiling: Qwen3-ASR-0.6B, batch B x 30 s clips, N new tokens, R repetitions (no timing claims)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_rs_b200 import AsrInference, config_0p6b, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
R = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = config_0p6b()
eng = AsrInference.from_weights(cfg, synth.make_weights(cfg, 1), device=0)
clips = [synth.make_clip(i, 30.0) for i in range(B)]
for _ in range(R):
    r = eng.transcribe_ids(clips, max_new_tokens=N)
print("stage_ms", r.stage_ms, "stats", eng.stats())
eng.close()
