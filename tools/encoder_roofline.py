"""Encoder GEMM throughput at batch B (BASELINE.json configs[2] shape): algorithmic encoder FLOPs / encoder stage time."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_rs_b200 import AsrInference, config_0p6b, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = config_0p6b()
eng = AsrInference.from_weights(cfg, synth.make_weights(cfg, 1), device=0)
clips = [synth.make_clip(i, 30.0) for i in range(B)]
GFLOP_PER_CLIP = 270.7      # SURVEY.md section 8d (conv1 0.83 + conv2 99.53 + conv3 25.88 + conv_out 5.37 + 18 layers 135.26 + attn 2.49 + head 1.34)
PREFILL_GFLOP = 356.7 + 18.8 + 0.31
out = {}
for planes in os.environ.get("PLANES", "3,2,1").split(","):
    eng.set_option("planes", planes)
    for _ in range(2):
        r = eng.transcribe_ids(clips, max_new_tokens=2)
    enc_ms, pre_ms = r.stage_ms["encoder"], r.stage_ms["prefill"]
    out[f"planes{planes}"] = {"batch": B, "encoder_ms": enc_ms, "encoder_algorithmic_tflops": GFLOP_PER_CLIP * B / enc_ms,
                              "prefill_ms": pre_ms, "prefill_algorithmic_tflops": PREFILL_GFLOP * B / pre_ms,
                              "mma_tflops_issued_encoder": GFLOP_PER_CLIP * B / enc_ms * int(planes)}
print(json.dumps(out, indent=1))
eng.close()
