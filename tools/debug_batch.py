"""Debug: batch-aware fused step vs one fused launch per sequence on the tiny config (ids must be identical)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qwen3_asr_rs_b200 import AsrInference, config_tiny, synth
from oracle import oracle as O
cfg = O.cfg_tiny()
w = synth.make_weights(cfg, 7)
eng = AsrInference.from_weights(config_tiny(), w, device=0)
def run(clips, lang=None, n=12):
    out = {}
    for mode in ("1", "0"):
        eng.set_option("batch_step", mode)
        out[mode] = eng.transcribe_ids(clips, language_ids=lang, max_new_tokens=n).ids
    bad = [(b, next((i for i in range(len(out["1"][b])) if i >= len(out["0"][b]) or out["1"][b][i] != out["0"][b][i]), None)) for b in range(len(clips))]
    return [x for x in bad if x[1] is not None]
cases = {
 "B3_nolang": ([2.5, 9.1, 5.0], None),
 "B3_lang": ([2.5, 9.1, 5.0], [None, [11528, 6364], [11528, 8453, 55]]),
 "B2": ([2.5, 5.0], None),
 "B2_a": ([2.5, 2.5], None),
 "B2_b": ([5.0, 5.0], None),
 "B2_c": ([5.0, 2.5], None),
 "B2_d": ([9.1, 2.5], None),
 "B2_e": ([0.7, 0.9], None),
 "B3_x": ([2.5, 5.0, 2.5], None),
 "B2_f": ([1.0, 5.0], None),
 "B2_g": ([2.5, 9.1], None),
 "B2_h": ([2.5, 4.5], None),
 "B2_i": ([3.0, 5.0], None),
 "B3_y": ([2.5, 5.0, 5.0], None),
 "B8": ([1.1, 2.3, 0.7, 4.9, 3.1, 1.9, 2.2, 0.9], None),
 "B8_same": ([3.0] * 8, None),
 "B4_short": ([0.7, 0.7, 0.7, 0.7], None),
 "B9": ([1.1, 2.3, 0.7, 4.9, 3.1, 1.9, 2.2, 0.9, 5.3], None),
 "B11": ([1.1, 2.3, 0.7, 4.9, 3.1, 1.9, 2.2, 0.9, 5.3, 1.4, 2.8], None),
 "B16_long": ([12.0] * 16, None),
 "B5_long": ([30.0, 12.0, 20.0, 3.0, 25.0], None),
}
for name, (secs, lang) in cases.items():
    clips = [synth.make_clip(80 + i, s) for i, s in enumerate(secs)]
    print(name, "mismatches (seq, first index):", run(clips, lang), flush=True)
print(eng.stats())
eng.close()
