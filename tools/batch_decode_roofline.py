"""Decode-step HBM roofline at batch B per GPU (north_star: "batch 8 x 30 s"; BASELINE configs[4]: 16 sequences at a
512-token KV cache): algorithmic bytes of one decoder forward step (bench.decode_step_bytes: all weights once + B x
fp32 KV cache + append) / measured step time.  Also checks that the batch-aware fused step (decode_batch.cu) yields the
same ids as one fused launch per sequence (decode_mega.cu)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_rs_b200 import AsrInference, config_0p6b, synth
from bench import decode_step_bytes

NEW = int(os.environ.get("NEW_TOKENS", "64"))
cfg = config_0p6b()
eng = AsrInference.from_weights(cfg, synth.make_weights(cfg, 1), device=0)
peak = 6581.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = []
for B in [int(a) for a in sys.argv[1:]] or [1, 8, 16]:
    clips = [synth.make_clip(i, 30.0) for i in range(B)]
    row = {"batch": B}
    ids = {}
    for mode in ((("1", "batch"),) if B > 1 else ()) + (("0", "per_seq"),):
        eng.set_option("batch_step", mode[0])
        for _ in range(2):
            r = eng.transcribe_ids(clips, max_new_tokens=NEW)
        ids[mode[1]] = r.ids
        steps = max(r.decode_steps, 1)
        us = 1e3 * r.stage_ms["decode"] / steps
        ctx = 405 + NEW / 2          # prompt of a 30 s clip (390 audio tokens + 15) + half the generated tokens
        by = decode_step_bytes(cfg, ctx, B)
        row[mode[1]] = {"us_per_step": us, "tokens_per_s": B * 1e6 / us, "bytes_per_step": by,
                        "achieved_gbps": by / us / 1e3, "frac_of_hbm_peak": by / us / 1e3 / peak,
                        "rtf": 30.0 * B / (r.stage_ms["total"] / 1e3), "stage_ms": r.stage_ms}
    if B > 1:
        row["ids_batch_equal_per_seq"] = ids["batch"] == ids["per_seq"]
        row["first_mismatch"] = next(((b, i) for b in range(B) for i in range(min(len(ids["batch"][b]), len(ids["per_seq"][b])))
                                      if ids["batch"][b][i] != ids["per_seq"][b][i]), None)
    row["stats"] = eng.stats()
    out.append(row)
    print(json.dumps(row), flush=True)
eng.close()
