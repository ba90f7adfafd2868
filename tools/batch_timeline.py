"""Debug: per-phase clock64 timeline of the batch-aware fused decode step (ASRB_MEGA_DEBUG=batch)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["ASRB_MEGA_DEBUG"] = "batch"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_rs_b200 import AsrInference, _lib, config_0p6b, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = config_0p6b()
eng = AsrInference.from_weights(cfg, synth.make_weights(cfg, 1), device=0)
clips = [synth.make_clip(i, 30.0) for i in range(B)]
for _ in range(2):
    r = eng.transcribe_ids(clips, max_new_tokens=32)
print("batch", B, "stage_ms", r.stage_ms, "us/step", 1e3 * r.stage_ms["decode"] / max(r.decode_steps, 1))
lib = _lib.load_library()
buf = (C.c_longlong * 4096)()
n = lib.asrb_debug_mega_timeline(buf, 4096)
t = np.array(buf[:2048], dtype=np.int64).reshape(2, -1)
gt = np.array(buf[2048:2048 + 8 * 148], dtype=np.int64).reshape(148, 8)
L = cfg.text.num_hidden_layers
names = ["x gather+norm", "qkv gemv", "attn items", "attn merge", "o_proj", "x gather+norm", "gate/up gemv", "down"]
NP = len(names)
for cta, row in zip(("cta0", "ctaLast(merger)"), t):
    marks = row[: 1 + NP * L + 2]
    d = np.diff(marks)
    per = d[: NP * L].reshape(L, NP)
    print(cta, "total cycles", int(marks[-1] - marks[0]), " final gather", int(d[NP * L]), " lm_head cycles", int(d[NP * L + 1]))
    for i, nm in enumerate(names):
        print(f"    {nm:14s} mean {per[2:, i].mean():9.0f}  min {per[2:, i].min():7d}  max {per[2:, i].max():7d}")
    print("    layer mean", per[2:].sum(1).mean(), " first layers:", per.sum(1)[:3])
for cta, row in zip(("cta0", "ctaLast"), t):
    f = row[600:800]
    n = int((f != 0).sum())
    print(cta, "layer-5 fine marks (cycles since first):", [int(v - f[0]) for v in f[:n]])
names_gt = ["x1 done", "qkv done", "items done", "merge done", "o_proj done", "gate/up done"]
t0 = gt[:, 0].min()
print("layer-5 wall clock per CTA (ns after the first CTA finished its x gather):", names_gt)
for c in list(range(0, 148, 6)) + [140, 143, 146, 147]:
    print(f"  cta {c:3d}:", " ".join(f"{int(v - t0):7d}" for v in gt[c, :6]))
print("  max over CTAs:", " ".join(f"{int(v - t0):7d}" for v in gt[:, :6].max(0)), " argmax", gt[:, :6].argmax(0))
print("  min over CTAs:", " ".join(f"{int(v - t0):7d}" for v in gt[:, :6].min(0)))
eng.close()
