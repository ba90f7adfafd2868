"""Debug: per-phase clock64 timeline of the fused decode step (ASRB_MEGA_DEBUG=1)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["ASRB_MEGA_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_asr_rs_b200 import AsrInference, _lib, config_0p6b, synth  # noqa: E402

cfg = config_0p6b()
eng = AsrInference.from_weights(cfg, synth.make_weights(cfg, 1), device=0)
x = synth.make_clip(0, 30.0)
for _ in range(2):
    r = eng.transcribe_ids([x], max_new_tokens=32)
print("stage_ms", r.stage_ms, "us/step", 1e3 * r.stage_ms["decode"] / max(r.decode_steps, 1))
lib = _lib.load_library()
buf = (C.c_longlong * 2048)()
n = lib.asrb_debug_mega_timeline(buf, 2048)
t = np.array(buf[:], dtype=np.int64).reshape(2, -1)
L = cfg.text.num_hidden_layers
names = ["p1_qkv", "p2_attn", "p3_oproj", "p4_gateup", "p5_down"]   # each includes the wait for its inputs
NP = len(names)
for cta, row in zip(("cta0", "ctaLast"), t):
    marks = row[: 2 + NP * L]
    d = np.diff(marks)
    per = d[: NP * L].reshape(L, NP)
    print(cta, "total cycles", int(marks[-1] - marks[0]), " lm_head cycles", int(d[NP * L]))
    print("  mean cycles per phase over layers:")
    for i, nm in enumerate(names):
        print(f"    {nm:10s} mean {per[:, i].mean():9.0f}  min {per[:, i].min():7d}  max {per[:, i].max():7d}")
    print("  layer totals (first 4):", per.sum(1)[:4])
for cta, row in zip(("cta0", "ctaLast"), t):
    f = row[400:424].astype(np.int64)
    d = lambda a, b: int(f[a] - f[b])
    print(cta, "layer-5 detail (cycles): P1 wait", d(1, 0), "gather", d(2, 1), "norm", d(3, 2), "gemv", d(4, 3),
          "| P4 wait", d(9, 8), "gather", d(10, 9), "norm", d(11, 10), "gemv", d(12, 11), "arrive", d(13, 12),
          "| P5 wait", d(17, 16), "gather", d(18, 17), "gemv", d(19, 18), "arrive", d(20, 19))
for cta, row in zip(("cta0", "ctaLast"), t):
    f = row[400:440].astype(np.int64)
    if f[24]:
        names2 = ["sync+poll qkv", "kv tile+sync", "append+sync", "warp partials+sync", "combine+publish", "release"]
        print(cta, "layer-5 P2/P3 detail (cycles):", ", ".join(f"{nm} {int(f[25 + i] - f[24 + i])}" for i, nm in enumerate(names2)),
              "| merge: poll", int(f[35] - f[30]), "math+publish", int(f[31] - f[35]), "| P3: gather", int(f[33] - f[32]), "o_proj", int(f[34] - f[33]))
for cta, row in zip(("cta0", "ctaLast"), t):
    f = row[440:464].astype(np.int64)
    n = int((f != 0).sum())
    print(cta, "P4 consume warp-0 marks (cycles from entry; entry, xr+sync, then per slot [weights ready, rows done], end):",
          [int(v - f[0]) for v in f[:n]])
g = t[0][512:512 + 3 * 148].reshape(148, 3).astype(np.int64)
t0 = g[:, 0].min()
print("layer-5 wall clock per CTA (ns after the first CTA left P1): [P1 done, partial published, P3 gather done]")
for c in list(range(0, 72)) + [100, 147]:
    print(f"  cta {c:3d}: {int(g[c,0]-t0):6d} {int(g[c,1]-t0) if g[c,1] else -1:6d} {int(g[c,2]-t0):6d}")
eng.close()
