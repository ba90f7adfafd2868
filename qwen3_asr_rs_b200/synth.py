"""Seeded synthetic Qwen3-ASR checkpoints and clips (no weights / no network exist here).

Tensor names and shapes are exactly the ones the reference loads
(/root/reference/src/audio_encoder.rs:37-55, src/layers.rs:135-150,185-227,262-281,
388-439, src/text_decoder.rs:54-79); values are bf16-representable so that the
reference's bf16->f32 up-cast (src/weights.rs:134-142) and our bf16 device storage
hold identical numbers.  Harness utility: torch is used only as a seeded RNG.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
import torch


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16)


def make_weights(cfg, seed: int = 0, peaked_head: bool = False) -> Dict[str, torch.Tensor]:
    """Returns name -> bf16 tensor.  ``cfg`` is any object with .audio / .text dims
    (oracle.AsrCfg or qwen3_asr_rs_b200.config.AsrConfig).

    Random weights give near-flat logits: the top-1/top-2 gap of a tied random head is a few 1e-5 of max|logit| at
    its worst step, the same size as fp32 summation-order noise, so exact-id parity becomes a coin toss for reasons
    unrelated to kernel correctness.  ``peaked_head=True`` (needs
    ``cfg.text.tie_word_embeddings == False``, the untied key of src/text_decoder.rs:75-79) draws the lm_head rows
    with log-normal norms (sigma 0.5): logits are dominated by a few hundred large-norm rows, the winner still depends
    on the hidden state's direction (dozens of distinct ids per 128 tokens), and the measured worst gap is >= 2e-4 of
    max|logit| -- an order of magnitude above the noise.  The embedding stays plain, so there is no self-token
    fixed point (a tied peaked head repeats one id forever)."""
    g = torch.Generator().manual_seed(seed)
    a, t = cfg.audio, cfg.text
    w: Dict[str, torch.Tensor] = {}

    def rn(*shape, std):
        return _bf16(torch.randn(*shape, generator=g, dtype=torch.float32) * std)

    def norm_w(n):
        return _bf16(1.0 + 0.1 * torch.randn(n, generator=g, dtype=torch.float32))

    p = "thinker.audio_tower"
    dsh, d, ffn = a.downsample_hidden_size, a.d_model, a.encoder_ffn_dim
    w[f"{p}.conv2d1.weight"] = rn(dsh, 1, 3, 3, std=0.45)
    w[f"{p}.conv2d1.bias"] = rn(dsh, std=0.1)
    for name in ("conv2d2", "conv2d3"):
        w[f"{p}.{name}.weight"] = rn(dsh, dsh, 3, 3, std=(2.0 / (9 * dsh)) ** 0.5)
        w[f"{p}.{name}.bias"] = rn(dsh, std=0.1)
    o = lambda l: (l - 1) // 2 + 1
    feat = dsh * o(o(o(a.num_mel_bins)))
    w[f"{p}.conv_out.weight"] = rn(d, feat, std=(1.0 / feat) ** 0.5)
    es = 1.2 / d ** 0.5
    for i in range(a.encoder_layers):
        q = f"{p}.layers.{i}"
        for ln in ("self_attn_layer_norm", "final_layer_norm"):
            w[f"{q}.{ln}.weight"] = norm_w(d)
            w[f"{q}.{ln}.bias"] = rn(d, std=0.05)
        for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[f"{q}.self_attn.{pr}.weight"] = rn(d, d, std=es * (2.0 if pr in ("q_proj", "k_proj") else 0.7))
            w[f"{q}.self_attn.{pr}.bias"] = rn(d, std=0.02)
        w[f"{q}.fc1.weight"] = rn(ffn, d, std=es)
        w[f"{q}.fc1.bias"] = rn(ffn, std=0.02)
        w[f"{q}.fc2.weight"] = rn(d, ffn, std=0.7 / ffn ** 0.5)
        w[f"{q}.fc2.bias"] = rn(d, std=0.02)
    w[f"{p}.ln_post.weight"] = norm_w(d)
    w[f"{p}.ln_post.bias"] = rn(d, std=0.05)
    w[f"{p}.proj1.weight"] = rn(d, d, std=es)
    w[f"{p}.proj1.bias"] = rn(d, std=0.02)
    w[f"{p}.proj2.weight"] = rn(a.output_dim, d, std=es)
    w[f"{p}.proj2.bias"] = rn(a.output_dim, std=0.02)

    p = "thinker.model"
    H, I, hd = t.hidden_size, t.intermediate_size, t.head_dim
    nq, nkv = t.num_attention_heads, t.num_key_value_heads
    hs = 1.0 / H ** 0.5
    w[f"{p}.embed_tokens.weight"] = rn(t.vocab_size, H, std=0.1)
    for i in range(t.num_hidden_layers):
        q = f"{p}.layers.{i}"
        w[f"{q}.input_layernorm.weight"] = norm_w(H)
        w[f"{q}.post_attention_layernorm.weight"] = norm_w(H)
        w[f"{q}.self_attn.q_proj.weight"] = rn(nq * hd, H, std=hs)
        w[f"{q}.self_attn.k_proj.weight"] = rn(nkv * hd, H, std=hs)
        w[f"{q}.self_attn.v_proj.weight"] = rn(nkv * hd, H, std=hs)
        w[f"{q}.self_attn.o_proj.weight"] = rn(H, nq * hd, std=0.1 / (nq * hd) ** 0.5)
        w[f"{q}.self_attn.q_norm.weight"] = norm_w(hd)
        w[f"{q}.self_attn.k_norm.weight"] = norm_w(hd)
        w[f"{q}.mlp.gate_proj.weight"] = rn(I, H, std=hs)
        w[f"{q}.mlp.up_proj.weight"] = rn(I, H, std=hs)
        w[f"{q}.mlp.down_proj.weight"] = rn(H, I, std=2.0 / I ** 0.5)
    w[f"{p}.norm.weight"] = norm_w(H)
    if not t.tie_word_embeddings:
        if peaked_head:
            g2 = torch.Generator().manual_seed(4242 + seed)
            e = torch.randn(t.vocab_size, H, generator=g2, dtype=torch.float32) * 0.1
            w["thinker.lm_head.weight"] = _bf16(e * torch.exp(0.5 * torch.randn(t.vocab_size, 1, generator=g2, dtype=torch.float32)))
        else:
            w["thinker.lm_head.weight"] = rn(t.vocab_size, H, std=0.1)
    else:
        assert not peaked_head, "peaked_head needs an untied lm_head (cfg.text.tie_word_embeddings = False)"
    return w


def write_checkpoint(model_dir: str, cfg, weights: Dict[str, torch.Tensor], shards: int = 1) -> None:
    """config.json + model.safetensors (or sharded + index.json), bf16 on disk like the
    real checkpoints (/root/reference/src/weights.rs:10-58)."""
    from safetensors.torch import save_file
    os.makedirs(model_dir, exist_ok=True)
    with open(os.path.join(model_dir, "config.json"), "w") as f:
        json.dump(cfg.to_config_json(), f, indent=1)
    names = sorted(weights)
    if shards <= 1:
        save_file({k: weights[k].contiguous() for k in names}, os.path.join(model_dir, "model.safetensors"))
        return
    per = (len(names) + shards - 1) // shards
    wm = {}
    for s in range(shards):
        part = names[s * per:(s + 1) * per]
        fn = f"model-{s + 1:05d}-of-{shards:05d}.safetensors"
        save_file({k: weights[k].contiguous() for k in part}, os.path.join(model_dir, fn))
        wm.update({k: fn for k in part})
    with open(os.path.join(model_dir, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": wm}, f)


def make_clip(index: int, seconds: float = 30.0, sample_rate: int = 16000) -> np.ndarray:
    """Synthetic speech-like clip i: 6 amplitude-modulated
    harmonics of f0 in U[90,250] Hz + N(0, 0.01^2) noise, peak-normalised to 0.5."""
    rng = np.random.default_rng(1234 + index)
    n = int(round(seconds * sample_rate))
    tt = np.arange(n, dtype=np.float64) / sample_rate
    f0 = rng.uniform(90.0, 250.0)
    x = np.zeros(n, dtype=np.float64)
    for h in range(1, 7):
        am = 0.5 * (1.0 + np.sin(2 * np.pi * rng.uniform(1.0, 5.0) * tt + rng.uniform(0, 2 * np.pi)))
        x += (1.0 / h) * am * np.sin(2 * np.pi * f0 * h * tt + rng.uniform(0, 2 * np.pi))
    x += rng.normal(0.0, 0.01, n) * np.abs(x).max() / 0.5
    x *= 0.5 / np.abs(x).max()
    return x.astype(np.float32)
