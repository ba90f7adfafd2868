"""Pins oracle/oracle.py against the model authors' independent HF implementations and
writes tests/golden/hf_pin.npz.   TEST INFRASTRUCTURE -- run in the build container:

    python oracle/pin_against_hf.py            # regenerates tests/golden/hf_pin.npz

The reference has no golden vectors for this path and cannot be
built here (Rust); HF transformers 5.5 ships the same architecture written by the model
authors: WhisperFeatureExtractor (mel; the reference claims to match it, src/mel.rs:39-45),
Qwen3OmniMoeAudioEncoder (audio tower) and Qwen3ForCausalLM (decoder).  The fixture holds
HF's OUTPUTS for seeded inputs (weights from qwen3_asr_rs_b200.synth, clips from
synth.make_clip); tests/test_oracle_pin.py re-runs the oracle on the same inputs.

Caveat handled here: transformers 5.5.0's eager audio attention ignores the window
blocks (modeling_qwen3_omni_moe.py:754-758 passes no mask); the model's own
``_prepare_attention_mask`` is injected into each layer so HF honours its cu_seqlens.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O                      # noqa: E402
from qwen3_asr_rs_b200 import synth                 # noqa: E402

SEED = 7
MEL_CLIP = (3, 3.07)        # (clip index, seconds)  -> ragged length, 308 frames
ENC_CLIP = (4, 11.55)       # 12 chunks (> 8 -> window mask active), tail chunk of 55 frames
DEC_STEPS = 12
VOCAB_STRIDE = 97


def hf_mel(x: np.ndarray) -> np.ndarray:
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor(feature_size=128, sampling_rate=16000, hop_length=160, n_fft=400)
    # the reference pads to a multiple of hop (src/mel.rs:51-53); HF pads to 30 s when called
    # through __call__, so use the underlying routine on the hop-padded waveform.
    n = ((len(x) + 159) // 160) * 160
    xp = np.zeros(n, np.float32)
    xp[: len(x)] = x
    return np.asarray(fe._np_extract_fbank_features(xp[None, :], "cpu"))[0].astype(np.float32)


def hf_encoder(cfg: O.AsrCfg, w, mel: torch.Tensor) -> np.ndarray:
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeAudioEncoderConfig
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import Qwen3OmniMoeAudioEncoder
    a = cfg.audio
    hc = Qwen3OmniMoeAudioEncoderConfig(
        num_mel_bins=a.num_mel_bins, encoder_layers=a.encoder_layers,
        encoder_attention_heads=a.encoder_attention_heads, encoder_ffn_dim=a.encoder_ffn_dim,
        d_model=a.d_model, max_source_positions=a.max_source_positions, n_window=a.n_window,
        output_dim=a.output_dim, n_window_infer=a.n_window_infer,
        downsample_hidden_size=a.downsample_hidden_size)
    hc._attn_implementation = "eager"
    enc = Qwen3OmniMoeAudioEncoder(hc).eval().to(torch.float32)
    sd = {k[len("thinker.audio_tower."):]: v.float() for k, v in w.items() if k.startswith("thinker.audio_tower.")}
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("positional_embedding" in m for m in missing), missing

    def inject_mask(module, args, kwargs):
        hs, cu = args[0], args[1]
        kwargs["attention_mask"] = enc._prepare_attention_mask(hs, cu)
        return args, kwargs
    for layer in enc.layers:
        layer.register_forward_pre_hook(inject_mask, with_kwargs=True)
    with torch.no_grad():
        out = enc(mel, feature_lens=torch.tensor([mel.shape[1]]))
    return out.last_hidden_state.numpy()


def hf_decoder(cfg: O.AsrCfg, w, hidden: torch.Tensor, steps: int):
    from transformers import Qwen3Config, Qwen3ForCausalLM
    t = cfg.text
    hc = Qwen3Config(vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
                     num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
                     num_key_value_heads=t.num_key_value_heads, head_dim=t.head_dim, rms_norm_eps=t.rms_norm_eps,
                     rope_theta=t.rope_theta, tie_word_embeddings=True, attention_bias=False,
                     max_position_embeddings=4096, use_sliding_window=False)
    hc.rope_parameters = {"rope_type": "default", "rope_theta": t.rope_theta}
    hc._attn_implementation = "eager"
    lm = Qwen3ForCausalLM(hc).eval().to(torch.float32)
    sd = {"model." + k[len("thinker.model."):]: v.float() for k, v in w.items() if k.startswith("thinker.model.")}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    lm.load_state_dict(sd, strict=True)
    ids, logits_rows = [], []
    with torch.no_grad():
        out = lm(inputs_embeds=hidden, use_cache=True)
        past = out.past_key_values
        nxt = out.logits[:, -1, :]
        logits_rows.append(nxt[0].numpy().copy())
        for _ in range(steps):
            tok = int(nxt.argmax(-1)[0])
            ids.append(tok)
            out = lm(input_ids=torch.tensor([[tok]]), past_key_values=past, use_cache=True)
            past = out.past_key_values
            nxt = out.logits[:, -1, :]
            logits_rows.append(nxt[0].numpy().copy())
    return ids, np.stack(logits_rows)


def main(out_path: str) -> None:
    torch.manual_seed(0)
    cfg = O.cfg_tiny()
    w = synth.make_weights(cfg, SEED)
    model = O.OracleModel(cfg, w)

    x_mel = synth.make_clip(*MEL_CLIP)
    mel_hf = hf_mel(x_mel)
    mel_or = O.extract_mel(x_mel).numpy()
    print("mel  oracle-vs-HF max abs diff", np.abs(mel_hf - mel_or).max(), mel_hf.shape)
    fb_hf = None
    from transformers import WhisperFeatureExtractor
    fb_hf = WhisperFeatureExtractor(feature_size=128).mel_filters.T.astype(np.float32)
    print("filterbank oracle-vs-HF max abs diff", np.abs(fb_hf - O.mel_filterbank()).max())

    x_enc = synth.make_clip(*ENC_CLIP)
    mel_enc = O.extract_mel(x_enc)
    enc_hf = hf_encoder(cfg, w, mel_enc)
    enc_or = model.encode(mel_enc).numpy()
    print("enc  oracle-vs-HF max abs diff", np.abs(enc_hf - enc_or).max(), enc_hf.shape,
          "scale", np.abs(enc_hf).max())

    # decoder: prompt + injected (oracle) audio embeddings -> HF greedy loop with KV cache
    ids, a0 = O.build_prompt(enc_or.shape[0])
    hidden = model.embed(ids).unsqueeze(0)
    hidden[0, a0:a0 + enc_or.shape[0]] = torch.from_numpy(enc_or)
    dec_ids, dec_logits = hf_decoder(cfg, w, hidden, DEC_STEPS)
    r = O.transcribe_ids(model, x_enc, max_new_tokens=DEC_STEPS, keep_logits=True)
    or_logits = np.stack([r.prefill_logits.numpy()] + [l.numpy() for l in r.step_logits])
    print("dec  ids equal:", dec_ids == r.ids, dec_ids)
    print("dec  logits oracle-vs-HF max abs diff", np.abs(dec_logits - or_logits).max(),
          "scale", np.abs(dec_logits).max())

    np.savez_compressed(
        out_path,
        seed=np.int64(SEED), mel_clip=np.array(MEL_CLIP), enc_clip=np.array(ENC_CLIP),
        mel_hf=mel_hf, filterbank_hf=fb_hf, enc_hf=enc_hf.astype(np.float32),
        dec_ids_hf=np.array(dec_ids, np.int64), dec_logits_hf=dec_logits[:, ::VOCAB_STRIDE].astype(np.float32),
        vocab_stride=np.int64(VOCAB_STRIDE))
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(os.path.join(ROOT, "tests", "golden", "hf_pin.npz"))
