"""Host-side mirror of the reference's public API for the hot path.

Reference: ``AsrInference::{load, transcribe}`` (/root/reference/src/inference.rs:19-213).
``transcribe()`` steps 2-8 (inference.rs:94-200: samples -> mel -> encoder -> prompt ->
prefill -> greedy ids) run inside libasr_b200.so; step 1 (file decode / resample,
src/audio.rs) and step 9 (detokenise, src/tokenizer.rs) are outside the hot path, so the
entry points here take f32 16 kHz samples and return token ids.  All arithmetic happens
in the CUDA library through the C ABI (include/asr_b200.h); numpy is only the container
for host buffers.  No fallback exists: a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .config import AsrConfig

EOS_TOKEN_IDS = (151643, 151645)      # inference.rs:154
MAX_NEW_TOKENS = 4096                 # inference.rs:153
MEL_SAMPLE_RATE = 16000               # inference.rs:16

_DT = {"float32": 0, "bfloat16": 1, "float16": 2}


def _dims_struct(cfg: AsrConfig) -> _lib.AsrbDims:
    a, t = cfg.audio, cfg.text
    d = _lib.AsrbDims()
    for k in ("d_model", "encoder_layers", "encoder_attention_heads", "encoder_ffn_dim", "num_mel_bins",
              "max_source_positions", "n_window", "n_window_infer", "downsample_hidden_size", "output_dim"):
        setattr(d, k, int(getattr(a, k)))
    for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
              "num_key_value_heads", "head_dim"):
        setattr(d, k, int(getattr(t, k)))
    d.tie_word_embeddings = int(bool(t.tie_word_embeddings))
    d.rms_norm_eps = float(t.rms_norm_eps)
    d.rope_theta = float(t.rope_theta)
    return d


@dataclass
class TranscribeResult:           # inference.rs:270-274
    text: str
    language: str
    raw_output: str
    ids: List[int]


@dataclass
class TranscribeIds:
    ids: List[List[int]]          # generated ids per utterance, EOS excluded
    stage_ms: Dict[str, float]    # device time per stage (CUDA events)
    kernels_launched: int
    decode_steps: int


class AsrInference:
    """``AsrInference`` of the reference, hot path only, batch-capable."""

    def __init__(self, cfg: AsrConfig, ctx, model):
        self.config = cfg
        self._lib = _lib.load_library()
        self._ctx, self._model = ctx, model
        self._session = None
        self._cap = None
        self._options: Dict[str, str] = {}
        self.tokenizer = None     # text.AsrTokenizer when model_dir has tokenizer.json

    # ---- construction ------------------------------------------------------------
    @staticmethod
    def _init_ctx(device: int):
        lib = _lib.load_library()
        ctx = C.c_void_p()
        _lib.check(lib.asrb_init(int(device), C.byref(ctx)))
        return lib, ctx

    @classmethod
    def load(cls, model_dir: str, device: int = 0) -> "AsrInference":
        """AsrInference::load (inference.rs:30-86): config.json + safetensors from ``model_dir``."""
        lib, ctx = cls._init_ctx(device)
        model = C.c_void_p()
        _lib.check(lib.asrb_model_load(ctx, os.fsencode(model_dir), C.byref(model)))
        d = _lib.AsrbDims()
        _lib.check(lib.asrb_model_dims(model, C.byref(d)))
        cfg = AsrConfig.from_file(os.path.join(model_dir, "config.json"))
        eng = cls(cfg, ctx, model)
        if os.path.exists(os.path.join(model_dir, "tokenizer.json")):
            from .text import AsrTokenizer
            eng.tokenizer = AsrTokenizer.from_dir(model_dir)
        return eng

    @classmethod
    def from_weights(cls, cfg: AsrConfig, weights: Dict[str, "object"], device: int = 0) -> "AsrInference":
        """Build from in-memory tensors (name -> torch tensor or numpy array) -- what load()
        does after reading safetensors (weights.rs:62-120).  bf16 torch tensors pass through
        bit-exactly."""
        lib, ctx = cls._init_ctx(device)
        model = C.c_void_p()
        dims = _dims_struct(cfg)
        _lib.check(lib.asrb_model_create(ctx, C.byref(dims), C.byref(model)))
        for name, t in weights.items():
            if hasattr(t, "detach"):          # torch tensor
                import torch
                t = t.detach().cpu().contiguous()
                if t.dtype == torch.bfloat16:
                    arr, code = t.view(torch.int16).numpy(), 1
                elif t.dtype == torch.float16:
                    arr, code = t.view(torch.int16).numpy(), 2
                else:
                    arr, code = t.to(torch.float32).numpy(), 0
            else:
                arr = np.ascontiguousarray(t, dtype=np.float32)
                code = 0
            shape = (C.c_int64 * arr.ndim)(*arr.shape)
            _lib.check(lib.asrb_model_set_tensor(model, name.encode(), code, shape, arr.ndim,
                                                 arr.ctypes.data_as(C.c_void_p)))
        _lib.check(lib.asrb_model_finalize(model))
        return cls(cfg, ctx, model)

    def close(self) -> None:
        if self._session is not None:
            self._lib.asrb_session_free(self._session)
            self._session = None
        if self._model is not None:
            self._lib.asrb_model_free(self._model)
            self._model = None
        if self._ctx is not None:
            self._lib.asrb_ctx_free(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- session management ----------------------------------------------------------
    def set_option(self, key: str, value: str) -> None:
        self._options[key] = value
        if self._session is not None:
            _lib.check(self._lib.asrb_session_set_option(self._session, key.encode(), value.encode()))

    def device_ids(self):
        """(ids_ptr, lens_ptr, row_stride, batch): device addresses of the last generated ids (asrb_session_device_ids)."""
        ids, lens = C.c_void_p(), C.c_void_p()
        stride, batch = C.c_int(), C.c_int()
        _lib.check(self._lib.asrb_session_device_ids(self._session, C.byref(ids), C.byref(lens), C.byref(stride), C.byref(batch)))
        return ids.value, lens.value, stride.value, batch.value

    def stats(self) -> Dict[str, int]:
        """Decoder-forward / GEMM path counters (asrb_session_stats): fallbacks are visible, never silent."""
        out = (C.c_int64 * 5)()
        if self._session is None:
            return {}
        _lib.check(self._lib.asrb_session_stats(self._session, out, 5))
        return dict(zip(("decode_batch_steps", "decode_fused_steps", "decode_phase_steps", "gemm_simt_fallbacks",
                         "gemm_tc_launches"), [int(v) for v in out]))

    def _ensure_session(self, batch: int, max_samples: int, max_lang: int, max_new: int):
        cap = self._cap
        if cap is None or batch > cap[0] or max_samples > cap[1] or max_lang > cap[2] or max_new > cap[3]:
            if self._session is not None:
                _lib.check(self._lib.asrb_session_free(self._session))
                self._session = None
            new_cap = (max(batch, cap[0] if cap else 0), max(max_samples, cap[1] if cap else 0),
                       max(max_lang, cap[2] if cap else 0), max(max_new, cap[3] if cap else 0))
            s = C.c_void_p()
            _lib.check(self._lib.asrb_session_create(self._model, new_cap[0], new_cap[1], new_cap[2], new_cap[3],
                                                     C.byref(s)))
            self._session, self._cap = s, new_cap
            for k, v in self._options.items():
                _lib.check(self._lib.asrb_session_set_option(s, k.encode(), v.encode()))
        return self._session

    @staticmethod
    def _pack_samples(clips: Sequence[np.ndarray]):
        arrs = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        ptrs = (C.POINTER(C.c_float) * len(arrs))(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
        lens = (C.c_int64 * len(arrs))(*[a.shape[0] for a in arrs])
        return arrs, ptrs, lens

    @staticmethod
    def _pack_lang(language_ids, batch):
        if language_ids is None:
            return None, None, None, 0
        keep, ptrs, lens, mx = [], [], [], 0
        for ids in language_ids:
            if ids is None:
                ptrs.append(C.POINTER(C.c_int64)())
                lens.append(0)
            else:
                a = np.ascontiguousarray(ids, dtype=np.int64)
                keep.append(a)
                ptrs.append(a.ctypes.data_as(C.POINTER(C.c_int64)))
                lens.append(len(a))
                mx = max(mx, len(a))
        return keep, (C.POINTER(C.c_int64) * batch)(*ptrs), (C.c_int32 * batch)(*lens), mx

    # ---- the hot path ----------------------------------------------------------------
    def transcribe_ids(self, clips: Sequence[np.ndarray], language_ids: Optional[Sequence] = None,
                       max_new_tokens: int = MAX_NEW_TOKENS) -> TranscribeIds:
        """transcribe() steps 2-8 for a batch: host f32 samples in, host token ids out."""
        B = len(clips)
        arrs, ptrs, lens = self._pack_samples(clips)
        keep, lptrs, llens, mx = self._pack_lang(language_ids, B)
        s = self._ensure_session(B, max(a.shape[0] for a in arrs), mx, max_new_tokens)
        ids = np.zeros((B, max_new_tokens), dtype=np.int32)
        n = np.zeros(B, dtype=np.int32)
        _lib.check(self._lib.asrb_transcribe_ids(
            s, ptrs, lens, B, lptrs, llens, int(max_new_tokens),
            ids.ctypes.data_as(C.POINTER(C.c_int32)), n.ctypes.data_as(C.POINTER(C.c_int32))))
        ms = (C.c_float * 6)()
        k, st = C.c_int64(), C.c_int64()
        _lib.check(self._lib.asrb_last_timings(s, ms, C.byref(k), C.byref(st)))
        names = ("h2d", "mel", "encoder", "prefill", "decode", "total")
        return TranscribeIds([ids[b, : n[b]].tolist() for b in range(B)], dict(zip(names, ms)), k.value, st.value)

    # ---- GPU-side audio ingest (step 1, src/audio.rs:162-245) -------------------------------------------
    _PCM_FMT = {"int16": 0, "float32": 1, "int32": 2}

    def _ingest(self, pcms: Sequence, rates: Sequence[int], max_lang: int, max_new: int):
        B = len(pcms)
        arrs = [np.ascontiguousarray(a if a.ndim == 2 else a.reshape(-1, 1)) for a in pcms]
        for a in arrs:
            if a.dtype.name not in self._PCM_FMT:
                raise ValueError(f"PCM dtype must be int16 / int32 / float32, got {a.dtype}")
        n_out = [-(-a.shape[0] * MEL_SAMPLE_RATE // int(r)) for a, r in zip(arrs, rates)]
        s = self._ensure_session(B, max(n_out), max_lang, max_new)
        ptrs = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
        frames = (C.c_int64 * B)(*[a.shape[0] for a in arrs])
        chans = (C.c_int32 * B)(*[a.shape[1] for a in arrs])
        rts = (C.c_int32 * B)(*[int(r) for r in rates])
        fmts = (C.c_int32 * B)(*[self._PCM_FMT[a.dtype.name] for a in arrs])
        n = (C.c_int64 * B)()
        _lib.check(self._lib.asrb_ingest_pcm(s, ptrs, frames, chans, rts, fmts, B, n))
        return s, arrs, list(n)

    def ingest_pcm(self, pcms: Sequence, rates: Sequence[int]) -> List[np.ndarray]:
        """asrb_ingest_pcm + read-back (tests): interleaved PCM arrays [frames, channels] -> mono f32 @ 16 kHz, on the GPU."""
        s, _keep, n = self._ingest(pcms, rates, 16, 64)
        out = []
        for b in range(len(pcms)):
            a = np.empty(n[b], dtype=np.float32)
            _lib.check(self._lib.asrb_ingested_read(s, b, a.ctypes.data_as(C.POINTER(C.c_float))))
            out.append(a)
        return out

    def transcribe_pcm(self, pcms: Sequence, rates: Sequence[int], language_ids: Optional[Sequence] = None,
                       max_new_tokens: int = MAX_NEW_TOKENS) -> TranscribeIds:
        """transcribe() steps 1-8 for a batch with step 1 on the GPU: raw PCM in, token ids out."""
        B = len(pcms)
        keep, lptrs, llens, mx = self._pack_lang(language_ids, B)
        s, _arrs, _n = self._ingest(pcms, rates, mx, max_new_tokens)
        ids = np.zeros((B, max_new_tokens), dtype=np.int32)
        n = np.zeros(B, dtype=np.int32)
        _lib.check(self._lib.asrb_transcribe_ingested(s, lptrs, llens, int(max_new_tokens),
                                                      ids.ctypes.data_as(C.POINTER(C.c_int32)), n.ctypes.data_as(C.POINTER(C.c_int32))))
        ms = (C.c_float * 6)()
        k, st = C.c_int64(), C.c_int64()
        _lib.check(self._lib.asrb_last_timings(s, ms, C.byref(k), C.byref(st)))
        names = ("h2d", "mel", "encoder", "prefill", "decode", "total")
        return TranscribeIds([ids[b, : n[b]].tolist() for b in range(B)], dict(zip(names, ms)), k.value, st.value)

    def transcribe(self, audio_path: str, language: Optional[str] = None,
                   max_new_tokens: int = MAX_NEW_TOKENS, gpu_ingest: bool = True) -> TranscribeResult:
        """AsrInference::transcribe (inference.rs:89-213): step 1 (WAV payload -> mono 16 kHz; on the GPU by default,
        `gpu_ingest=False` = the host loader) -> steps 2-8 on the GPU -> step 9 (detokenise + parse, host; needs
        tokenizer.json, else raw_output is the id list as text)."""
        from .audio import load_wav, read_wav_pcm
        from .text import language_prompt_ids, parse_asr_output
        lang_ids = language_prompt_ids(self.tokenizer, language)
        if gpu_ingest:
            pcm, rate = read_wav_pcm(audio_path)
            r = self.transcribe_pcm([pcm], [rate], language_ids=[lang_ids] if lang_ids is not None else None,
                                    max_new_tokens=max_new_tokens)
        else:
            samples = load_wav(audio_path, MEL_SAMPLE_RATE)
            r = self.transcribe_ids([samples], language_ids=[lang_ids] if lang_ids is not None else None,
                                    max_new_tokens=max_new_tokens)
        ids = r.ids[0]
        raw = self.tokenizer.decode(ids) if self.tokenizer is not None else " ".join(str(i) for i in ids)
        lang, text = parse_asr_output(raw, language is not None) if self.tokenizer is not None else ("unknown", raw)
        return TranscribeResult(text=text, language=lang, raw_output=raw, ids=ids)

    # ---- stage-level calls (the calls transcribe() makes; used by the parity tests) ----
    def mel(self, clips: Sequence[np.ndarray], max_new_tokens: int = 64, max_lang: int = 16) -> List[np.ndarray]:
        """WhisperFeatureExtractor::extract (mel.rs:49-96) -> [128, F] per utterance."""
        B = len(clips)
        arrs, ptrs, lens = self._pack_samples(clips)
        s = self._ensure_session(B, max(a.shape[0] for a in arrs), max_lang, max_new_tokens)
        frames = (C.c_int64 * B)()
        _lib.check(self._lib.asrb_mel(s, ptrs, lens, B, frames))
        out = []
        nm = self.config.audio.num_mel_bins
        for b in range(B):
            a = np.empty((nm, frames[b]), dtype=np.float32)
            _lib.check(self._lib.asrb_mel_read(s, b, a.ctypes.data_as(C.POINTER(C.c_float))))
            out.append(a)
        self._B = B
        return out

    def encode(self) -> List[np.ndarray]:
        """AudioEncoder::forward (audio_encoder.rs:79-169) on the mel of the last mel() call."""
        B = self._B
        toks = (C.c_int64 * B)()
        _lib.check(self._lib.asrb_encode(self._session, toks))
        out = []
        for b in range(B):
            a = np.empty((toks[b], self.config.audio.output_dim), dtype=np.float32)
            _lib.check(self._lib.asrb_encode_read(self._session, b, a.ctypes.data_as(C.POINTER(C.c_float))))
            out.append(a)
        return out

    def prefill(self, language_ids: Optional[Sequence] = None, want_logits: bool = True):
        """prompt + embed/inject + MRoPE + prefill (inference.rs:105-149) -> (seq_lens, last-row logits)."""
        B = self._B
        keep, lptrs, llens, mx = self._pack_lang(language_ids, B)
        seq = (C.c_int64 * B)()
        logits = np.empty((B, self.config.text.vocab_size), dtype=np.float32) if want_logits else None
        _lib.check(self._lib.asrb_prefill(self._session, lptrs, llens, seq,
                                          logits.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None))
        return list(seq), logits

    def decode_step(self, want_logits: bool = True):
        """One greedy iteration (inference.rs:160-200) -> (next ids, logits after the forward)."""
        B = self._B
        nxt = (C.c_int64 * B)()
        logits = np.empty((B, self.config.text.vocab_size), dtype=np.float32) if want_logits else None
        _lib.check(self._lib.asrb_decode_step(self._session, nxt,
                                              logits.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None))
        return list(nxt), logits

    def generate(self, max_new_tokens: int) -> List[List[int]]:
        B = self._B
        ids = np.zeros((B, max_new_tokens), dtype=np.int32)
        n = np.zeros(B, dtype=np.int32)
        _lib.check(self._lib.asrb_generate(self._session, int(max_new_tokens),
                                           ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                           n.ctypes.data_as(C.POINTER(C.c_int32))))
        return [ids[b, : n[b]].tolist() for b in range(B)]
