"""Minimal audio ingest for the CLI: WAV (PCM 8/16/24/32-bit or float32) -> mono f32 @ 16 kHz.

The reference decodes with FFmpeg/swresample or hound + rubato sinc (src/audio.rs:7-245); neither is reproducible
here, so parity of the hot path starts at the f32 16 kHz sample vector.  This loader uses
scipy's polyphase resampler; it is host-side I/O outside the timed region.
"""
from __future__ import annotations

import struct
import wave
from math import gcd

import numpy as np


def load_wav(path: str, target_rate: int = 16000) -> np.ndarray:
    with wave.open(path, "rb") as w:
        nch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        x = v.astype(np.float32) / 8388608.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise ValueError(f"unsupported sample width {width}")
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1)                 # mono mixdown (audio.rs:193-206)
    if rate != target_rate:
        from scipy.signal import resample_poly
        g = gcd(rate, target_rate)
        x = resample_poly(x.astype(np.float64), target_rate // g, rate // g).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def read_wav_pcm(path: str):
    """WAV payload as-is for the GPU ingest path (asrb_ingest_pcm): (interleaved array [frames, channels] of int16 / int32 /
    float32, sample rate).  8-bit and 24-bit files are widened to int32 on the host (rare formats; hound does the same
    widening, src/audio.rs:181-189)."""
    with wave.open(path, "rb") as w:
        nch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2")
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4")
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (np.where(v & 0x800000, v - (1 << 24), v) << 8).astype(np.int32)      # same value / 2^23 as int32 / 2^31
    elif width == 1:
        x = ((np.frombuffer(raw, dtype=np.uint8).astype(np.int32) - 128) << 24).astype(np.int32)
    else:
        raise ValueError(f"unsupported sample width {width}")
    return np.ascontiguousarray(x.reshape(-1, nch)), rate
