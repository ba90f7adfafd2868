"""CPU: the oracle restatement vs the committed HF-generated golden vectors
(tests/golden/hf_pin.npz, produced by oracle/pin_against_hf.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from qwen3_asr_rs_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "hf_pin.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def model(gold):
    cfg = O.cfg_tiny()
    return O.OracleModel(cfg, synth.make_weights(cfg, int(gold["seed"])))


def test_filterbank_matches_hf(gold):
    assert np.abs(O.mel_filterbank() - gold["filterbank_hf"]).max() < 1e-7


def test_mel_matches_hf(gold):
    idx, sec = gold["mel_clip"]
    x = synth.make_clip(int(idx), float(sec))
    mel = O.extract_mel(x).numpy()
    assert mel.shape == gold["mel_hf"].shape == (128, (len(x) + 159) // 160)
    # HF computes the STFT with numpy (f64 FFT), the reference/oracle with an fp32 FFT
    assert np.abs(mel - gold["mel_hf"]).max() < 1e-4


def test_encoder_matches_hf(gold, model):
    idx, sec = gold["enc_clip"]
    x = synth.make_clip(int(idx), float(sec))
    mel = O.extract_mel(x)
    _, valid = model.chunk_plan(mel.shape[1])
    assert len(valid) == 12 and model.window_mask(sum(valid), valid) is not None   # windows active
    enc = model.encode(mel).numpy()
    assert enc.shape == gold["enc_hf"].shape
    assert np.abs(enc - gold["enc_hf"]).max() < 2e-5


def test_decoder_greedy_matches_hf(gold, model):
    idx, sec = gold["enc_clip"]
    x = synth.make_clip(int(idx), float(sec))
    steps = len(gold["dec_ids_hf"])
    r = O.transcribe_ids(model, x, max_new_tokens=steps, keep_logits=True)
    assert r.ids == gold["dec_ids_hf"].tolist()
    stride = int(gold["vocab_stride"])
    logits = np.stack([r.prefill_logits.numpy()] + [l.numpy() for l in r.step_logits])[:, ::stride]
    assert np.abs(logits - gold["dec_logits_hf"]).max() < 1e-4


def test_last_only_lm_head_is_identical(model):
    x = synth.make_clip(9, 2.0)
    a = O.transcribe_ids(model, x, max_new_tokens=4, lm_head_all_rows=True)
    b = O.transcribe_ids(model, x, max_new_tokens=4, lm_head_all_rows=False)
    assert a.ids == b.ids
    assert np.abs(a.prefill_logits.numpy() - b.prefill_logits.numpy()).max() < 1e-5
