"""CPU: host-side helpers around the hot path -- parse_asr_output (src/inference.rs:276-305) and WAV ingest."""
import wave

import numpy as np
import pytest

from qwen3_asr_rs_b200.audio import load_wav
from qwen3_asr_rs_b200.text import capitalize_first, parse_asr_output


@pytest.mark.parametrize("raw,forced,exp", [
    ("language English<asr_text>Hello there.", False, ("English", "Hello there.")),
    ("  language Chinese <asr_text> 你好 ", False, ("Chinese", "你好")),
    ("language English Hello there.", False, ("English", "Hello there.")),
    ("Hello there.", False, ("unknown", "Hello there.")),
    ("  some text ", True, ("forced", "some text")),
    ("language ", False, ("unknown", "language")),
])
def test_parse_asr_output(raw, forced, exp):
    assert parse_asr_output(raw, forced) == exp


def test_capitalize_first():
    assert capitalize_first("english") == "English" and capitalize_first("") == "" and capitalize_first("étoile") == "Étoile"


def test_load_wav_resamples_24k_stereo_to_16k_mono(tmp_path):
    rate, secs = 24000, 0.5
    t = np.arange(int(rate * secs)) / rate
    left = 0.4 * np.sin(2 * np.pi * 440 * t)
    right = 0.2 * np.sin(2 * np.pi * 440 * t)
    pcm = (np.stack([left, right], 1) * 32767).astype("<i2")
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(rate); w.writeframes(pcm.tobytes())
    x = load_wav(str(p))
    assert x.dtype == np.float32 and abs(len(x) - int(16000 * secs)) <= 1
    ref = 0.3 * np.sin(2 * np.pi * 440 * np.arange(len(x)) / 16000)
    assert np.abs(x[200:-200] - ref[200:-200]).max() < 5e-3
