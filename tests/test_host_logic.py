"""CPU: host-side logic -- config parsing (src/config.rs), chunk/window plan
(src/audio_encoder.rs:83-121,172-209,263-266), prompt (src/inference.rs:215-257)."""
import json

import numpy as np
import pytest

from oracle import oracle as O
from qwen3_asr_rs_b200 import config as Cfg, synth


def test_config_defaults_are_0p6b():
    c = Cfg.AsrConfig.from_dict({"thinker_config": {"audio_config": {}, "text_config": {}}})
    assert c.audio.d_model == 896 and c.audio.n_window_infer == 800 and c.text.num_key_value_heads == 8
    assert c.text.mrope_section == (24, 20, 20) and not c.text.mrope_interleaved


def test_config_roundtrip(tmp_path):
    c = Cfg.config_1p7b()
    p = tmp_path / "config.json"
    p.write_text(json.dumps(c.to_config_json()))
    c2 = Cfg.AsrConfig.from_file(str(p))
    assert c2 == c
    o = O.AsrCfg.from_config_json(json.loads(p.read_text()))
    assert o.text.hidden_size == 2048 and o.audio.encoder_layers == 24


@pytest.mark.parametrize("frames,exp", [(100, 13), (1, 1), (2, 1), (8, 1), (9, 2), (16, 2), (55, 7), (60, 8), (99, 13)])
def test_feat_extract_output_length(frames, exp):
    assert O.feat_extract_output_length(frames) == exp


def test_sample_fixture_shapes_from_survey():
    # sample1 -> 800 frames / 104 tokens / no mask; 30 s -> 390 tokens, windows 104/104/104/78
    m = O.OracleModel(O.cfg_0p6b(), {})
    _, v = m.chunk_plan(800)
    assert sum(v) == 104 and m.window_mask(104, v) is None
    _, v = m.chunk_plan(416)
    assert sum(v) == 54
    _, v = m.chunk_plan(3000)
    assert sum(v) == 390
    mask = m.window_mask(390, v)
    blocks = (mask[0, 0] == 0).sum(1)
    assert blocks[0] == 104 and blocks[389] == 78


def test_prompt_layout():
    ids, a0 = O.build_prompt(5)
    assert len(ids) == 5 + 15 and a0 == 9 and ids[a0:a0 + 5] == [O.AUDIO_PAD] * 5 and ids[-2:] == [77091, 198]
    ids2, _ = O.build_prompt(5, [11528, 6364])
    assert ids2[:-2] == ids and ids2[-2:] == [11528, 6364]


def test_synth_weights_are_bf16_exact_and_named():
    cfg = O.cfg_tiny()
    w = synth.make_weights(cfg, 0)
    assert "thinker.audio_tower.layers.1.fc1.weight" in w and "thinker.model.layers.2.self_attn.q_norm.weight" in w
    assert w["thinker.audio_tower.conv_out.weight"].shape == (128, 32 * 16)
    import torch
    assert all(t.dtype == torch.bfloat16 for t in w.values())


def test_make_clip_deterministic():
    a, b = synth.make_clip(3, 1.0), synth.make_clip(3, 1.0)
    assert a.dtype == np.float32 and len(a) == 16000 and np.array_equal(a, b) and abs(np.abs(a).max() - 0.5) < 1e-6
