"""Model configuration -- mirrors /root/reference/src/config.rs (serde structs; every
field defaults to the Qwen3-ASR-0.6B value, config.rs:52-62, 90-99)."""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from typing import Tuple


@dataclass
class AudioEncoderConfig:            # config.rs:27-62
    d_model: int = 896
    encoder_layers: int = 18
    encoder_attention_heads: int = 14
    encoder_ffn_dim: int = 3584
    num_mel_bins: int = 128
    max_source_positions: int = 1500
    n_window: int = 50
    n_window_infer: int = 800
    conv_chunksize: int = 500
    downsample_hidden_size: int = 480
    output_dim: int = 1024


@dataclass
class TextDecoderConfig:             # config.rs:66-113
    vocab_size: int = 151936
    hidden_size: int = 1024
    intermediate_size: int = 3072
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    tie_word_embeddings: bool = True
    mrope_section: Tuple[int, ...] = (24, 20, 20)
    mrope_interleaved: bool = False


@dataclass
class AsrConfig:                     # config.rs:5-24 (thinker_config.{audio_config,text_config})
    audio: AudioEncoderConfig = field(default_factory=AudioEncoderConfig)
    text: TextDecoderConfig = field(default_factory=TextDecoderConfig)

    @staticmethod
    def from_dict(d: dict) -> "AsrConfig":
        th = d["thinker_config"]
        a = {k: v for k, v in th.get("audio_config", {}).items() if k in AudioEncoderConfig.__dataclass_fields__}
        tj = dict(th.get("text_config", {}))
        rs = tj.pop("rope_scaling", None) or {}
        t = {k: v for k, v in tj.items() if k in TextDecoderConfig.__dataclass_fields__}
        if "mrope_section" in rs:
            t["mrope_section"] = tuple(rs["mrope_section"])
        t["mrope_interleaved"] = bool(rs.get("mrope_interleaved", False) or rs.get("interleaved", False))
        return AsrConfig(AudioEncoderConfig(**a), TextDecoderConfig(**t))

    @staticmethod
    def from_file(path: str) -> "AsrConfig":      # config.rs:116-120
        with open(path) as f:
            return AsrConfig.from_dict(json.load(f))

    def to_config_json(self) -> dict:
        t = asdict(self.text)
        sec, inter = t.pop("mrope_section"), t.pop("mrope_interleaved")
        t["rope_scaling"] = {"rope_type": "default", "mrope_section": list(sec), "mrope_interleaved": bool(inter)}
        return {"thinker_config": {"audio_config": asdict(self.audio), "text_config": t,
                                   "audio_start_token_id": 151669, "audio_end_token_id": 151670,
                                   "audio_token_id": 151676}}


def config_0p6b() -> AsrConfig:
    return AsrConfig()


def config_1p7b() -> AsrConfig:
    """Dims recalled from the HF model card; a real config.json overrides."""
    return AsrConfig(AudioEncoderConfig(d_model=1024, encoder_layers=24, encoder_attention_heads=16,
                                        encoder_ffn_dim=4096, output_dim=2048),
                     TextDecoderConfig(hidden_size=2048, intermediate_size=6144))


def config_tiny() -> AsrConfig:
    """Small same-structure config for fast parity tests (vocab kept: prompt ids are > 151000)."""
    return AsrConfig(AudioEncoderConfig(d_model=128, encoder_layers=2, encoder_attention_heads=2,
                                        encoder_ffn_dim=256, downsample_hidden_size=32, output_dim=256),
                     TextDecoderConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=3,
                                       num_attention_heads=4, num_key_value_heads=2, head_dim=128))
