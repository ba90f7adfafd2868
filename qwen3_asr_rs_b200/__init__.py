"""B200-native Qwen3-ASR hot path (mel -> audio encoder -> greedy decode) behind a C ABI.

The compute lives in ``libasr_b200.so`` (hand-written sm_100a CUDA, see csrc/); this package
is the thin host-side mirror of the reference's ``AsrInference`` API for that path.
"""
from .config import AsrConfig, AudioEncoderConfig, TextDecoderConfig, config_0p6b, config_1p7b, config_tiny  # noqa: F401
from .inference import AsrInference, TranscribeIds  # noqa: F401
