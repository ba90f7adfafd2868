"""CPU: the C-ABI library loads and exports every symbol include/asr_b200.h declares;
the product path fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    return sorted(set(re.findall(r"ASRB_API\s+[\w\s\*]+?\b(asrb_\w+)\s*\(", src)))


def test_header_symbols_exported():
    from qwen3_asr_rs_b200 import _lib
    lib = _lib.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in asr_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_version_and_defaults_no_gpu_needed():
    from qwen3_asr_rs_b200 import _lib
    lib = _lib.load_library()
    assert b"sm_100a" in lib.asrb_version()
    d = _lib.AsrbDims()
    assert lib.asrb_dims_default(C.byref(d)) == 0
    assert (d.d_model, d.encoder_layers, d.hidden_size, d.num_hidden_layers, d.vocab_size) == (896, 18, 1024, 28, 151936)
    assert d.rope_theta == 1e6 and d.rms_norm_eps == 1e-6


def test_null_arguments_return_status_not_crash():
    from qwen3_asr_rs_b200 import _lib
    lib = _lib.load_library()
    assert lib.asrb_dims_default(None) != 0
    assert b"null" in lib.asrb_last_error()
    assert lib.asrb_session_create(None, 1, 16000, 0, 8, None) != 0


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "qwen3_asr_rs_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
