import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # gpu tests need a CUDA device; on the CPU container they are deselected by -m "not gpu",
    # and if someone runs them anyway without a GPU they fail loudly (no skip, no fallback).
    pass


_REPORT = {}


@pytest.fixture(scope="session")
def report():
    """Collects measured deviations; written to gpurun_out/parity_report.json at session end."""
    yield _REPORT
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="session")
def tiny():
    """(cfg, weights, oracle model) for the tiny same-structure config, seed 7."""
    from oracle import oracle as O
    from qwen3_asr_rs_b200 import synth
    cfg = O.cfg_tiny()
    w = synth.make_weights(cfg, 7)
    return cfg, w, O.OracleModel(cfg, w)


@pytest.fixture(scope="session")
def tiny_engine(tiny):
    from qwen3_asr_rs_b200 import AsrInference, config_tiny
    _, w, _ = tiny
    eng = AsrInference.from_weights(config_tiny(), w, device=0)
    yield eng
    eng.close()
