import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


_ENGINES = {}


@pytest.fixture(scope="session")
def engine_factory():
    """Session-cached engines keyed by (cfg name, weight seed, max_batch)."""
    from mcm_b200 import synth
    from mcm_b200.engine import McmEngine

    def make(cfg_name, wseed=5, max_batch=64):
        key = (cfg_name, wseed, max_batch)
        if key not in _ENGINES:
            cfg = synth.CFGS[cfg_name]
            sd = synth.synth_vision_state_dict(cfg, wseed)
            _ENGINES[key] = (McmEngine.from_state_dict(sd, cfg, max_batch=max_batch), sd, cfg)
        return _ENGINES[key]

    yield make
    for e, _, _ in _ENGINES.values():
        e.close()
    _ENGINES.clear()
