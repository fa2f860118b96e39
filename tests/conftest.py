import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")

from golden_cases import CASES, build_case, smpl_models  # noqa: E402  (table shared with oracle/make_golden.py)

SINGLE_CASES = [n for n, c in CASES.items() if len(c["scenes"]) == 1]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


_cache = {}


def load_case(name):
    """(scene, state_dict, golden npz dict) for a committed single-subject golden case."""
    if name not in _cache:
        g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        scenes, sd, _, _ = build_case(CASES[name])
        assert len(scenes) == 1
        _cache[name] = (scenes[0], sd, g)
    return _cache[name]


def load_batch_case(name):
    """(scenes, state_dict, sp_input, tp_input, smpl models by gender, golden) for a B > 1 golden case."""
    key = ("batch", name)
    if key not in _cache:
        g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        scenes, sd, sp, tp = build_case(CASES[name])
        _cache[key] = (scenes, sd, sp, tp, smpl_models(CASES[name]), g)
    return _cache[key]


@pytest.fixture(params=SINGLE_CASES)
def case(request):
    return (request.param,) + load_case(request.param)
