import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# scene kwargs of each golden case -- must match oracle/make_golden.py::CASES
CASE_SCENES = {
    "plain": dict(kind="thuman", seed=0),
    "stress": dict(kind="thuman", seed=1, novel_pose=True),
    "h36m": dict(kind="h36m", seed=2, H=500, W=500, novel_pose=True, t_vertices_from="file"),
}


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
    """(scene, state_dict, golden npz dict) for a committed golden case."""
    if name not in _cache:
        from mpsnerf_b200 import synthetic
        g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        scene = synthetic.make_scene(**CASE_SCENES[name])
        sd = synthetic.seeded_state_dict(scene.seed, float(g["alpha_gain"]))
        _cache[name] = (scene, sd, g)
    return _cache[name]


@pytest.fixture(params=["plain", "stress", "h36m"])
def case(request):
    return (request.param,) + load_case(request.param)
