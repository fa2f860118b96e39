"""Training step (BASELINE config 4): the CPU oracle's gradients against the reference's own autograd
(tests/golden/train_grads.npz, oracle/make_golden_train.py), and the host-side bucket logic."""
import os

import numpy as np
import torch

from oracle import oracle as O
from oracle import train_oracle as TO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_grads.npz")


def check_grads_against_golden(grads, g, rtol_norm, rtol_val):
    """grads: {name: numpy array}.  Per live parameter: L2 norm and 512 strided entries vs the reference's."""
    worst = 0.0
    for k in TO.dense_keys() + TO.TRUNK_KEYS:
        got = np.asarray(grads[k].detach().cpu() if torch.is_tensor(grads[k]) else grads[k], dtype=np.float64).reshape(-1)
        n_ref = float(g["norm/" + k])
        assert n_ref > 0, k
        assert abs(np.linalg.norm(got) - n_ref) <= rtol_norm * n_ref, (k, np.linalg.norm(got), n_ref)
        sel, val = g["idx/" + k], g["val/" + k].astype(np.float64)
        err = np.abs(got[sel] - val).max() / max(np.abs(val).max(), 1e-30)
        worst = max(worst, err)
        assert err <= rtol_val, (k, err)
    return worst


def test_oracle_training_gradients_match_reference_autograd():
    from golden_cases import build_train_case
    g = np.load(GOLD)
    scene, sd, ids, S, u, target, msk = build_train_case()
    loss, grads, out = TO.loss_and_grads(O.smpl_tensors(scene.smpl), sd, scene.sp_input, scene.tp_input, scene.rays_o[ids],
                                         scene.rays_d[ids], scene.near[ids], scene.far[ids], S, target, msk, u=u)
    assert out["n_active"] == int(g["n_active"])
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(out["rgb_map"], g["rgb_map"], atol=5e-4)
    np.testing.assert_allclose(out["acc_map"], g["acc_map"], atol=2e-4)
    # two fp32 CPU implementations of the same graph: gradient norms agree to ~1e-4; single entries move by up to
    # ~7e-3 of the largest entry where a pre-activation sits within rounding of a ReLU's kink (a sample's whole
    # contribution to that entry switches on or off)
    check_grads_against_golden(grads, g, rtol_norm=1e-3, rtol_val=2e-2)


def test_dense_key_order_is_the_engine_order():
    from mpsnerf_b200.engine import DENSE_FP32_ORDER
    assert TO.dense_keys() == list(DENSE_FP32_ORDER) and len(DENSE_FP32_ORDER) == 46
