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


def test_bucket_automatic_step_follows_the_reference_loop():
    """DenseBucket without TrainStep -- the reference's own loop body (render -> loss.backward() -> optimizer.step(),
    run_nerf_batch.py:544-563): the first render node opens the step, a callback on the autograd engine closes it when
    the backward pass is over, and the parameters end up with .grad = kernels' share + whatever autograd accumulated
    itself (smooth steps), with torch's accumulate / zero_grad semantics.  The stand-in node below does to the bucket
    exactly what train._RenderNode does (node_forward / node_backward_begin / write the views / node_done)."""
    from mpsnerf_b200.engine import DENSE_FP32_ORDER
    from mpsnerf_b200.train import DenseBucket

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(0)
            self._names = {}
            for i, k in enumerate(DENSE_FP32_ORDER):
                p = torch.nn.Parameter(torch.randn(3 + i % 5, 2 + i % 3, generator=g))
                self.register_parameter("p%d" % i, p)
                self._names[k] = p
            self.encoder_2d = torch.nn.Linear(4, 3)

        def named_parameters(self, *a, **k):
            return list(self._names.items()) + [("encoder_2d." + n, p) for n, p in self.encoder_2d.named_parameters()]

    net = Tiny()
    b = DenseBucket(net)

    class Node(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, share):
            b.node_forward()
            ctx.share = share
            return x * 1.0

        @staticmethod
        def backward(ctx, g):
            b.node_backward_begin()
            for v in b.views:                      # what the backward kernels do
                v += ctx.share
            b.node_done()
            return g, None

    def step(smooth):
        x = net.encoder_2d(torch.ones(2, 4))       # the trunk under autograd, upstream of the nodes
        loss = Node.apply(x, 1.0).sum() + Node.apply(x, 10.0).sum()          # two subjects
        if smooth:                                 # a term autograd differentiates with respect to dense parameters
            loss = loss + sum((p ** 2).sum() for p in b.params[:3])
        loss.backward()

    step(smooth=False)
    assert not b.open and not b.explicit
    for p, v in zip(b.params, b.views):
        assert p.grad is not None and p.grad.data_ptr() == v.data_ptr() and torch.allclose(p.grad, torch.full_like(p, 11.0))
    assert net.encoder_2d.weight.grad is not None
    # no zero_grad in between: gradients accumulate, as with any torch parameter
    step(smooth=True)
    for i, p in enumerate(b.params):
        want = torch.full_like(p, 22.0) + (2 * p.detach() if i < 3 else 0)
        assert torch.allclose(p.grad, want), i
    # zero_grad (either flavour) starts from zero again
    for set_to_none in (True, False):
        torch.optim.SGD(list(net.parameters()), lr=0.0).zero_grad(set_to_none=set_to_none)
        step(smooth=True)
        for i, p in enumerate(b.params):
            want = torch.full_like(p, 11.0) + (2 * p.detach() if i < 3 else 0)
            assert torch.allclose(p.grad, want), (set_to_none, i)
    # an explicit step (TrainStep) is left alone by the automatic machinery
    for p in net.parameters():
        p.grad = None
    b.begin_step(1)
    step(smooth=True)
    assert b.open and b.explicit and all(p.grad is None for p in b.params[3:])      # nothing finalised behind its back
    b.absorb_autograd()
    b.finish()
    for i, p in enumerate(b.params):
        want = torch.full_like(p, 11.0) + (2 * p.detach() if i < 3 else 0)
        assert torch.allclose(p.grad, want), i
