"""Generate tests/golden/smooth_grads.npz: the smooth-loss terms of one training step (normals of the occupancy field by
double backward, run_nerf_batch.py:60-79 + lib/skinnning_batch.py:408-412, 496-504) and the gradients of
other_loss[0][0] computed by the UNMODIFIED reference's own autograd on CPU.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_smooth.py

Scene and rays of oracle/make_golden_train.py with a density head of moderate gain (tests/golden_cases.py SMOOTH_CASE:
the terms are ill-conditioned where the occupancy saturates), global_step = 0 and
smooth_interval = 4 (an interval step); the perturbation delta_x ~ N(0, 0.01) that the reference draws from the global
RNG is drawn here from a seeded generator and stored, and handed to the reference through its module-level
``perturb_distri``.  Stores: delta_x, other_loss (4), the two normal fields of the unperturbed pass (columns 17:23 of
the network output), and per live parameter the gradient's L2 norm and 512 evenly strided entries.

Runs with ONE CPU thread: the reference's compute_normal (lib/skinnning_batch.py:29-41) scatters with
``norm[faces[:, s]] += n``, which keeps one contribution per repeated index and is a race between threads -- two
multi-threaded runs of the reference disagree on ~4 500 of the 6 890 vertex normals.  One thread executes it in index
order (the last face wins), which is the definition mps-nerf_b200/smooth.py:vertex_normals reproduces on any device.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mpsnerf_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402
from oracle import train_oracle as TO  # noqa: E402
from golden_cases import build_smooth_case  # noqa: E402

SMOOTH_KEYS = [k for k in TO.dense_keys() if not k.startswith(("feature_linear", "views_linear", "rgb_linear"))] + TO.TRUNK_KEYS


class _FixedDelta:
    def __init__(self, delta):
        self.delta = delta

    def sample(self, shape):
        assert tuple(shape) == tuple(self.delta.shape[:-1]), (shape, self.delta.shape)
        return self.delta.clone()


def main():
    torch.set_num_threads(1)
    scene, sd, ids, S, u, target, msk = build_smooth_case()
    work = tempfile.mkdtemp(prefix="mpsnerf_ref_")
    ref_shims.install(work, scene.smpl)
    R = ref_shims.load_reference(n_samples=S)
    from model_selection import return_model
    R.global_args.N_samples = S
    R.global_args.smooth_loss = 1
    torch.manual_seed(0)
    net = return_model(R.global_args)
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert net.smooth_loss
    net.train()
    wrapped = ref_shims.ScatterLike(net)
    rays, near, far = synthetic.rays_tensor(scene, ids)
    delta = (torch.randn(1, len(ids) * S, 3, generator=torch.Generator().manual_seed(7)) * 0.01).float()
    R.perturb_distri = _FixedDelta(delta)
    sp = dict(scene.sp_input)
    sp["global_step"] = torch.zeros(1, dtype=torch.long)
    sp["smooth_interval"] = torch.full((1,), 4, dtype=torch.long)
    seen = {}

    def query(i, v, f, sp_input=None, tp_input=None):
        out, other = R.run_network(i, v, f, sp_input=sp_input, tp_input=tp_input)
        seen["out"] = out
        return out, other

    orig_rand = torch.rand
    torch.rand = lambda *a, **k: torch.from_numpy(u)[None].clone()
    try:
        rgb, disp, acc, extras = R.render(chunk=len(ids), rays=rays, near=near, far=far, sp_input=sp, tp_input=scene.tp_input,
                                          network_query_fn=query, perturb=1.0, N_samples=S, network_fn=wrapped,
                                          use_viewdirs=True, N_importance=0)
    finally:
        torch.rand = orig_rand
    other = extras["other_loss"]
    full = seen["out"].reshape(-1, seen["out"].shape[-1])
    assert full.shape[-1] == 23, full.shape
    other[0][0].backward()
    named = dict(net.named_parameters())
    out = {"delta": delta[0].numpy(), "other_loss": other.detach().numpy().reshape(4).astype(np.float64),
           "occ_normal": full[:, 17:20].detach().numpy(), "smpl_normal": full[:, 20:23].detach().numpy(),
           "n_active": np.int64(int(extras["pts_mask"].sum()))}
    for k in SMOOTH_KEYS:
        g = named[k].grad
        assert g is not None, k
        flat = g.reshape(-1).numpy()
        sel = np.linspace(0, len(flat) - 1, min(512, len(flat))).astype(np.int64)
        out["norm/" + k] = np.float64(np.linalg.norm(flat.astype(np.float64)))
        out["idx/" + k] = sel
        out["val/" + k] = flat[sel]
    extra = [k for k, p in named.items() if p.grad is not None and k not in SMOOTH_KEYS and float(p.grad.abs().max()) > 0]
    print("other_loss", out["other_loss"], "active", int(out["n_active"]), "params with non-zero grad outside the list:", extra)
    path = os.path.join(ROOT, "tests", "golden", "smooth_grads.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
