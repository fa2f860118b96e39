"""Generate tests/golden/train_grads.npz: gradients of one training step computed by the UNMODIFIED reference's own
autograd on CPU (training mode, smooth term off; run_nerf_batch.py:544-570).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_train.py

Stores, per live parameter, the gradient's L2 norm and 512 evenly strided entries (the full set is 1.3 M floats),
the loss and the rendered rgb / acc of the step.
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
from golden_cases import TRAIN_CASE, build_train_case  # noqa: E402


def main():
    scene, sd, ids, S, u, target, msk = build_train_case()
    work = tempfile.mkdtemp(prefix="mpsnerf_ref_")
    ref_shims.install(work, scene.smpl)
    R = ref_shims.load_reference(n_samples=S)
    from model_selection import return_model
    R.global_args.N_samples = S
    R.global_args.smooth_loss = 0
    torch.manual_seed(0)
    net = return_model(R.global_args)
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    net.train()
    wrapped = ref_shims.ScatterLike(net)
    rays, near, far = synthetic.rays_tensor(scene, ids)
    orig_rand = torch.rand
    torch.rand = lambda *a, **k: torch.from_numpy(u)[None].clone()
    try:
        rgb, disp, acc, extras = R.render(
            chunk=len(ids), rays=rays, near=near, far=far, sp_input=scene.sp_input, tp_input=scene.tp_input,
            network_query_fn=lambda i, v, f, sp_input=None, tp_input=None: R.run_network(i, v, f, sp_input=sp_input, tp_input=tp_input),
            perturb=1.0, N_samples=S, network_fn=wrapped, use_viewdirs=True, N_importance=0)
    finally:
        torch.rand = orig_rand
    loss = torch.mean((rgb - torch.from_numpy(target)[None]) ** 2) + torch.mean((torch.from_numpy(msk)[None] - acc) ** 2)
    loss.backward()
    named = dict(net.named_parameters())
    out = {"loss": np.float64(loss.item()), "rgb_map": rgb[0].detach().numpy(), "acc_map": acc[0].detach().numpy(),
           "n_active": np.int64(int(extras["pts_mask"].sum()))}
    names = TO.dense_keys() + TO.TRUNK_KEYS
    for k in names:
        g = named[k].grad
        assert g is not None, k
        flat = g.reshape(-1).numpy()
        sel = np.linspace(0, len(flat) - 1, min(512, len(flat))).astype(np.int64)
        out["norm/" + k] = np.float64(np.linalg.norm(flat.astype(np.float64)))
        out["idx/" + k] = sel
        out["val/" + k] = flat[sel]
    dead = [k for k, p in named.items() if p.grad is not None and k not in names]
    print("loss", loss.item(), "active", int(out["n_active"]), "params with grad outside the live list:", dead)
    path = os.path.join(ROOT, "tests", "golden", "train_grads.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
