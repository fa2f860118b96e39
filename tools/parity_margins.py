"""Print the bf16 path's error margins against every single-subject golden case (rgb max-abs, PSNR, raw max-abs / scale)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASES, SINGLE_CASES, load_case  # noqa: E402
from test_gpu_parity import _configure, _cuda_dict, make_net, psnr  # noqa: E402


def main():
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    for name in SINGLE_CASES:
        scene, sd, g = load_case(name)
        spec = CASES[name]
        net = R.NetworkHandle(make_net(scene, sd, "bf16"))
        ids, S = g["ray_ids"], int(g["S"])
        rays, near, far = synthetic.rays_tensor(scene, ids, device="cuda")
        kw = dict(network_fn=net, N_samples=S, perturb=1.0 if "u" in g else False, white_bkgd=bool(spec.get("white_bkgd", False)))
        if "u" in g:
            kw["perturb_u"] = torch.from_numpy(g["u"])[None].cuda()
        _configure("bf16", spec.get("occupancy", 0))
        rgb, disp, acc, ex = R.render(rays=rays, near=near, far=far, sp_input=_cuda_dict(scene.sp_input),
                                      tp_input=_cuda_dict(scene.tp_input), use_viewdirs=True, **kw)
        _configure()
        m = (ex["pts_mask"][0, ..., 0].cpu().numpy() > 0.5) & (g["pts_mask"][..., 0] > 0)
        ok = ((ex["pts_mask"][0, ..., 0].cpu().numpy() > 0.5) == (g["pts_mask"][..., 0] > 0)).all(1)
        d = np.abs(rgb[0].cpu().numpy()[ok] - g["rgb_map"][ok])
        scale = max(1.0, float(np.abs(g["raw"][g["pts_mask"][..., 0] > 0]).max()))
        dr = np.abs(ex["raw"][0].cpu().numpy()[m] - g["raw"][m]).max() / scale
        print(f"{name:10s} rgb max {d.max():.5f}  psnr {psnr(rgb[0].cpu().numpy()[ok], g['rgb_map'][ok]):.1f}  raw/scale {dr:.5f}")


if __name__ == "__main__":
    main()
