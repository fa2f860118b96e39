"""Run a few full-frame render steps (same workload as bench.py) -- the command profiled by ncu."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from mpsnerf_b200 import synthetic
    from mpsnerf_b200 import run_nerf_batch as R
    scene, net, args = bench.build_scene_and_net(a.precision, 1)
    handle = R.NetworkHandle(net).cuda().eval()
    cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else cuda(v) if isinstance(v, dict) else v) for k, v in d.items()}
    sp, tp = cuda(scene.sp_input), cuda(scene.tp_input)
    rays, near, far = synthetic.rays_tensor(scene, None, device="cuda")
    for _ in range(a.steps):
        rgb, disp, acc, ex = R.render(rays=rays, near=near, far=far, sp_input=sp, tp_input=tp, network_fn=handle,
                                      N_samples=64, perturb=False, use_viewdirs=True)
    torch.cuda.synchronize()
    print("ok", float(acc.mean()))


if __name__ == "__main__":
    main()
