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
    if os.environ.get("MPSNERF_TC_PROF"):
        import ctypes
        from mpsnerf_b200 import _lib
        buf = (ctypes.c_ulonglong * 32)()
        _lib.check(_lib.load().mpsnerf_debug_read_prof(buf), "read_prof")
        names = ["epi_wait_mma", "epi_work", "tile_load", "mma_wait_A", "mma_wait_W", "mma_total", "prod_wait_slot", "tiles"]
        for k, kn in enumerate(("T", "M")):
            v = [buf[16 * k + i] for i in range(16)]
            tiles = max(v[7], 1)
            print(kn, "tiles(cta0-thread0 sum)", v[7], {n: round(x / tiles) for n, x in zip(names[:7], v[:7])}, "cycles/tile")
            if k == 0:
                sn = ["publish", "bar1", "dots", "bar2", "softmax_o", "LN2", "GELU", "LN1"]
                print("  T sections", {n: round(x / tiles) for n, x in zip(sn, v[8:16])})


if __name__ == "__main__":
    main()
