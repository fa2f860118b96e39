"""Density-grid query (BASELINE.json configs[4] / extract_thuman_mesh.py:114-125): an R^3 grid of points over
the body's box through ``network_fn`` -- target-pose space (mask + deformation + network) and canonical space
(``set_extract_mesh``: every point evaluated) -- timed on one GPU.  With torchrun the grid is split in z-slabs
over the ranks (no collective on the data path).

    python tools/grid_query.py --res 256 [--chunk 1048576]
prints one JSON line per mode: points/s, ms, active points.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=1 << 22)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    scene, net, args = bench.build_scene_and_net("bf16", None)
    from mpsnerf_b200 import run_nerf_batch as R
    from mpsnerf_b200.parallel import ray_block, max_over_ranks
    handle = R.NetworkHandle(net).to(dev).eval()
    dd = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else dd(v) if isinstance(v, dict) else v) for k, v in d.items()}
    sp, tp = dd(scene.sp_input), dd(scene.tp_input)
    z0, z1 = ray_block(a.res, rank, world)                      # this rank's z-slab
    for mode in ("target_pose", "canonical"):
        v = (scene.tp_input["vertices"] if mode == "target_pose" else scene.sp_input["t_vertices"])[0].numpy()
        lo, hi = v.min(0) - 0.05, v.max(0) + 0.05
        ax = [torch.linspace(float(lo[k]), float(hi[k]), a.res, device=dev) for k in range(3)]
        gz, gy, gx = torch.meshgrid(ax[2][z0:z1], ax[1], ax[0], indexing="ij")
        pts = torch.stack([gx, gy, gz], -1).reshape(1, -1, 3).contiguous()
        net.set_extract_mesh(mode == "canonical")
        times, n_act = [], 0
        for rep in range(a.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dens = []
            for s in range(0, pts.shape[1], a.chunk):
                out = handle(sp, tp, pts[:, s:s + a.chunk], None)
                dens.append(out[0, :, 3].clone())
                n_act += net.engine().last_active if rep == a.reps else 0
            e1.record()
            e1.synchronize()
            if rep:
                times.append(e0.elapsed_time(e1))
        ms = max_over_ranks(float(np.median(times)), dev)
        if rank == 0:
            print(json.dumps({"workload": f"density_grid_{a.res}^3_{mode}", "n_gpus": world, "points": a.res ** 3, "ms": ms,
                              "points_per_s": a.res ** 3 / (ms * 1e-3), "active_points_rank0": int(n_act), "chunk": a.chunk,
                              "split": f"z-slabs over {world} rank(s)"}))
    net.set_extract_mesh(False)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
