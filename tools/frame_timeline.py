"""Kernel timeline of one full render() step (torch.profiler / CUPTI): start, end, duration, stream, name."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    workload = sys.argv[1] if len(sys.argv) > 1 else "thuman"
    scene, net, args = bench.build_scene_and_net("bf16", None, workload)
    handle = R.NetworkHandle(net).cuda().eval()
    cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else cuda(v) if isinstance(v, dict) else v) for k, v in d.items()}
    sp, tp = cuda(scene.sp_input), cuda(scene.tp_input)
    rays, near, far = synthetic.rays_tensor(scene, None, device="cuda")
    kw = dict(network_fn=handle, N_samples=64, perturb=False, use_viewdirs=True)
    fn = lambda: R.render(rays=rays, near=near, far=far, sp_input=sp, tp_input=tp, **kw)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        flush.fill_(1)
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if evs:
        t0 = evs[0].time_range.start
        for e in evs:
            st = getattr(e, "device_index", -1)
            print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:100]}")


if __name__ == "__main__":
    main()
