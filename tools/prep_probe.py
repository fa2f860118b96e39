"""Kernel timeline of one frame-prep replay (torch.profiler / CUPTI): which kernels overlap, where the gaps are."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from mpsnerf_b200 import run_nerf_batch as R
    scene, net, args = bench.build_scene_and_net("bf16", 1)
    handle = R.NetworkHandle(net).cuda().eval()
    cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else cuda(v) if isinstance(v, dict) else v) for k, v in d.items()}
    sp, tp = cuda(scene.sp_input), cuda(scene.tp_input)
    from mpsnerf_b200.run_nerf_batch import _select
    sp0, tp0 = _select(sp, 0), _select(tp, 0)
    eng = net.engine()
    smpl = net._smpl_for(sp0["gender"])
    fn = lambda: eng.prepare_frame(sp0, tp0, smpl)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(20):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    print("prep avg ms (back to back, warm L2):", ev[0].elapsed_time(ev[1]) / 20)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if evs:
        t0 = evs[0].time_range.start
        for e in evs:
            print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")


if __name__ == "__main__":
    main()
