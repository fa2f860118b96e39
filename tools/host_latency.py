"""Host-side cost of one render() call vs its device time: 30 calls enqueued back to back without synchronising."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from mpsnerf_b200 import run_nerf_batch as R, synthetic
    scene, net, args = bench.build_scene_and_net("bf16", None, "thuman")
    handle = R.NetworkHandle(net).cuda().eval()
    cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else cuda(v) if isinstance(v, dict) else v) for k, v in d.items()}
    sp, tp = cuda(scene.sp_input), cuda(scene.tp_input)
    rays, near, far = synthetic.rays_tensor(scene, None, device="cuda")
    kw = dict(network_fn=handle, N_samples=64, perturb=False, use_viewdirs=True)
    fn = lambda: R.render(rays=rays, near=near, far=far, sp_input=sp, tp_input=tp, **kw)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    host = []
    for _ in range(n):
        h0 = time.perf_counter()
        fn()
        host.append(time.perf_counter() - h0)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host per call: median {1e3 * sorted(host)[n // 2]:.3f} ms (min {1e3 * min(host):.3f}, max {1e3 * max(host):.3f}); "
          f"enqueue of {n} calls took {1e3 * (t1 - t0):.1f} ms, device span {e0.elapsed_time(e1) / n:.3f} ms per call, "
          f"wall {1e3 * (t2 - t0) / n:.3f} ms per call")
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
