"""Print the M-kernel event trace of CTA 0's second tile (run with MPSNERF_TC_PROF=2)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MPSNERF_TC_PROF", "2")
import bench
from mpsnerf_b200 import synthetic, _lib
from mpsnerf_b200 import run_nerf_batch as R
scene, net, args = bench.build_scene_and_net("bf16", 1)
handle = R.NetworkHandle(net).cuda().eval()
cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else cuda(v) if isinstance(v, dict) else v) for k, v in d.items()}
sp, tp = cuda(scene.sp_input), cuda(scene.tp_input)
rays, near, far = synthetic.rays_tensor(scene, None, device="cuda")
for _ in range(2):
    R.render(rays=rays, near=near, far=far, sp_input=sp, tp_input=tp, network_fn=handle, N_samples=64, perturb=False, use_viewdirs=True)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 512)()
_lib.check(_lib.load().mpsnerf_debug_read_trace(buf), "read_trace")
names = {1: "mma wait_a0", 2: "mma wait_a1", 3: "mma got_a0", 4: "mma got_a1", 5: "mma issued h0 (commit d0)", 6: "mma issued h1K0 (commit k)",
         7: "mma issued h1K1 (commit d1)", 10: "epi wait_d0", 11: "epi wait_d1", 12: "epi got_d0", 13: "epi got_d1", 14: "epi wait_k", 15: "epi got_k",
         16: "epi arrive a0", 17: "epi arrive a1", 18: "epi ld done", 19: "epi st done", 20: "mma slot begin", 21: "mma slot acquired", 22: "mma slot issued", 30: "mma tile start", 31: "mma step complete", 32: "PRELUDE issue cycles (value)", 33: "PRELUDE total cycles (value)", 23: "epi vouch begin", 24: "epi vouch done"}
ev = []
for base, who in ((0, "MMA"), (256, "EPI")):
    for i in range(256):
        v = buf[base + i]
        if v == 0:
            continue
        if 32 <= (v >> 48) < 48:
            t = (v >> 48) - 32
            print("PRELUDE cfg", t // 2, "issue" if t % 2 == 0 else "total", v & 0xffffffffffff, "cycles per burst of 8 MMAs")
            continue
        ev.append((v & 0xffffffffffff, v >> 48, who))
ev.sort()
t0 = ev[0][0] if ev else 0
if os.environ.get("ONLY_MMA"):
    ev = [e for e in ev if e[2] == "MMA"]
for t, tag, who in ev[:int(os.environ.get("N_EV", "140"))]:
    print(f"{t - t0:8d}  {'' if who == 'MMA' else '                              '}{names.get(tag, tag)}")
