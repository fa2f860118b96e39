import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from mpsnerf_b200 import extract_thuman_mesh as X, synthetic
sc = synthetic.make_scene("thuman", seed=0, H=64, W=64, novel_pose=True)
q, _, _, _ = X.grid_points(False, 256)
flat = torch.from_numpy(q.reshape(-1, 3)).cuda()
raw = torch.randn(flat.shape[0], 4, device="cuda")
verts = torch.as_tensor(np.asarray(sc.tp_input["vertices"])).reshape(-1, 3).float()
nrm = torch.randn(verts.shape[0], 3)
for _ in range(2):
    X.occupancy_post(flat, raw, verts, None, normals=nrm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    occ = X.occupancy_post(flat, raw, verts, None, normals=nrm)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("occupancy_fix 256^3 x 6890: %.2f ms  = %.2f T pair-evals/s" % (ms, flat.shape[0] * verts.shape[0] / ms / 1e9))

# the whole per-frame block of the reference script (grid -> network -> post-step), 256^3 points
import json, time
from mpsnerf_b200 import run_nerf_batch as R
sd = synthetic.seeded_state_dict(0, 300.0)
import bench
scene, net, args = bench.build_scene_and_net("bf16", 1)
handle = R.NetworkHandle(net).cuda().eval()
cuda = lambda d: {k: (v.cuda() if torch.is_tensor(v) else cuda(v) if isinstance(v, dict) else v) for k, v in d.items()}
sp, tp = cuda(scene.sp_input), cuda(scene.tp_input)
res = {}
for can in (False, True):
    X.estimate_occupancy(handle, sp, tp, scene.smpl["f"], can_flag=can, n=64)       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    occ_v, START, SIZE, RANGE = X.estimate_occupancy(handle, sp, tp, scene.smpl["f"], can_flag=can, n=256)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res["canonical" if can else "target_pose"] = {"points": int(occ_v.size), "wall_ms_incl_d2h": round(dt * 1e3, 1),
                                                  "frac_occupied_gt_30": float((occ_v > 30).mean())}
res["occupancy_fix_ms_256cubed_x_6890"] = round(ms, 2)
print(json.dumps(res))
