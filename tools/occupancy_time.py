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
