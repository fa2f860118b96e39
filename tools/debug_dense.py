"""Compare the T kernel's output tokens (tok0/tok1 in the dense workspace) and raw against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case
from test_gpu_parity import make_net, _cuda_dict, _engine_consts
from mpsnerf_b200.run_nerf_batch import _select
from oracle import oracle as O

scene, sd, g = load_case("plain")
net = make_net(scene, sd, "bf16")
ids, S = g["ray_ids"], int(g["S"])
sp, tp = _select(_cuda_dict(scene.sp_input), 0), _select(_cuda_dict(scene.tp_input), 0)
ctx = net.frame_context(sp, tp)
eng = net.engine(); eng.debug = {}
rays8 = torch.from_numpy(np.concatenate([scene.rays_o[ids], scene.rays_d[ids], scene.near[ids, None], scene.far[ids, None]], 1)).cuda()
res = eng.run(ctx, rays8=rays8, S=S, t_vals=torch.linspace(0, 1, S, device="cuda"))
torch.cuda.synchronize()
n = res["n_active"]
ws = eng._ws["dense"]
tokb = ((n * 160 * 2 + 255) // 256) * 256
tok0 = ws[:n * 320].view(torch.bfloat16).reshape(n, 160).float().cpu().numpy()
tok1 = ws[tokb:tokb + n * 320].view(torch.bfloat16).reshape(n, 160).float().cpu().numpy()
dbg = {k: (torch.cat(v).cpu().numpy() if isinstance(v, list) else v.cpu().numpy()) for k, v in eng.debug.items()}
order = np.argsort(dbg["act_pid"], kind="stable")
smpl = O.smpl_tensors(scene.smpl)
sp_c, tp_c = O.squeeze_inputs(scene.sp_input, scene.tp_input)
c = _engine_consts(ctx, O.frame_constants(smpl, sp_c, tp_c))
z = O.sample_z(scene.near[ids], scene.far[ids], S, None)
pts = O.sample_points(scene.rays_o[ids], scene.rays_d[ids], z).reshape(-1, 3)
raw17, st = O.forward_points(smpl, sd, sp_c, tp_c, pts, bf16=True, latent=ctx.latent.permute(0, 3, 1, 2).cpu(), consts=c, return_stages=True)
to = st["tok_out"]
for name, mine, ref in (("tok0", tok0[order][:, :155], to[:, 0]), ("tok1", tok1[order][:, :155], to[:, 1])):
    d = np.abs(mine - ref)
    print(name, "max abs diff", d.max(), "mean", d.mean(), "ref absmax", np.abs(ref).max(), "nan", np.isnan(mine).sum(),
          "worst cols", np.argsort(-d.max(0))[:8], "worst rows", np.argsort(-d.max(1))[:8])
print("pad cols tok0", np.abs(tok0[:, 155:]).max())
raw = res["raw"].cpu().numpy(); m = st["mask"]
d = np.abs(raw[m] - raw17[m, :4]); print("raw max diff", d.max(0), "rows bad", (d.max(1) > 0.1).sum(), "of", m.sum())
