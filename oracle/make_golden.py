"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py [case ...]     # default: every case of tests/golden_cases.py

Each case = one or more seeded synthetic scenes (mpsnerf_b200.synthetic, regenerated from
their seeds at test time), the seeded live weights (synthetic.seeded_state_dict) loaded
into the reference's own SKinningBatch, and the outputs of the reference's own
``run_nerf_batch.render`` plus per-stage tensors captured by wrapping (not editing)
reference functions.  The case table lives in tests/golden_cases.py so that the tests
rebuild exactly the same scenes.
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mpsnerf_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402
from golden_cases import CASES, build_case  # noqa: E402

TOKENS_KEEP = 192


def run_case(R, name, spec):
    from model_selection import return_model
    import lib.skinnning_batch as SB

    scenes, sd, sp_input, tp_input = build_case(spec)
    n_rays, S, perturb = spec["n_rays"], spec["S"], spec.get("perturb", False)
    B = len(scenes)
    torch.manual_seed(0)
    R.global_args.N_samples = S
    R.global_args.occupancy = int(spec.get("occupancy", 0))
    net = return_model(R.global_args)
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    net.eval()
    wrapped = ref_shims.ScatterLike(net)

    ids = [synthetic.inbox_ray_subset(sc, n_rays) for sc in scenes]
    parts = [synthetic.rays_tensor(sc, i) for sc, i in zip(scenes, ids)]
    rays, near, far = (torch.cat([p[k] for p in parts], 0) for k in range(3))
    u = None
    if perturb:
        u = torch.from_numpy(np.random.RandomState(99).uniform(0, 1, (B, n_rays, S)).astype(np.float32))

    rec = {"proj_in": [], "uv": [], "tok_in": [], "tok_out": []}
    orig_proj = SB.SKinningBatch.projection

    def proj(self, q, Rm, T, K):
        out = orig_proj(self, q, Rm, T, K)
        rec["proj_in"].append(q.detach().clone())
        rec["uv"].append(out.detach().clone())
        return out

    SB.SKinningBatch.projection = proj

    def tok_hook(m, i, o):
        rec["tok_in"].append(i[0].detach().clone())
        rec["tok_out"].append(o.detach().clone())

    hook = net.transformer.register_forward_hook(tok_hook)
    orig_rand = torch.rand
    if perturb:
        torch.rand = lambda *a, **k: u.clone()
    ref_shims.KNN_LOG.clear()
    try:
        with torch.no_grad():
            rgb, disp, acc, extras = R.render(
                chunk=n_rays, rays=rays, near=near, far=far, sp_input=sp_input, tp_input=tp_input,
                network_query_fn=lambda i, v, f, sp_input=None, tp_input=None: R.run_network(i, v, f, sp_input=sp_input, tp_input=tp_input),
                perturb=1.0 if perturb else False, N_samples=S, network_fn=wrapped, use_viewdirs=True, N_importance=0,
                white_bkgd=bool(spec.get("white_bkgd", False)))
    finally:
        torch.rand = orig_rand
        SB.SKinningBatch.projection = orig_proj
        hook.remove()
        R.global_args.occupancy = 0
    assert len(ref_shims.KNN_LOG) == 3 * B, len(ref_shims.KNN_LOG)
    out = {
        "ray_ids": np.stack(ids).astype(np.int64) if B > 1 else ids[0].astype(np.int64), "S": np.int64(S),
        "alpha_gain": np.float32(spec["alpha_gain"]),
        "rgb_map": rgb.numpy(), "disp_map": disp.numpy(), "acc_map": acc.numpy(),
        "raw": extras["raw"].numpy(), "pts_mask": extras["pts_mask"].numpy().astype(np.uint8),
        "smpl_query_pts": extras["smpl_query_pts"].numpy(), "smpl_src_pts": extras["smpl_src_pts"].numpy(),
    }
    for b in range(B):
        (_, d2_all, _), (q_act, _, idx2), (xc, _, idx3) = ref_shims.KNN_LOG[3 * b:3 * b + 3]
        sfx = "" if B == 1 else f"_{b}"
        out.update({
            "d2_all" + sfx: d2_all.numpy(), "q_active" + sfx: q_act.numpy(), "idx2" + sfx: idx2.numpy().astype(np.int32),
            "xc" + sfx: xc.numpy(), "idx3" + sfx: idx3.numpy().astype(np.int32),
            "xw" + sfx: rec["proj_in"][b].numpy(), "uv" + sfx: rec["uv"][b].numpy(),
            "tok_in" + sfx: rec["tok_in"][b][:TOKENS_KEEP].numpy(), "tok_out" + sfx: rec["tok_out"][b][:TOKENS_KEEP].numpy()})
    if B == 1:      # the layout of the round-1 goldens: no batch dim on the outputs
        for k in ("rgb_map", "disp_map", "acc_map", "raw", "pts_mask", "smpl_query_pts", "smpl_src_pts"):
            out[k] = out[k][0]
    if u is not None:
        out["u"] = u[0].numpy() if B == 1 else u.numpy()
    m = out["pts_mask"][..., 0] == 1
    n_act = int(m.sum())
    print(f"[{name}] B {B} rays {n_rays} S {S} active {n_act} ({100.0 * n_act / (B * n_rays * S):.1f}%) "
          f"raw alpha range [{out['raw'][..., 3][m].min():.3f}, {out['raw'][..., 3][m].max():.3f}] "
          f"acc max {out['acc_map'].max():.4f}, rays with acc > 0.99: {int((out['acc_map'] > 0.99).sum())}")
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    work = tempfile.mkdtemp(prefix="mpsnerf_ref_")
    ref_shims.install(work, synthetic.make_smpl("n", 0))
    R = ref_shims.load_reference(n_samples=64)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    files = {"male": "basicmodel_m_lbs_10_207_0_v1.0.0.pkl", "female": "basicmodel_f_lbs_10_207_0_v1.0.0.pkl",
             "neutral": "SMPL_NEUTRAL.pkl"}
    for name in names:
        spec = CASES[name]
        # the reference reads its three SMPL pickles from ./assets at construction: rewrite them for this case
        models = spec_smpl_models(spec)
        for g, f in files.items():
            with open(os.path.join(work, "assets", f), "wb") as fh:
                pickle.dump(models[g], fh)
        out = run_case(R, name, spec)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def spec_smpl_models(spec):
    from golden_cases import smpl_models
    return smpl_models(spec)


if __name__ == "__main__":
    main()
