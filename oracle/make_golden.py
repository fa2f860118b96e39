"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/{plain,stress,h36m}.npz

Each case = a seeded synthetic scene (mpsnerf_b200.synthetic, regenerated from its
seed at test time), the seeded live weights (synthetic.seeded_state_dict) loaded
into the reference's own SKinningBatch, and the outputs of the reference's own
``run_nerf_batch.render`` plus per-stage tensors captured by wrapping (not editing)
reference functions.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mpsnerf_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

CASES = {
    # name: scene kwargs, n_rays, S, alpha_gain, perturb
    "plain": (dict(kind="thuman", seed=0), 512, 64, 1.0, False),
    "stress": (dict(kind="thuman", seed=1, novel_pose=True), 384, 64, 1000.0, True),
    "h36m": (dict(kind="h36m", seed=2, H=500, W=500, novel_pose=True, t_vertices_from="file"), 256, 128, 300.0, False),
}
TOKENS_KEEP = 192


def run_case(R, name, scene_kw, n_rays, S, alpha_gain, perturb):
    from model_selection import return_model
    import lib.skinnning_batch as SB

    scene = synthetic.make_scene(**scene_kw)
    torch.manual_seed(0)
    R.global_args.N_samples = S
    net = return_model(R.global_args)
    missing = net.load_state_dict(synthetic.seeded_state_dict(scene.seed, alpha_gain), strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    net.eval()
    wrapped = ref_shims.ScatterLike(net)

    ids = synthetic.inbox_ray_subset(scene, n_rays)
    rays, near, far = synthetic.rays_tensor(scene, ids)
    u = None
    if perturb:
        u = torch.from_numpy(np.random.RandomState(99).uniform(0, 1, (1, n_rays, S)).astype(np.float32))

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
                chunk=n_rays, rays=rays, near=near, far=far, sp_input=scene.sp_input, tp_input=scene.tp_input,
                network_query_fn=lambda i, v, f, sp_input=None, tp_input=None: R.run_network(i, v, f, sp_input=sp_input, tp_input=tp_input),
                perturb=1.0 if perturb else False, N_samples=S, network_fn=wrapped, use_viewdirs=True, N_importance=0)
    finally:
        torch.rand = orig_rand
        SB.SKinningBatch.projection = orig_proj
        hook.remove()
    assert len(ref_shims.KNN_LOG) == 3, len(ref_shims.KNN_LOG)
    (_, d2_all, _), (q_act, _, idx2), (xc, _, idx3) = ref_shims.KNN_LOG
    out = {
        "ray_ids": ids.astype(np.int64), "S": np.int64(S), "alpha_gain": np.float32(alpha_gain),
        "rgb_map": rgb[0].numpy(), "disp_map": disp[0].numpy(), "acc_map": acc[0].numpy(),
        "raw": extras["raw"][0].numpy(), "pts_mask": extras["pts_mask"][0].numpy().astype(np.uint8),
        "smpl_query_pts": extras["smpl_query_pts"][0].numpy(), "smpl_src_pts": extras["smpl_src_pts"][0].numpy(),
        "d2_all": d2_all.numpy(), "q_active": q_act.numpy(), "idx2": idx2.numpy().astype(np.int32),
        "xc": xc.numpy(), "idx3": idx3.numpy().astype(np.int32),
        "xw": rec["proj_in"][0].numpy(), "uv": rec["uv"][0].numpy(),
        "tok_in": rec["tok_in"][0][:TOKENS_KEEP].numpy(), "tok_out": rec["tok_out"][0][:TOKENS_KEEP].numpy(),
    }
    if u is not None:
        out["u"] = u[0].numpy()
    n_act = int(out["pts_mask"].sum())
    print(f"[{name}] rays {n_rays} S {S} active {n_act} ({100.0 * n_act / (n_rays * S):.1f}%) "
          f"raw alpha range [{out['raw'][..., 3][out['pts_mask'][..., 0] == 1].min():.3f}, "
          f"{out['raw'][..., 3][out['pts_mask'][..., 0] == 1].max():.3f}] acc max {out['acc_map'].max():.3f}")
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    smpl_by_seed = {}
    work = tempfile.mkdtemp(prefix="mpsnerf_ref_")
    first = CASES[names[0]][0]
    ref_shims.install(work, synthetic.make_smpl(first.get("gender", "n"), first["seed"]))
    R = ref_shims.load_reference(n_samples=64)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in names:
        scene_kw, n_rays, S, gain, perturb = CASES[name]
        # the reference reads its SMPL from ./assets at construction: rewrite for this case's seed
        import pickle
        smpl = synthetic.make_smpl(scene_kw.get("gender", "n"), scene_kw["seed"])
        for f in os.listdir(os.path.join(work, "assets")):
            with open(os.path.join(work, "assets", f), "wb") as fh:
                pickle.dump(smpl, fh)
        out = run_case(R, name, scene_kw, n_rays, S, gain, perturb)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
