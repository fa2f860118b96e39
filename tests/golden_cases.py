"""Table of the golden cases (tests/golden/<name>.npz), shared by oracle/make_golden.py -- which runs the
UNMODIFIED reference on them -- and by the tests, which rebuild the same seeded scenes and weights.

A case = one or more subjects (``scenes``: kwargs of synthetic.make_scene; more than one = a DataLoader batch
B > 1), ray count per subject, samples per ray, the seeded weights (``alpha_gain`` / ``alpha_bias`` of
synthetic.seeded_state_dict, seeded by ``weights_seed`` or the first scene's seed) and the render variant
(``perturb``: stratified sampling with supplied uniforms; ``white_bkgd``; ``occupancy`` = --occupancy 1,
run_nerf_batch.py:383-386).  ``smpl`` = "single": the three SMPL pickles hold the first scene's model (round-1
cases); "by_gender": male / female / neutral pickles differ (m/f/n template files) and the subject's ``gender``
selects one (lib/skinnning_batch.py:335-340).
"""
CASES = {
    # ---- round 1
    "plain": dict(scenes=[dict(kind="thuman", seed=0)], n_rays=512, S=64, alpha_gain=1.0),
    "stress": dict(scenes=[dict(kind="thuman", seed=1, novel_pose=True)], n_rays=384, S=64, alpha_gain=1000.0, perturb=True),
    "h36m": dict(scenes=[dict(kind="h36m", seed=2, H=500, W=500, novel_pose=True, t_vertices_from="file")], n_rays=256, S=128,
                 alpha_gain=300.0),
    # ---- round 2
    # opaque surfaces: a +40 bias on the density head saturates the transmittance within a few samples (acc > 0.99),
    # stratified jitter on, white background composited in
    "opaque": dict(scenes=[dict(kind="thuman", seed=3, novel_pose=True)], n_rays=384, S=64, alpha_gain=300.0, alpha_bias=40.0,
                   perturb=True, white_bkgd=True),
    # --occupancy 1 compositing variant.  There alpha = wide_sigmoid(raw_3) directly, without the x dist (~0.02) of the
    # density form: a density head of gain 300 would be ~40x more sensitive than anything in density mode (a bf16-level
    # error of 0.1 in raw_3 moves alpha by 0.025), so this case uses gain 60 -- raw_3 in -2 .. +3, still saturating
    "occupancy": dict(scenes=[dict(kind="thuman", seed=4)], n_rays=256, S=64, alpha_gain=60.0, alpha_bias=0.0, occupancy=1),
    # H36M at its real size: 1000 x 1000 input views, latent 250 x 250
    "h36m_full": dict(scenes=[dict(kind="h36m", seed=6, novel_pose=True)], n_rays=256, S=64, alpha_gain=300.0),
    # B = 2 subjects in one call, genders 1 (male tables) and 0 (female tables)
    "batch2": dict(scenes=[dict(kind="thuman", seed=7, gender="m", smpl_seed=7),
                           dict(kind="thuman", seed=8, gender="f", novel_pose=True, smpl_seed=7)],
                   n_rays=192, S=48, alpha_gain=300.0, alpha_bias=8.0, smpl="by_gender"),
}


def smpl_models(spec):
    from mpsnerf_b200 import synthetic
    first = spec["scenes"][0]
    if spec.get("smpl", "single") == "by_gender":
        return synthetic.smpl_by_gender(first.get("smpl_seed", first["seed"]))
    m = synthetic.make_smpl(first.get("gender", "n"), first["seed"])
    return {"male": m, "female": m, "neutral": m}


def build_case(spec):
    """-> (scenes, state_dict, sp_input, tp_input) with the DataLoader batch dim B = len(scenes)."""
    from mpsnerf_b200 import synthetic
    scenes = [synthetic.make_scene(**kw) for kw in spec["scenes"]]
    sd = synthetic.seeded_state_dict(spec.get("weights_seed", scenes[0].seed), float(spec["alpha_gain"]), spec.get("alpha_bias"))
    if len(scenes) == 1:
        sp, tp = scenes[0].sp_input, scenes[0].tp_input
    else:
        sp, tp = synthetic.batch_scenes(scenes)
    return scenes, sd, sp, tp


# ---- training step (BASELINE config 4): one subject, rays with stratified jitter, a seeded target image
TRAIN_CASE = dict(scene=dict(kind="thuman", seed=11, H=128, W=128, novel_pose=True), n_rays=320, S=48, alpha_gain=300.0,
                  alpha_bias=2.0)


# The smooth-loss step is checked on the same scene and rays with a density head of moderate gain.  Its terms are
# functions of NORMALISED gradients of the occupancy sigmoid(alpha): at the bench weight scale (gain 300) most active
# points sit deep in the sigmoid's saturation, their raw gradient is ~1e-30, and the derivative of the normalisation
# (~1 / |g|) amplifies fp32 rounding into 5-10 % run-to-run noise on single gradient entries (the order of the active
# points differs between runs: K1 compacts with atomics).  Gain 4 keeps alpha within a few units of 0: well conditioned.
SMOOTH_CASE = dict(TRAIN_CASE, alpha_gain=4.0, alpha_bias=0.0)


def build_smooth_case():
    return build_train_case(SMOOTH_CASE)


def build_train_case(spec=None):
    """-> (scene, state_dict, ray ids, S, u (n_rays, S), target rgb (n_rays, 3), bkgd_msk (n_rays,))."""
    import numpy as np
    from mpsnerf_b200 import synthetic
    spec = spec or TRAIN_CASE
    scene = synthetic.make_scene(**spec["scene"])
    sd = synthetic.seeded_state_dict(scene.seed, spec["alpha_gain"], spec["alpha_bias"])
    ids = synthetic.inbox_ray_subset(scene, spec["n_rays"])
    rng = np.random.RandomState(1234)
    u = rng.uniform(0, 1, (spec["n_rays"], spec["S"])).astype(np.float32)
    target = rng.uniform(0, 1, (spec["n_rays"], 3)).astype(np.float32)
    msk = (rng.uniform(0, 1, spec["n_rays"]) > 0.5).astype(np.float32)
    return scene, sd, ids, spec["S"], u, target, msk
