"""Run the UNMODIFIED reference's own render() on CPU for a seeded synthetic scene.  TEST / BENCH INFRASTRUCTURE ONLY.

Used by ``bench.py``'s CPU arm when a reference tree is present ($MPSNERF_REF, /root/reference or baseline/_ref --
the build container; the GPU box never has one, the tree cannot travel) so that the baseline is the reference
itself (``cpu_baseline.kind = "reference"``) and not the oracle port.  The import shims are those of
oracle/ref_shims.py; ``.cuda()`` is forced to the identity so that the CPU arm stays on
the host cores even on a box with a GPU.
"""
import os
import sys
import tempfile

import torch

_STATE = {}


def find_reference():
    """Path of a reference tree that can be executed here, or None."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.environ.get("MPSNERF_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if p and os.path.isfile(os.path.join(p, "run_nerf_batch.py")) and os.path.isfile(os.path.join(p, "lib", "skinnning_batch.py")):
            return p
    return None


def load(scene, sd, n_samples=64):
    """Import the reference (once per process) and build its network with the seeded weights."""
    ref = find_reference()
    if ref is None:
        raise FileNotFoundError("no reference tree")
    from oracle import ref_shims
    if "R" not in _STATE:
        ref_shims.REF = ref
        cwd = os.getcwd()
        work = tempfile.mkdtemp(prefix="mpsnerf_ref_")
        ref_shims.install(work, scene.smpl)          # chdir(work): the reference reads ./assets at construction
        torch.Tensor.cuda = lambda self, *a, **k: self
        argv = list(sys.argv)
        try:
            _STATE["R"] = ref_shims.load_reference(n_samples=n_samples)
        finally:
            sys.argv = argv
        _STATE["cwd"], _STATE["work"] = cwd, work
    R = _STATE["R"]
    os.chdir(_STATE["work"])
    try:
        from model_selection import return_model
        R.global_args.N_samples = n_samples
        net = return_model(R.global_args)
    finally:
        os.chdir(_STATE["cwd"])
    net.load_state_dict(sd, strict=False)
    net.eval()
    return R, ref_shims.ScatterLike(net)


def render(R, wrapped, scene, ids, S=64):
    """The reference's run_nerf_batch.render on rays ``ids`` of ``scene`` -> [rgb, disp, acc, extras] (CPU tensors)."""
    from mpsnerf_b200 import synthetic
    rays, near, far = synthetic.rays_tensor(scene, ids)
    with torch.no_grad():
        return R.render(chunk=len(ids), rays=rays, near=near, far=far, sp_input=scene.sp_input, tp_input=scene.tp_input,
                        network_query_fn=lambda i, v, f, sp_input=None, tp_input=None: R.run_network(i, v, f, sp_input=sp_input, tp_input=tp_input),
                        perturb=False, N_samples=S, network_fn=wrapped, use_viewdirs=True, N_importance=0)
