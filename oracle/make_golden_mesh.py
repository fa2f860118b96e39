"""Golden vectors for the mesh-extraction post-step: executes the reference's OWN source lines
(/root/reference/extract_thuman_mesh.py: normalize_v3 / compute_normal :19-41 and the per-frame block :125-163)
under the import shims, on a small seeded grid, and writes tests/golden/mesh_post.npz.

Run here (the GPU box has no /root/reference):  python -m oracle.make_golden_mesh
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402


def main():
    import tempfile
    from mpsnerf_b200 import synthetic
    sc = synthetic.make_scene("thuman", seed=5, H=64, W=64, novel_pose=True)
    ref_shims.install(tempfile.mkdtemp(prefix="mpsnerf_ref_"), sc.smpl)
    src = open(os.path.join(ref_shims.REF, "extract_thuman_mesh.py")).read().split("\n")
    helpers = "\n".join(src[18:41])                       # :19-41  normalize_v3, compute_normal
    assert helpers.lstrip().startswith("def normalize_v3") and "def compute_normal" in helpers
    i0 = next(i for i, l in enumerate(src) if "occupancy = shifted_softplus(raw[...,3])" in l)      # :125
    i1 = next(i for i, l in enumerate(src) if "mcubes.marching_cubes" in l)                          # :164
    assert (i0, i1) == (124, 163), (i0, i1)
    block = [src[i0].strip()] + src[i0 + 1:i1]             # the first line sits inside `with torch.no_grad():`
    indent = min(len(l) - len(l.lstrip()) for l in block[1:] if l.strip())
    body = "\n".join((l[indent:] if l.strip() else "") if k else l for k, l in enumerate(block))

    verts = np.asarray(sc.tp_input["vertices"]).reshape(-1, 3).astype(np.float32)
    faces = np.asarray(sc.smpl["f"]).astype(np.int32)
    rng = np.random.RandomState(11)
    lo, hi = verts.min(0) - 0.25, verts.max(0) + 0.25
    n = 22
    g = np.stack(np.meshgrid(*[np.linspace(lo[k], hi[k], n) for k in range(3)]), -1).astype(np.float32)
    sh = g.shape
    flat_np = g.reshape(-1, 3)
    # a third of the points are moved next to the surface so that both mask branches are exercised
    pick = rng.rand(len(flat_np)) < 0.35
    flat_np[pick] = verts[rng.randint(0, len(verts), pick.sum())] + rng.normal(0, 0.03, (pick.sum(), 3)).astype(np.float32)
    raw_np = rng.normal(0.0, 3.0, (len(flat_np), 4)).astype(np.float32)

    from pytorch3d.ops.knn import knn_points
    from lib.run_nerf_helpers import shifted_softplus
    ns = dict(torch=torch, np=np, os=os, knn_points=knn_points, shifted_softplus=shifted_softplus,
              read_pickle=lambda path: {"f": faces}, data_root="synthetic_M")
    exec(compile(helpers, "extract_thuman_mesh.py[19:41]", "exec"), ns)
    ns.update(flat=torch.from_numpy(flat_np.copy()), sh=sh, t_vertices=verts.copy(),
              raw=torch.from_numpy(raw_np.copy()).reshape(list(sh[:-1]) + [4]))
    exec(compile(body, "extract_thuman_mesh.py[125:159]", "exec"), ns)
    out = os.path.join(ROOT, "tests", "golden", "mesh_post.npz")
    np.savez_compressed(out, flat=flat_np, raw=raw_np, verts=verts, faces=faces, sh=np.array(sh),
                        occupancy=np.asarray(ns["occupancy"], np.float32), pts_mask=ns["pts_mask"].numpy(),
                        outside_msk=ns["outside_msk"].numpy(), vert_ids=ns["vert_ids"].squeeze(0).numpy(),
                        normals=ns["smpl_pts_normal"].numpy())
    occ = np.asarray(ns["occupancy"])
    print("wrote", out, "points", len(flat_np), "mask frac", float(ns["pts_mask"].float().mean()),
          "occ==100 frac", float((occ == 100).mean()), "occ==0 frac", float((occ == 0).mean()))


if __name__ == "__main__":
    main()
