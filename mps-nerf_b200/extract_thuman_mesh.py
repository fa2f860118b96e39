"""Density-grid query + occupancy post-step of the reference's mesh extraction (extract_thuman_mesh.py:92-158).

The reference is a script; its per-frame body is mirrored here as functions with the same names for the
quantities it computes (``t_1/t_2/t_3``, ``START/SIZE/RANGE``, ``occupancy``, ``pts_mask``, ``outside_msk``).
Marching cubes itself (PyMCubes in the reference, :160) is left to the caller -- ``occupancy`` is exactly
the array the reference hands to ``mcubes.marching_cubes(occupancy, threshold)``.

All heavy work runs in the CUDA library: the grid query through the render engine (K1..K5 on points, a21 in
SURVEY 8), the post-step in one kernel (csrc/occupancy.cu).  There is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib

threshold = 30      # :71
N = 256             # :72


def normalize_v3(arr):
    """:19-27 (in place)."""
    lens = torch.sqrt(arr[:, 0] ** 2 + arr[:, 1] ** 2 + arr[:, 2] ** 2)
    lens[lens < 0.00000001] = 0.00000001
    arr[:, 0] /= lens
    arr[:, 1] /= lens
    arr[:, 2] /= lens
    return arr


def compute_normal(vertices, faces):
    """:29-41 -- vertex normals; like the reference this uses indexed ``+=`` (the last face of a vertex wins
    per corner slot, it does not accumulate), so the result is only meaningful as the reference defines it."""
    norm = torch.zeros(vertices.shape, dtype=vertices.dtype, device=vertices.device)
    tris = vertices[faces]
    n = torch.cross(tris[::, 1] - tris[::, 0], tris[::, 2] - tris[::, 0], dim=-1)
    normalize_v3(n)
    norm[faces[:, 0]] += n
    norm[faces[:, 1]] += n
    norm[faces[:, 2]] += n
    normalize_v3(norm)
    return norm


def grid_axes(can_flag=False, n=N):
    """:95-112 -> (t_1, t_2, t_3 float64 numpy, START, SIZE, RANGE)."""
    if can_flag:
        t_1, t_2, t_3 = np.linspace(-1.0, 1.0, n), np.linspace(-1.0, 1.0, n), np.linspace(-0.25, 0.25, n // 4)
        START, SIZE, RANGE = np.array([-1.0, -1.0, -0.25]), np.array([2.0, 2.0, 0.5]), np.array([n, n, n // 4])
    else:
        t_1, t_2, t_3 = np.linspace(0.0, 2.0, n), np.linspace(0.6, 2.6, n), np.linspace(0.0, 2.0, n)
        START, SIZE, RANGE = np.array([0.0, 0.6, 0.0]), np.array([2.0, 2.0, 2.0]), np.array([n, n, n])
    return t_1, t_2, t_3, START, SIZE, RANGE


def grid_points(can_flag=False, n=N, device=None):
    """:95-115 -> (query_pts (n,n,n|n/4,3) float32, START, SIZE, RANGE).

    ``device=None``: numpy, literally ``np.stack(np.meshgrid(t_1, t_2, t_3), -1).astype(np.float32)``.
    With a CUDA device the same array is assembled there by broadcasting the three float32-rounded axes
    (meshgrid only replicates values; 'xy' indexing puts t_2 on the first axis), which avoids building and
    uploading a 200 MB host array per frame."""
    t_1, t_2, t_3, START, SIZE, RANGE = grid_axes(can_flag, n)
    if device is None:
        return np.stack(np.meshgrid(t_1, t_2, t_3), -1).astype(np.float32), START, SIZE, RANGE
    a1, a2, a3 = (torch.from_numpy(t.astype(np.float32)).to(device) for t in (t_1, t_2, t_3))
    shape = (len(t_2), len(t_1), len(t_3))
    query_pts = torch.stack([a1[None, :, None].expand(shape), a2[:, None, None].expand(shape), a3[None, None, :].expand(shape)], -1)
    return query_pts, START, SIZE, RANGE


def occupancy_post(flat, raw, t_vertices, faces, want_debug=False, normals=None):
    """:125-158 for flat points (P,3), raw (P,>=4) and the mesh the reference tests against.

    Returns occupancy (P,) float32 CUDA tensor (and, with ``want_debug``, pts_mask, outside_msk, idx5, d2).
    ``normals`` overrides compute_normal(t_vertices, faces) (whose duplicate-index ``+=`` is not reproducible
    run to run, in the reference as here)."""
    if not flat.is_cuda:
        raise RuntimeError("mpsnerf_b200 has no CPU path: inputs must be CUDA tensors")
    lib = _lib.load()
    dev = flat.device
    flat = flat.float().contiguous()
    raw = raw.float().contiguous()
    verts = t_vertices.to(dev).float().contiguous()
    normals = (normals.to(dev).float() if normals is not None
               else compute_normal(verts, torch.as_tensor(faces, device=dev).long())).contiguous()
    P = flat.shape[0]
    occ = torch.empty(P, device=dev)
    mask = torch.empty(P, dtype=torch.int32, device=dev) if want_debug else None
    outside = torch.empty(P, dtype=torch.uint8, device=dev) if want_debug else None
    idx5 = torch.empty(P, 5, dtype=torch.int32, device=dev) if want_debug else None
    d2 = torch.empty(P, device=dev) if want_debug else None
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.mpsnerf_occupancy_fix(_lib.ptr(flat), P, _lib.ptr(verts), _lib.ptr(normals), verts.shape[0],
                                         _lib.ptr(raw), raw.shape[1], _lib.ptr(occ), _lib.ptr(mask), _lib.ptr(outside),
                                         _lib.ptr(idx5), _lib.ptr(d2), stream), "occupancy_fix")
    _lib.count_launches(1)
    return (occ, mask, outside, idx5, d2) if want_debug else occ


@torch.no_grad()
def estimate_occupancy(net_fn, sp_input, tp_input, faces, can_flag=False, n=N, chunk=200000 * 5):
    """The reference's per-frame block :92-158: grid -> network -> occupancy volume (numpy, as handed to
    marching cubes).  ``faces`` = SMPL ``f`` (:129-130)."""
    net = net_fn.module if hasattr(net_fn, "module") else net_fn
    dev = sp_input["img_all"].device
    query_pts, START, SIZE, RANGE = grid_points(can_flag, n, device=dev)
    sh = query_pts.shape
    flat = query_pts.reshape([-1, 3])
    if can_flag:
        net.set_extract_mesh(True)
        t_vertices = sp_input["t_vertices"].reshape(-1, 3)
    else:
        t_vertices = tp_input["vertices"].reshape(-1, 3)
    try:
        raw = torch.cat([net_fn(sp_input, tp_input, flat[i:i + chunk], torch.zeros_like(flat[i:i + chunk]))[0, ..., 0:4]
                         for i in range(0, flat.shape[0], chunk)], 0)
    finally:
        if can_flag:
            net.set_extract_mesh(False)
    occupancy = occupancy_post(flat, raw, t_vertices, faces)
    return occupancy.reshape(list(sh[:-1])).cpu().numpy(), START, SIZE, RANGE
