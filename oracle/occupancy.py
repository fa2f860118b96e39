"""CPU restatement of the mesh-extraction post-step (extract_thuman_mesh.py:19-41, 125-158).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header): imported by tests/ and nothing else.
Pinned by tests/golden/mesh_post.npz, which oracle/make_golden_mesh.py produces by executing the reference's
own lines (read from /root/reference at generation time) under the import shims of oracle/ref_shims.py.
kNN follows the repo-wide contract: d2 = (dx*dx + dy*dy) + dz*dz in fp32, ties -> lowest
index, neighbours sorted by (d2, index).
"""
import numpy as np
import torch
import torch.nn.functional as F


def normalize_v3(arr):                       # :19-27
    lens = torch.sqrt(arr[:, 0] ** 2 + arr[:, 1] ** 2 + arr[:, 2] ** 2)
    lens[lens < 0.00000001] = 0.00000001
    arr[:, 0] /= lens
    arr[:, 1] /= lens
    arr[:, 2] /= lens
    return arr


def compute_normal(vertices, faces):         # :29-41 (indexed +=: no accumulation over repeated indices)
    norm = torch.zeros(vertices.shape, dtype=vertices.dtype)
    tris = vertices[faces]
    n = torch.cross(tris[::, 1] - tris[::, 0], tris[::, 2] - tris[::, 0], dim=-1)
    normalize_v3(n)
    norm[faces[:, 0]] += n
    norm[faces[:, 1]] += n
    norm[faces[:, 2]] += n
    normalize_v3(norm)
    return norm


def knn5(flat, verts):
    """(P,5) d2 and indices, sorted by (d2, index); pinned fp32 distances."""
    q = np.asarray(flat, np.float32)
    v = np.asarray(verts, np.float32)
    d_out = np.empty((len(q), 5), np.float32)
    i_out = np.empty((len(q), 5), np.int64)
    for s in range(0, len(q), 4096):
        c = q[s:s + 4096]
        dx, dy, dz = c[:, 0:1] - v[None, :, 0], c[:, 1:2] - v[None, :, 1], c[:, 2:3] - v[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        order = np.argsort(d2, axis=1, kind="stable")[:, :5]          # stable: ties -> lowest index
        i_out[s:s + 4096] = order
        d_out[s:s + 4096] = np.take_along_axis(d2, order, 1)
    return d_out, i_out


def occupancy_post(flat, raw, t_vertices, faces=None, normals=None):
    """-> dict(occupancy (P,), pts_mask (P,) int32, outside (P,) bool, idx5 (P,5), dot (P,))  -- :125-163.

    ``normals``: vertex normals to use instead of compute_normal(t_vertices, faces).  The reference's
    compute_normal is NOT reproducible -- ``norm[faces[:, k]] += n`` with repeated indices keeps whichever
    duplicate the multi-threaded index_put happens to write last (run-to-run differences of up to a sign flip
    were observed here) -- so parity of the post-step is defined with the normals as an input."""
    flat_t = torch.as_tensor(np.asarray(flat, np.float32))
    verts = torch.as_tensor(np.asarray(t_vertices, np.float32))
    occupancy = F.softplus(torch.as_tensor(np.asarray(raw, np.float32))[..., 3] - 1)          # :125
    d5, i5 = knn5(flat, t_vertices)
    pts_mask = (d5[:, 0] < np.float32(0.05 ** 2)).astype(np.int32)                           # :132-137
    normals = (torch.as_tensor(np.asarray(normals, np.float32)) if normals is not None
               else compute_normal(verts.clone(), torch.as_tensor(np.asarray(faces)).long()))      # :146
    ids = torch.as_tensor(i5)
    pts_dir = flat_t - verts[ids].mean(dim=1)                                                # :150
    pts_dir = pts_dir / torch.norm(pts_dir, dim=-1, keepdim=True)
    face_normal = normals[ids].mean(dim=1)                                                   # :154
    dot = (pts_dir * face_normal).sum(dim=-1)
    outside = dot > 0                                                                        # :155
    m = torch.as_tensor(pts_mask)
    occupancy = occupancy.clone()
    occupancy[m == 0] = 0.                                                                   # :157
    occupancy[(m == 0) & (outside == 0)] = 100.                                              # :159
    return dict(occupancy=occupancy.numpy(), pts_mask=pts_mask, outside=outside.numpy(), idx5=i5, dot=dot.numpy(),
                normals=normals.numpy())
