"""Host-side helpers mirroring the names in the reference's ``lib/run_nerf_helpers.py``.

Only what the render hot path and its callers need:
activations/metrics (ref :15-19), ``SMPL_to_tensor`` (ref :141-150), the SMPL
rigid-transform chain (ref :174-254) and ``PositionalEncoding`` (ref :313-353).
Everything here is per-frame or per-model work; per-point work lives in CUDA.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SMPL_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)


def img2mse(x, y):
    return torch.mean((x - y) ** 2)


def mse2psnr(x):
    return -10.0 * torch.log(x) / math.log(10.0)


def to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def shifted_softplus(x):
    return F.softplus(x - 1)


def wide_sigmoid(x):
    return (1 + 2 * 0.0001) * torch.sigmoid(x) - 0.0001


def SMPL_to_tensor(params):
    """Convert a SMPL pickle dict to torch tensors (ref :141-150)."""
    out = dict(params)
    for key in ("v_template", "shapedirs", "J_regressor", "kintree_table", "f", "weights", "posedirs"):
        val = params[key]
        if key == "J_regressor":
            val = val.toarray() if hasattr(val, "toarray") else np.asarray(val)
            out[key] = torch.tensor(val.astype(float)).float()
        elif key in ("kintree_table", "f"):
            out[key] = torch.tensor(np.array(val).astype(float)).long()
        else:
            out[key] = torch.tensor(np.array(val).astype(float)).float()
    return out


def _walk(d, fn):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            out[k] = fn(v)
        elif isinstance(v, dict):
            out[k] = _walk(v, fn)
        else:
            out[k] = v
    return out


def to_cuda(device, sp_input, tp_input=None):
    """Move every tensor of the input dicts to ``device`` (ref :119-139)."""
    sp = _walk(sp_input, lambda t: t.to(device))
    if tp_input is None:
        return sp
    return sp, _walk(tp_input, lambda t: t.to(device))


def sequeeze_0(sp_input, tp_input=None):
    """Drop the DataLoader batch dim of every tensor (ref :152-172).

    Unlike the reference this returns new dicts and leaves the caller's alone.
    """
    fn = lambda t: torch.squeeze(t, 0).float()
    sp = _walk(sp_input, fn)
    if tp_input is None:
        return sp
    return sp, _walk(tp_input, fn)


def batch_rodrigues_torch(poses):
    """Axis-angle (N,3) -> rotation matrices (N,3,3); 1e-8 added before the norm (ref :174-192)."""
    angle = torch.norm(poses + 1e-8, p=2, dim=1, keepdim=True)
    axis = poses / angle
    c = torch.cos(angle)[:, None]
    s = torch.sin(angle)[:, None]
    rx, ry, rz = axis[:, 0:1], axis[:, 1:2], axis[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=poses.dtype, device=poses.device)[None]
    return eye + s * K + (1 - c) * torch.matmul(K, K)


def get_rigid_transformation_torch(rot_mats, joints, parents):
    """24-joint kinematic chain with the rest joints removed (ref :195-224)."""
    n = joints.shape[0]
    rel = joints.clone()
    rel[1:] = rel[1:] - joints[parents[1:]]
    top = torch.cat([rot_mats, rel[..., None]], dim=2)
    bottom = torch.zeros(n, 1, 4, dtype=joints.dtype, device=joints.device)
    bottom[..., 3] = 1
    local = torch.cat([top, bottom], dim=1)
    chain = [local[0]]
    for i in range(1, n):
        chain.append(torch.matmul(chain[int(parents[i])], local[i]))
    G = torch.stack(chain, dim=0)
    jh = torch.cat([joints, torch.zeros(n, 1, dtype=joints.dtype, device=joints.device)], dim=1)
    G = G.clone()
    G[..., 3] = G[..., 3] - torch.sum(G * jh[:, None], dim=2)
    return G


def get_transform_params_torch(smpl, params):
    """LBS transforms A (24,4,4) plus R, Th, joints for one pose (ref :227-254)."""
    dev = params["shapes"].device
    v_shaped = smpl["v_template"].to(dev) + torch.sum(smpl["shapedirs"].to(dev) * params["shapes"][None], dim=2).float()
    rot = batch_rodrigues_torch(params["poses"].reshape(-1, 3))
    joints = torch.matmul(smpl["J_regressor"].to(dev), v_shaped)
    parents = smpl["kintree_table"][0].to(dev)
    A = get_rigid_transformation_torch(rot, joints, parents)
    return A, params["R"], params["Th"], joints


class PositionalEncoding(torch.nn.Module):
    """NeRF positional encoding with the reference's buffer names (ref :313-353).

    Output order: input, then per frequency (sin xyz, cos xyz); cos is evaluated
    as sin(f*x + fl(pi/2)).
    """

    def __init__(self, num_freqs=6, d_in=3, freq_factor=np.pi, include_input=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.d_in = d_in
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = num_freqs * 2 * d_in + (d_in if include_input else 0)
        self.include_input = include_input
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        phases = torch.zeros(2 * num_freqs)
        phases[1::2] = np.pi * 0.5
        self.register_buffer("_phases", phases.view(1, -1, 1))

    def forward(self, x):
        emb = x.unsqueeze(1).repeat(1, self.num_freqs * 2, 1)
        emb = torch.sin(torch.addcmul(self._phases, emb, self._freqs))
        emb = emb.view(x.shape[0], self.num_freqs * 2 * self.d_in)
        if self.include_input:
            emb = torch.cat((x, emb), dim=-1)
        return emb
