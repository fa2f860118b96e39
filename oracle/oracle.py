"""CPU ORACLE for the MPS-NeRF render hot path.  TEST INFRASTRUCTURE ONLY.

This file is a clean-room CPU restatement of the reference algorithm and exists
only as the *checker*: it may be imported by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs.  The product package never imports it and has no CPU path.

Parity status: PINNED against the reference itself.  The reference ships no
tests or golden vectors, so ``oracle/make_golden.py`` runs the
reference's own files (under the import shims of ``oracle/ref_shims.py``) in the
build container and commits its outputs to ``tests/golden/``;
``tests/test_oracle_vs_golden.py`` checks this restatement against them.  One
dependency of the reference, ``pytorch3d.ops.knn_points`` (un-vendored, no
version pinned, ``lib/skinnning_batch.py:4``), is absent offline: its published
semantics (K nearest by squared L2, ascending, returns (dists, idx, nn)) are
restated here and in the shim, so the nearest-vertex contract is defined by this
file: ``d2 = (dx*dx + dy*dy) + dz*dz`` with every fp32 operation rounded
separately (no FMA), ties -> lowest vertex index, mask = ``d2 < fl32(0.05**2)``.

"Pinned arithmetic": every stage whose result feeds an integer output (sample
positions, world->SMPL, distances, LBS blends, 3x3 inverses) is written as an
explicit sequence of individually rounded fp32 operations, identical to the
sequence the CUDA kernels execute with ``__fmul_rn/__fadd_rn/...``.  That makes
mask / vertex indices bit-exact by construction rather than by luck.

Reference files followed (all under /root/reference):
  run_nerf_batch.py:42-135, 369-444      render / render_rays / raw2outputs
  lib/skinnning_batch.py:177-300, 333-514  projection, inverse LBS, forward
  lib/encoder.py:12-62, 225-306          grid_sample, SpatialEncoder
  lib/transformer.py:13-86               cross-view transformer
  lib/run_nerf_helpers.py:174-254, 313-353  SMPL chain, positional encoding
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32
THRESH = F32(0.05 ** 2)


# --------------------------------------------------------------------------- per-frame
def rodrigues(poses):
    """ref lib/run_nerf_helpers.py:174-192"""
    angle = torch.norm(poses + 1e-8, p=2, dim=1, keepdim=True)
    axis = poses / angle
    c, s = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(axis, 1, dim=1)
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).reshape(-1, 3, 3)
    return torch.eye(3)[None] + s * K + (1 - c) * torch.matmul(K, K)


def smpl_transforms(smpl, poses, shapes):
    """A (24,4,4): ref lib/run_nerf_helpers.py:195-254 (shape blend, joints, chain, rest removal)."""
    v_shaped = smpl["v_template"] + torch.sum(smpl["shapedirs"] * shapes.reshape(1, 1, 10), dim=2).float()
    rot = rodrigues(poses.reshape(-1, 3))
    joints = torch.matmul(smpl["J_regressor"], v_shaped)
    parents = smpl["kintree_table"][0]
    rel = joints.clone()
    rel[1:] -= joints[parents[1:]]
    T = torch.cat([torch.cat([rot, rel[..., None]], dim=2), torch.tensor([0.0, 0, 0, 1]).expand(24, 1, 4)], dim=1)
    chain = [T[0]]
    for i in range(1, 24):
        chain.append(torch.matmul(chain[int(parents[i])], T[i]))
    G = torch.stack(chain)
    jh = torch.cat([joints, torch.zeros(24, 1)], dim=1)
    G[..., 3] = G[..., 3] - torch.sum(G * jh[:, None], dim=2)
    return G


def big_pose():
    """ref lib/skinnning_batch.py:193-201"""
    p = torch.zeros(1, 72)
    p[0, 5] = 45 / 180 * torch.tensor(np.pi)
    p[0, 8] = -45 / 180 * torch.tensor(np.pi)
    p[0, 23] = -30 / 180 * torch.tensor(np.pi)
    p[0, 26] = 30 / 180 * torch.tensor(np.pi)
    return p


def _a12(A):
    return A[:, :3, :].reshape(24, 12).numpy().astype(F32)


def _rodrigues64(r):
    """ref lib/run_nerf_helpers.py:174-192 in float64; r (24,3)."""
    angle = np.sqrt(((r + 1e-8) ** 2).sum(1, keepdims=True))
    k = r / angle
    s, c = np.sin(angle)[:, :, None], (1.0 - np.cos(angle))[:, :, None]
    z = np.zeros(len(r))
    K = np.stack([z, -k[:, 2], k[:, 1], k[:, 2], z, -k[:, 0], -k[:, 1], k[:, 0], z], 1).reshape(-1, 3, 3)
    return np.eye(3)[None] + s * K + c * np.einsum("nik,nkj->nij", K, K)


def smpl_transforms64(smpl, poses, shapes):
    """``smpl_transforms`` (ref lib/run_nerf_helpers.py:195-254) evaluated in float64 on the fp32 inputs and
    rounded to fp32 once at the end: the arithmetic contract of K0 (csrc/frame_prep.cu).  Returns (24,12) fp32."""
    f64 = lambda t: (t.detach().numpy() if torch.is_tensor(t) else np.asarray(t)).astype(F32).astype(np.float64)
    v_shaped = f64(smpl["v_template"]) + (f64(smpl["shapedirs"]) * f64(shapes).reshape(1, 1, 10)).sum(2)
    joints = f64(smpl["J_regressor"]) @ v_shaped
    rot = _rodrigues64(f64(poses).reshape(24, 3))
    parents = [int(p) for p in smpl["kintree_table"][0]]
    G = np.zeros((24, 3, 4))
    for j in range(24):
        T = np.concatenate([rot[j], (joints[j] - (joints[parents[j]] if j else 0.0))[:, None]], 1)
        if j == 0:
            G[0] = T
        else:
            Gp = G[parents[j]]
            G[j, :, :3] = Gp[:, :3] @ T[:, :3]
            G[j, :, 3] = Gp[:, :3] @ T[:, 3] + Gp[:, 3]
    A = G.copy()
    A[:, :, 3] = G[:, :, 3] - np.einsum("jab,jb->ja", G[:, :, :3], joints)
    return A.reshape(24, 12).astype(F32)


def frame_constants(smpl, sp, tp, fp32_reference=False):
    """Everything that is computed once per (source, target) pair.

    ``sp``/``tp`` are already squeezed (no batch dim).  Returns numpy fp32 arrays.  Default = the float64 ->
    fp32 contract that K0 reproduces bit for bit; ``fp32_reference=True`` evaluates the same formulas with fp32
    torch ops exactly as the reference does (the two agree to ~1 fp32 ulp: tests/test_oracle_vs_golden.py).
    """
    tpp, spp = tp["params"], sp["params"]
    if fp32_reference:
        xf = lambda poses, shapes: _a12(smpl_transforms(smpl, poses, shapes))
        rinv = torch.inverse(spp["R"].reshape(3, 3)).numpy().astype(F32)
    else:
        xf = lambda poses, shapes: smpl_transforms64(smpl, poses, shapes)
        rinv = np.linalg.inv(spp["R"].reshape(3, 3).numpy().astype(F32).astype(np.float64)).astype(F32)
    c = {
        "A_tp": xf(tpp["poses"], tpp["shapes"]),
        "A_big_tp": xf(big_pose(), tpp["shapes"]),
        "A_big_sp": xf(big_pose(), spp["shapes"]),
        "A_sp": xf(spp["poses"], spp["shapes"]),
        "R_tp": tpp["R"].numpy().astype(F32).reshape(3, 3),
        "Th_tp": tpp["Th"].numpy().astype(F32).reshape(3),
        "Rinv_sp": rinv,
        "Th_sp": spp["Th"].numpy().astype(F32).reshape(3),
        "W": smpl["weights"].numpy().astype(F32),
        "t_vertices": sp["t_vertices"].numpy().astype(F32),
    }
    c["verts_smpl"] = world_to_smpl(tp["vertices"].numpy().astype(F32), c["Th_tp"], c["R_tp"])
    return c


# --------------------------------------------------------------------------- pinned stages
def sample_z(near, far, S, u=None):
    """ref run_nerf_batch.py:411-422.  near/far (N,), u optional (N,S) uniforms."""
    t = torch.linspace(0.0, 1.0, steps=S).numpy().astype(F32)
    near, far = near.astype(F32)[:, None], far.astype(F32)[:, None]
    z = near * (F32(1.0) - t) + far * t
    if u is not None:
        mids = F32(0.5) * (z[:, 1:] + z[:, :-1])
        upper = np.concatenate([mids, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mids], -1)
        z = lower + (upper - lower) * u.astype(F32)
    return z.astype(F32)


def sample_points(o, d, z):
    """ref run_nerf_batch.py:424: p = o + d*z, mul and add rounded separately."""
    return (o[:, None, :] + d[:, None, :] * z[:, :, None]).astype(F32)


def world_to_smpl(p, Th, R):
    """ref lib/skinnning_batch.py:347: q = (p - Th) @ R as ((d0*R0k + d1*R1k) + d2*R2k)."""
    dd = (p - Th).astype(F32)
    out = np.empty_like(dd)
    for k in range(3):
        out[..., k] = (dd[..., 0] * R[0, k] + dd[..., 1] * R[1, k]) + dd[..., 2] * R[2, k]
    return out


def knn1(q, verts, chunk=8192):
    """K=1 nearest vertex (restates pytorch3d knn_points K=1 as used at
    lib/skinnning_batch.py:214,256,357).  Returns (d2min fp32, idx int64)."""
    qt, vt = torch.from_numpy(np.ascontiguousarray(q)), torch.from_numpy(np.ascontiguousarray(verts))
    d2o = torch.empty(len(qt))
    ido = torch.empty(len(qt), dtype=torch.int64)
    vx, vy, vz = vt[:, 0][None], vt[:, 1][None], vt[:, 2][None]
    for s in range(0, len(qt), chunk):
        c = qt[s:s + chunk]
        dx, dy, dz = c[:, 0:1] - vx, c[:, 1:2] - vy, c[:, 2:3] - vz
        d2 = dx * dx
        d2 += dy * dy
        d2 += dz * dz
        m, i = torch.min(d2, dim=1)
        d2o[s:s + chunk], ido[s:s + chunk] = m, i
    return d2o.numpy(), ido.numpy()


def _blend(w, A12):
    """Sequential 24-term blend, each product and sum rounded: (P,24)x(24,12)->(P,12)."""
    acc = w[:, 0:1] * A12[0][None]
    for j in range(1, 24):
        acc = acc + w[:, j:j + 1] * A12[j][None]
    return acc.astype(F32)


def _inv3(M):
    """Adjugate inverse of (P,12) blended transforms' 3x3 part; op order is the contract."""
    a = lambda r, c: M[:, 4 * r + c]
    c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)
    c01 = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)
    c02 = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)
    c10 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)
    c11 = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)
    c12 = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)
    c20 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)
    c21 = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)
    c22 = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)
    det = (a(0, 0) * c00 + a(0, 1) * c10) + a(0, 2) * c20
    r = F32(1.0) / det
    return [[c00 * r, c01 * r, c02 * r], [c10 * r, c11 * r, c12 * r], [c20 * r, c21 * r, c22 * r]]


def _apply_inv(M, x):
    inv = _inv3(M)
    v = [x[:, k] - M[:, 4 * k + 3] for k in range(3)]
    return np.stack([(inv[i][0] * v[0] + inv[i][1] * v[1]) + inv[i][2] * v[2] for i in range(3)], 1).astype(F32)


def _apply_fwd(M, x):
    return np.stack([((M[:, 4 * i] * x[:, 0] + M[:, 4 * i + 1] * x[:, 1]) + M[:, 4 * i + 2] * x[:, 2]) + M[:, 4 * i + 3]
                     for i in range(3)], 1).astype(F32)


def target2canonical(q, idx2, c):
    """ref lib/skinnning_batch.py:203-251 (mean_shape=0)."""
    w = c["W"][idx2]
    xT = _apply_inv(_blend(w, c["A_tp"]), q)
    return _apply_fwd(_blend(w, c["A_big_tp"]), xT)


def canonical2source(xc, c):
    """ref lib/skinnning_batch.py:253-300 (weights_correction=0, mean_shape=0)."""
    _, idx3 = knn1(xc, c["t_vertices"])
    w = c["W"][idx3]
    s = w[:, 0].copy()
    for j in range(1, 24):
        s = s + w[:, j]
    w = (w / s[:, None]).astype(F32)
    xT = _apply_inv(_blend(w, c["A_big_sp"]), xc)
    xs = _apply_fwd(_blend(w, c["A_sp"]), xT)
    Ri, Th = c["Rinv_sp"], c["Th_sp"]
    xw = np.stack([((xs[:, 0] * Ri[0, k] + xs[:, 1] * Ri[1, k]) + xs[:, 2] * Ri[2, k]) + Th[k] for k in range(3)], 1)
    return idx3, xs, xw.astype(F32), w


# --------------------------------------------------------------------------- tolerance stages
def projection(xw, R_all, T_all, K_all):
    """ref lib/skinnning_batch.py:177-184 -> (V,P,2) pixel coords."""
    x = torch.from_numpy(xw)[None].expand(R_all.shape[0], -1, -1)
    cam = torch.bmm(x, R_all.transpose(1, 2)) + T_all.transpose(1, 2)
    img = torch.bmm(cam, K_all.transpose(1, 2))
    return img[:, :, :2] / (img[:, :, 2:] + 1e-5)


def bilinear_border(image, uv, size_wh):
    """ref lib/encoder.py:12-62 + :225-253: image (V,C,IH,IW), uv (V,P,2) px -> (V,P,C)."""
    V, C, IH, IW = image.shape
    g = 2.0 * uv / torch.tensor(size_wh, dtype=torch.float32) - 1.0
    ix = ((g[..., 0] + 1) / 2) * (IW - 1)
    iy = ((g[..., 1] + 1) / 2) * (IH - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wnw, wne = (x1 - ix) * (y1 - iy), (ix - x0) * (y1 - iy)
    wsw, wse = (x1 - ix) * (iy - y0), (ix - x0) * (iy - y0)
    cx = lambda t: t.clamp(0, IW - 1).long()
    cy = lambda t: t.clamp(0, IH - 1).long()
    flat = image.reshape(V, C, IH * IW)
    tap = lambda yy, xx: torch.gather(flat, 2, (cy(yy) * IW + cx(xx))[:, None, :].expand(-1, C, -1))
    out = tap(y0, x0) * wnw[:, None] + tap(y0, x1) * wne[:, None] + tap(y1, x0) * wsw[:, None] + tap(y1, x1) * wse[:, None]
    return out.transpose(1, 2)


def posenc(x, nfreq):
    """ref lib/run_nerf_helpers.py:313-353"""
    freqs = torch.repeat_interleave(np.pi * 2.0 ** torch.arange(0, nfreq), 2).view(1, -1, 1)
    phases = torch.zeros(2 * nfreq)
    phases[1::2] = np.pi * 0.5
    e = x.unsqueeze(1).repeat(1, nfreq * 2, 1)
    e = torch.sin(torch.addcmul(phases.view(1, -1, 1), e, freqs)).reshape(x.shape[0], nfreq * 6)
    return torch.cat((x, e), dim=-1)


def encode_images(img_all, sd):
    """ref lib/encoder.py:256-306 with num_layers=2, feature_scale=0.5 (eval-mode BN)."""
    p = "encoder_2d.model."
    x = F.interpolate(img_all, scale_factor=0.5, mode="area", recompute_scale_factor=True)

    def bn(x, n):
        return F.batch_norm(x, sd[p + n + ".running_mean"], sd[p + n + ".running_var"], sd[p + n + ".weight"],
                            sd[p + n + ".bias"], False, 0.0, 1e-5)

    x = F.relu(bn(F.conv2d(x, sd[p + "conv1.weight"], None, 2, 3), "bn1"))
    lat0 = x
    for b in range(3):
        y = F.relu(bn(F.conv2d(x, sd[p + f"layer1.{b}.conv1.weight"], None, 1, 1), f"layer1.{b}.bn1"))
        y = bn(F.conv2d(y, sd[p + f"layer1.{b}.conv2.weight"], None, 1, 1), f"layer1.{b}.bn2")
        x = F.relu(x + y)
    lat1 = F.interpolate(x, lat0.shape[-2:], mode="bilinear", align_corners=True)
    lat0 = F.interpolate(lat0, lat0.shape[-2:], mode="bilinear", align_corners=True)
    return torch.cat([lat0, lat1], dim=1)


class _Lin:
    """fp32 linear, or its bf16-operand emulation (inputs and weights rounded to bf16,
    fp32 accumulate) used to test the tensor-core path tightly."""

    def __init__(self, bf16):
        self.bf16 = bf16

    def __call__(self, x, w, b=None):
        if self.bf16:
            x, w = x.bfloat16().float(), w.bfloat16().float()
        return F.linear(x, w, b)


def _ln_linear(x, g, b, w, bias, lin):
    """linear(LayerNorm(x)).  fp32: literally that.  bf16 emulation: the tensor-core path folds the LayerNorm affine into
    the weights (mps-nerf_b200/pack.py): operand = bf16(xhat), weights = bf16(W diag(g)), and W b (+ the layer's own
    bias) is a bf16 K column against a constant 1.0 of the operand."""
    if not lin.bf16:
        return F.linear(F.layer_norm(x, (x.shape[-1],), g, b, 1e-5), w, bias)
    xh = F.layer_norm(x, (x.shape[-1],), None, None, 1e-5).bfloat16().float()
    wb = w @ b
    extra = (wb if bias is None else (bias + wb)).bfloat16().float()       # a bf16 K column against a constant 1.0
    return F.linear(xh, (w * g[None, :]).bfloat16().float(), None) + extra


def transformer(tok, sd, lin):
    """ref lib/transformer.py:13-86: tok (P,V,155) -> (P,V,155); depth 2, 4 heads x 64."""
    x = tok
    for l in range(2):
        p = f"transformer.layers.{l}."
        qkv = _ln_linear(x, sd[p + "0.fn.norm.weight"], sd[p + "0.fn.norm.bias"], sd[p + "0.fn.fn.to_qkv.weight"], None, lin)
        P, V, _ = qkv.shape
        q, k, v = (t.reshape(P, V, 4, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
        att = torch.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * (64 ** -0.5), dim=-1)      # (attention itself stays fp32)
        o = torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).reshape(P, V, 256)
        if lin.bf16:    # the tensor-core path runs the out-projection as an fp16 x fp16 GEMM (o is produced in fp16)
            x = x + F.linear(o.half().float(), sd[p + "0.fn.fn.to_out.0.weight"].half().float(), sd[p + "0.fn.fn.to_out.0.bias"])
        else:
            x = x + lin(o, sd[p + "0.fn.fn.to_out.0.weight"], sd[p + "0.fn.fn.to_out.0.bias"])
        h = F.gelu(_ln_linear(x, sd[p + "1.fn.norm.weight"], sd[p + "1.fn.norm.bias"], sd[p + "1.fn.fn.net.0.weight"],
                              sd[p + "1.fn.fn.net.0.bias"], lin))
        x = x + lin(h, sd[p + "1.fn.fn.net.3.weight"], sd[p + "1.fn.fn.net.3.bias"])
    return x


def nerf_mlp(xc, tok0, tok1, sd, lin):
    """ref lib/skinnning_batch.py:449-473 -> rgb (P,3), alpha (P,1)."""
    x = torch.cat((posenc(xc, 6), tok0), dim=1)
    h = x
    for i in range(8):
        h = F.relu(lin(h, sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"]))
        if i == 4:
            h = torch.cat([x, h], -1)
    alpha = F.linear(h, sd["alpha_linear.weight"], sd["alpha_linear.bias"]) if not lin.bf16 else \
        lin(h, sd["alpha_linear.weight"], sd["alpha_linear.bias"])
    feat = lin(h, sd["feature_linear.weight"], sd["feature_linear.bias"])
    g = F.relu(lin(torch.cat([feat, tok1], -1), sd["views_linear.weight"], sd["views_linear.bias"]))
    rgb = lin(g, sd["rgb_linear.weight"], sd["rgb_linear.bias"])
    return rgb, alpha


# --------------------------------------------------------------------------- network forward
def squeeze_inputs(sp, tp):
    """ref lib/run_nerf_helpers.py:152-172, non-mutating."""
    def walk(d):
        return {k: (v.squeeze(0).float() if torch.is_tensor(v) else walk(v) if isinstance(v, dict) else v) for k, v in d.items()}
    return walk(sp), walk(tp)


def smpl_tensors(smpl_np):
    """ref lib/run_nerf_helpers.py:141-150"""
    out = {}
    for k in ("v_template", "shapedirs", "weights", "posedirs"):
        out[k] = torch.tensor(np.array(smpl_np[k]).astype(float)).float()
    J = smpl_np["J_regressor"]
    out["J_regressor"] = torch.tensor((J.toarray() if hasattr(J, "toarray") else np.asarray(J)).astype(float)).float()
    for k in ("kintree_table", "f"):
        out[k] = torch.tensor(np.array(smpl_np[k]).astype(float)).long()
    return out


def forward_points(smpl, sd, sp, tp, pts, bf16=False, latent=None, consts=None, extract_mesh=False,
                   return_stages=False):
    """SKinningBatch.forward for the shipped configs (ref lib/skinnning_batch.py:333-514).

    pts (P,3) world fp32 numpy.  Returns raw (P,17) numpy (or (P,4) when
    ``extract_mesh``) and optionally the per-stage tensors.
    """
    lin = _Lin(bf16)
    c = consts or frame_constants(smpl, sp, tp)
    if latent is None:
        latent = encode_images(sp["img_all"], sd)
    P = len(pts)
    q_all = world_to_smpl(pts.astype(F32), c["Th_tp"], c["R_tp"])
    d2, idx_all = knn1(q_all, c["verts_smpl"])
    mask = d2 < THRESH
    act = np.nonzero(mask)[0]
    q, idx2 = q_all[act], idx_all[act]
    xc = target2canonical(q, idx2, c)
    if extract_mesh:                       # ref :394-396 canonical_pts = world_query_pts, mask all ones
        xc = pts.astype(F32)
        act = np.arange(P)
        mask = np.ones(P, dtype=bool)
    idx3, xs, xw, w2 = canonical2source(xc, c)
    st = {"mask": mask, "d2": d2, "active": act, "q": q, "idx2": idx2, "xc": xc, "idx3": idx3, "xs": xs, "xw": xw}
    raw4 = np.full((P, 4), -80.0, dtype=F32)
    if len(act):
        uv = projection(xw, sp["R_all"], sp["T_all"], sp["K_all"])
        H, Wd = sp["img_all"].shape[-2:]
        feat = bilinear_border(latent, uv, (Wd, H))                       # (V,Pa,128)
        rgbs = bilinear_border(sp["img_all"], uv, (Wd, H))                # (V,Pa,3)
        V = rgbs.shape[0]
        rgb_code = posenc(rgbs.reshape(-1, 3), 4).reshape(V, -1, 27)
        tok = torch.cat((feat, rgb_code), dim=-1).transpose(0, 1).contiguous()   # (Pa,V,155)
        tout = transformer(tok, sd, lin)
        rgb, alpha = nerf_mlp(torch.from_numpy(xc), tout[:, 0], tout[:, 1], sd, lin)
        raw4[act] = torch.cat([rgb, alpha], -1).numpy()
        st.update(uv=uv.numpy(), tokens=tok.numpy(), tok_out=tout.numpy())
    if extract_mesh:
        out = raw4
    else:
        out = np.zeros((P, 17), dtype=F32)
        out[:, :4] = raw4
        out[:, 4] = mask
        out[act, 11:14] = q
        out[act, 14:17] = xs
    return (out, st) if return_stages else out


# --------------------------------------------------------------------------- compositing / render
def raw2outputs(raw, z, rays_d, occupancy=False):
    """ref run_nerf_batch.py:369-398.  raw (N,S,4), z (N,S), rays_d (N,3) torch."""
    rgb = (1 + 2 * 0.0001) * torch.sigmoid(raw[..., :3]) - 0.0001
    if not occupancy:
        dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], -1)
        dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
        alpha = 1.0 - torch.exp(-F.softplus(raw[..., 3] - 1) * dists)
    else:
        alpha = (1 + 2 * 0.0001) * torch.sigmoid(raw[..., 3]) - 0.0001
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    w = alpha * T
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth = torch.sum(w * z, -1)
    acc = torch.sum(w, -1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    return rgb_map, disp, acc, w, depth


def render(smpl, sd, sp_b, tp_b, rays_o, rays_d, near, far, S=64, u=None, bf16=False, chunk=4096,
           return_stages=False, occupancy=False, white_bkgd=False):
    """run_nerf_batch.render for one subject (ref :100-135, :401-444); numpy in, dict of numpy out.
    ``occupancy`` = --occupancy 1 (ref :383-386), ``white_bkgd`` = ref :394-396."""
    sp, tp = squeeze_inputs(sp_b, tp_b)
    c = frame_constants(smpl, sp, tp)
    latent = encode_images(sp["img_all"], sd)
    N = len(rays_o)
    z = sample_z(near.reshape(-1), far.reshape(-1), S, u)
    out17 = np.empty((N, S, 17), dtype=F32)
    stages = []
    for s in range(0, N, chunk):
        p = sample_points(rays_o[s:s + chunk].astype(F32), rays_d[s:s + chunk].astype(F32), z[s:s + chunk])
        r = forward_points(smpl, sd, sp, tp, p.reshape(-1, 3), bf16=bf16, latent=latent, consts=c,
                           return_stages=return_stages)
        if return_stages:
            r, st = r
            st["offset"] = s * S
            stages.append(st)
        out17[s:s + chunk] = r.reshape(-1, S, 17)
    rgb, disp, acc, w, depth = raw2outputs(torch.from_numpy(out17[..., :4]), torch.from_numpy(z), torch.from_numpy(rays_d.astype(F32)),
                                           occupancy=occupancy)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    res = {"rgb_map": rgb.numpy(), "disp_map": disp.numpy(), "acc_map": acc.numpy(), "depth_map": depth.numpy(),
           "raw": out17[..., :4], "pts_mask": out17[..., 4:5], "smpl_query_pts": out17[..., 11:14],
           "smpl_src_pts": out17[..., 14:17], "z_vals": z}
    if return_stages:
        res["stages"] = stages
    return res
