"""Smooth-loss terms of the training step: normals of the occupancy field by double backward.

Reference: run_nerf_batch.py:60-79 (the second, perturbed network pass and the two losses) and
lib/skinnning_batch.py:408-412, 496-504 (``occ_normal = d wide_sigmoid(alpha) / d canonical_pts`` with
``create_graph=True``, normalised; the normal of the nearest template vertex beside it).  Every
``smooth_interval``-th step of the shipped configs takes this path.

What is differentiated.  The canonical point x_c is a leaf (:386); alpha depends on it through (a) the positional
code of x_c in front of the MLP and (b) the pixel-aligned tokens: x_c -> source pose (inverse / forward LBS with the
blend weights of the nearest template vertex, :253-300) -> cameras (:177-184) -> bilinear reads of the encoder latent
and of the images (:417-435) -> transformer -> MLP.  The nearest-vertex indices, the human-region mask and x_c itself
carry no gradient and come from the CUDA kernels of the render path (K1 / K3: pinned arithmetic, the same values the
oracle computes); the differentiable chain is restated here in plain torch ops, so that autograd can differentiate it
twice -- the loss is a function of the *gradient* of alpha, and its parameter gradients (MLP, transformer, encoder
trunk through the latent) need the second derivative of every stage, including the bilinear read (torch's fused
``grid_sample`` has no double backward, which is why the reference carries a gather-based one as well,
lib/encoder.py:12-62).  This first version of the step is torch-autograd on the few thousand active points of a
training batch, not hand-written kernels; the render forward / backward of the same step still runs on the kernels
(train.py).  Device-agnostic: tests/test_smooth_cpu.py runs these functions on CPU tensors against the reference's
own autograd (tests/golden/smooth_grads.npz); the product path only ever calls them with CUDA tensors.
"""
import torch

from . import _lib
from .lib.run_nerf_helpers import wide_sigmoid

# float offsets of the fields of mpsnerf_frame (include/mpsnerf.h)
_OFF = {}
_o = 0
for _name, _n in (("Th_tp", 3), ("R_tp", 9), ("Rinv_sp", 9), ("Th_sp", 3), ("A_tp", 288), ("A_big_tp", 288),
                  ("A_big_sp", 288), ("A_sp", 288), ("cam_R", _lib.MAX_VIEWS * 9), ("cam_T", _lib.MAX_VIEWS * 3),
                  ("cam_K", _lib.MAX_VIEWS * 9)):
    _OFF[_name] = (_o, _n)
    _o += _n


def frame_constants(frame_dev, n_views):
    """Views into the device-resident mpsnerf_frame that K0 filled (no copy, no synchronisation)."""
    f = frame_dev.view(torch.float32)
    get = lambda k: f[_OFF[k][0]:_OFF[k][0] + _OFF[k][1]]
    return {"A_big_sp": get("A_big_sp").view(24, 12), "A_sp": get("A_sp").view(24, 12), "Rinv_sp": get("Rinv_sp").view(3, 3),
            "Th_sp": get("Th_sp"), "cam_R": get("cam_R").view(-1, 3, 3)[:n_views], "cam_T": get("cam_T").view(-1, 3)[:n_views],
            "cam_K": get("cam_K").view(-1, 3, 3)[:n_views]}


def inverse3(A):
    """(n,3,3) -> inverses by the adjugate (the formula K3 uses, csrc/deform.cu): elementwise ops only, so no
    cuSOLVER call (and none of its host synchronisations) and a cheap double backward."""
    a, b, c, d, e, f, g, h, i = (A[:, r, k] for r in range(3) for k in range(3))
    c00, c01, c02 = e * i - f * h, c * h - b * i, b * f - c * e
    c10, c11, c12 = f * g - d * i, a * i - c * g, c * d - a * f
    c20, c21, c22 = d * h - e * g, b * g - a * h, a * e - b * d
    det = a * c00 + b * c10 + c * c20
    return torch.stack([c00, c01, c02, c10, c11, c12, c20, c21, c22], -1).view(-1, 3, 3) / det[:, None, None]


def canonical_to_pixels(xc, bw, fr):
    """x_c (n,3) -> pixel coordinates (V,n,2) in the input views: coarse_deform_c2source (:253-300, nearest vertex
    frozen: ``bw`` = its normalised blend weights (n,24)) and projection (:177-184).  fr: frame_constants()."""
    A = (bw @ fr["A_big_sp"]).view(-1, 3, 4)                      # big pose -> T pose, inverted
    q = xc - A[:, :, 3]
    q = (inverse3(A[:, :, :3]) * q[:, None]).sum(2)
    A = (bw @ fr["A_sp"]).view(-1, 3, 4)                          # T pose -> source pose
    s = (A[:, :, :3] * q[:, None]).sum(2) + A[:, :, 3]
    w = s @ fr["Rinv_sp"] + fr["Th_sp"]                           # SMPL space -> world (:297-298)
    cam = torch.matmul(w[None], fr["cam_R"].transpose(1, 2)) + fr["cam_T"][:, None]
    pix = torch.matmul(cam, fr["cam_K"].transpose(1, 2))
    return pix[..., :2] / (pix[..., 2:] + 1e-5)


def bilinear_read(image, uv, size):
    """image (V,C,IH,IW), uv (V,n,2) pixels of a ``size`` = (W, H) image -> (V,n,C).  The reference's gather-based
    bilinear read (lib/encoder.py:12-62 behind SpatialEncoder.index, :238-244): weights from the unclamped corner
    coordinates, reads clamped to the border; differentiable any number of times in uv and in the image."""
    V, C, IH, IW = image.shape
    g = 2.0 * uv / size - 1.0
    ix = ((g[..., 0] + 1) / 2) * (IW - 1)
    iy = ((g[..., 1] + 1) / 2) * (IH - 1)
    x0, y0 = torch.floor(ix.detach()), torch.floor(iy.detach())
    wx1, wy1 = ix - x0, iy - y0
    wx0, wy0 = (x0 + 1) - ix, (y0 + 1) - iy
    flat = image.reshape(V, C, IH * IW)

    def corner(dx, dy):
        xi = (x0 + dx).clamp(0, IW - 1)
        yi = (y0 + dy).clamp(0, IH - 1)
        idx = (yi * IW + xi).long()[:, None, :].expand(V, C, -1)
        return torch.gather(flat, 2, idx)                        # (V,C,n)

    out = corner(0, 0) * (wx0 * wy0)[:, None] + corner(1, 0) * (wx1 * wy0)[:, None] + \
        corner(0, 1) * (wx0 * wy1)[:, None] + corner(1, 1) * (wx1 * wy1)[:, None]
    return out.transpose(1, 2)


def alpha_of_canonical(net, fr, latent, img, xc, bw):
    """Density logit of the active points as a differentiable function of x_c (lib/skinnning_batch.py:417-466, the
    alpha branch only).  latent (V,128,Hf,Wf) NCHW under autograd, img (V,3,H,W), xc (n,3), bw (n,24)."""
    H, W = img.shape[-2:]
    size = torch.tensor([float(W), float(H)], device=xc.device, dtype=xc.dtype)
    uv = canonical_to_pixels(xc, bw, fr)
    feat = bilinear_read(latent.to(xc.dtype), uv, size)                                   # (V,n,128)
    rgb = bilinear_read(img.to(xc.dtype), uv, size)                                       # (V,n,3)
    V, n = rgb.shape[:2]
    code = net.view_enc(rgb.reshape(-1, 3)).reshape(V, n, 27)
    tok = net.transformer(torch.cat((feat, code), -1).transpose(0, 1))               # (n,V,155)
    x = torch.cat((net.pos_enc(xc), tok[:, 0]), 1)
    h = x
    for i, lin in enumerate(net.pts_linears):
        h = torch.relu(lin(h))
        if i in net.skips:
            h = torch.cat([x, h], -1)
    return net.alpha_linear(h)


def occupancy_normals(net, fr, latent, img, xc, bw):
    """:496-499 -> (n,3) unit gradient of wide_sigmoid(alpha) with respect to x_c, still attached to the graph."""
    xc = xc.detach().clone().requires_grad_(True)
    occ = wide_sigmoid(alpha_of_canonical(net, fr, latent, img, xc, bw))
    g = torch.autograd.grad(occ, [xc], grad_outputs=torch.ones_like(occ), create_graph=True)[0]
    return g / (torch.norm(g, dim=-1, keepdim=True) + 1e-8)


def normal_fields(net, fr, latent, img, skin_w, vertex_normals, n_points, first, second):
    """The fields the network output carries in columns 17:23 on a smooth step (:484-503), for the two passes of
    run_nerf_batch.py:62-67 at once: ``first`` / ``second`` = dict(act_pid (n), xc (n,3), idx3 (n)) of the unperturbed
    and of the perturbed sample points.  -> (occ0, smpl0, occ1), each (n_points,3): occupancy normal of pass 0, normal
    of the nearest template vertex of pass 0, occupancy normal of pass 1; zero outside the human region.  The active
    points of both passes go through the chain as ONE batch (every stage is per point), which halves its launches."""
    dev, dt = first["xc"].device, first["xc"].dtype
    n0 = first["act_pid"].numel()
    xc = torch.cat((first["xc"], second["xc"]), 0)
    idx3 = torch.cat((first["idx3"], second["idx3"]), 0).long()
    zeros = lambda: torch.zeros(n_points, 3, device=dev, dtype=dt)
    if xc.shape[0] == 0:
        return zeros(), zeros(), zeros()
    bw = skin_w[idx3]
    bw = bw / bw.sum(-1, keepdim=True)                                               # :261-262
    occ = occupancy_normals(net, fr, latent, img, xc, bw)
    pid0, pid1 = first["act_pid"].long(), second["act_pid"].long()
    return (zeros().index_put((pid0,), occ[:n0]), zeros().index_put((pid0,), vertex_normals[idx3[:n0]]),
            zeros().index_put((pid1,), occ[n0:]))


def smooth_losses(occ0, smpl0, occ1):
    """run_nerf_batch.py:66-78 -> other_loss (1,4): [0.1 * normal_smooth + 0.1 * smpl_normal, normal_smooth, 0,
    smpl_normal].  occ0 / smpl0: fields of the unperturbed pass, occ1: occupancy normals at the perturbed points;
    any leading batch shape (the means run over all subjects, as on the gathered DataParallel output)."""
    normal_smooth = torch.mean((occ1 - occ0) ** 2)
    smpl_normal = torch.mean((smpl0 + occ0) ** 2)
    zero = torch.zeros((), device=occ0.device, dtype=occ0.dtype)
    return torch.stack([0.1 * normal_smooth + 0.1 * smpl_normal, normal_smooth, zero, smpl_normal]).reshape(1, 4)


def sample_points(ray_batch, t_vals, u):
    """World positions of the samples of a ray batch (run_nerf_batch.py:406-424; K1 generates the same points on the
    fly and never stores them): ray_batch (..., >=8) = o, d, near, far; u (..., S) uniforms or None -> (..., S, 3)."""
    o, d = ray_batch[..., 0:3], ray_batch[..., 3:6]
    near, far = ray_batch[..., 6:7], ray_batch[..., 7:8]
    z = near * (1.0 - t_vals) + far * t_vals
    if u is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * u
    return o[..., None, :] + d[..., None, :] * z[..., :, None]


def vertex_normals(vertices, faces):
    """compute_normal of the reference (lib/skinnning_batch.py:29-41) with its scatter made deterministic.  The
    reference writes ``norm[faces[:, s]] += n`` for the three corner slots s: an indexed read-modify-write that does
    NOT accumulate over repeated indices -- of all faces that have vertex v at corner s, one contribution survives,
    and which one is a race on a multi-threaded CPU and on the GPU (two runs of the reference disagree with each
    other).  Executed sequentially (one thread, or numpy) the LAST such face wins; that is the definition used here,
    written as an arg-max over face indices so that it is the same on every device: normal(v) = normalise(sum over s of
    n[last face with v at corner s]).  tests/golden/smooth_grads.npz is generated with one CPU thread."""
    nv, nf = vertices.shape[0], faces.shape[0]
    faces = faces.long()
    tri = vertices[faces]
    n = torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=-1)
    n = n / torch.sqrt((n ** 2).sum(-1, keepdim=True)).clamp_min(1e-8)
    ids = torch.arange(nf, device=vertices.device)
    out = torch.zeros_like(vertices)
    for s in range(3):
        last = torch.full((nv,), -1, dtype=torch.long, device=vertices.device).scatter_reduce(0, faces[:, s], ids, "amax")
        out = out + torch.where((last >= 0)[:, None], n[last.clamp_min(0)], torch.zeros_like(out))
    return out / torch.sqrt((out ** 2).sum(-1, keepdim=True)).clamp_min(1e-8)
