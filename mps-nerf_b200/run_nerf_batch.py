"""Render entry points with the reference's names and contracts (run_nerf_batch.py:42-135,
301-444): ``render``, ``batchify_rays``, ``render_rays``, ``run_network``, ``raw2outputs``,
``create_nerf``.

Differences that a caller can observe are limited to what SURVEY.md section 8b allows:
the input dicts are never mutated, the arguments are not parsed at import time (call
``configure(args)``; defaults = the parser's), and the sampling + network + compositing of
``render_rays`` run as fused CUDA kernels instead of through ``network_query_fn``.
Returned ``extras`` hold the same keys/shapes/values; ``correction``/``correction_`` are
zero tensors (correction_field = 0 in the shipped configs).
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from .lib.run_nerf_helpers import shifted_softplus, wide_sigmoid, img2mse, mse2psnr, to8b  # noqa: F401 (API)
from .model_selection import return_model
from .parser_config import config_parser

global_args = config_parser().parse_args([])
density_actfn = shifted_softplus
rgb_actfn = wide_sigmoid


def configure(args):
    """Install the parsed flags (the reference does this at import, run_nerf_batch.py:23-24)."""
    global global_args
    global_args = args
    return args


class NetworkHandle(nn.Module):
    """What ``create_nerf`` returns as ``network_fn``: exposes ``.module`` like the reference's
    DataParallel/DDP wrapper (run_nerf_batch.py:344-350) without replicating anything."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, sp_input, tp_input, pts, viewdirs=None):
        if pts.dim() == 2:      # extract_thuman_mesh.py:122: flat (P,3) points with the DataLoader-batched dicts
            return self.module(sp_input, tp_input, pts, None)
        outs = []
        for b in range(pts.shape[0]):
            outs.append(self.module(_select(sp_input, b), _select(tp_input, b), pts[b], None))
        return torch.cat(outs, 0)


def _net_of(network_fn):
    return network_fn.module if hasattr(network_fn, "module") else network_fn


def _select(d, b):
    """Subject ``b`` of a batched input dict, batch dim dropped, nothing mutated."""
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            if not v.is_floating_point():      # gender / indices: a view (keeps the storage identity the gender cache
                out[k] = v[b] if v.dim() > 0 else v      # keys on; a dtype conversion would also cost a launch each)
            else:
                out[k] = v[b].float() if v.dim() > 0 else v
        elif isinstance(v, dict):
            out[k] = _select(v, b)
        else:
            out[k] = v
    return out


def run_network(inputs, viewdirs, fn, sp_input=None, tp_input=None):
    """ref :42-82.  inputs (B,C,S,3) -> (outputs (B,C,S,17), other_loss (1,4))."""
    flat = torch.reshape(inputs, [inputs.shape[0], -1, inputs.shape[-1]])
    out = fn(sp_input, tp_input, flat, None)
    out = torch.reshape(out, list(inputs.shape[:-1]) + [out.shape[-1]])
    net = _net_of(fn)
    if net.training and global_args.smooth_loss and torch.is_grad_enabled():
        raise NotImplementedError("the smooth-loss second pass runs inside render() (mps-nerf_b200/train.py + smooth.py: it "
                                  "needs the active-point lists of both passes); run_network() itself is forward only")
    return out, torch.zeros(1, 4, device=inputs.device)


def raw2outputs(raw, z_vals, rays_d, white_bkgd=False):
    """ref :369-398 -> (rgb_map, disp_map, acc_map, weights, depth_map, T_s); raw (B,C,S,4)."""
    lib = _lib.load()
    B, C, S = z_vals.shape
    dev = raw.device
    rays8 = torch.zeros(B * C, 8, device=dev)
    rays8[:, 3:6] = rays_d.reshape(-1, 3)
    rawc = raw.reshape(-1, 4).float().contiguous()
    z = z_vals.reshape(-1, S).float().contiguous()
    rgb, disp, acc, depth = (torch.empty(B * C, 3, device=dev), torch.empty(B * C, device=dev),
                             torch.empty(B * C, device=dev), torch.empty(B * C, device=dev))
    w, ts = torch.empty(B * C, S, device=dev), torch.empty(B * C, S, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mpsnerf_composite(_lib.ptr(rawc), _lib.ptr(rays8), B * C, S, None, None, _lib.ptr(z),
                                     1 if global_args.occupancy else 0, _lib.ptr(rgb), _lib.ptr(disp), _lib.ptr(acc),
                                     _lib.ptr(depth), _lib.ptr(w), _lib.ptr(ts), st), "composite")
    _lib.count_launches(1)
    rgb, acc = rgb.reshape(B, C, 3), acc.reshape(B, C)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    return rgb, disp.reshape(B, C), acc, w.reshape(B, C, S), depth.reshape(B, C), ts.reshape(B, C, S)


def render_rays(ray_batch, network_fn, network_query_fn=None, N_samples=64, perturb=0.0, N_importance=0,
                network_fine=None, white_bkgd=False, sp_input=None, tp_input=None, perturb_u=None, _rays_ready=None,
                smooth_delta=None):
    """ref :401-444.  ray_batch (B,C,8|11) = [o, d, near, far(, viewdirs)] -> dict of outputs.

    ``perturb_u`` (B,C,S) optionally supplies the stratified-sampling uniforms (otherwise drawn
    with torch.rand on the device, as the reference draws them from the global RNG); ``smooth_delta`` (B,C*S,3)
    likewise supplies the perturbation of the smooth-loss second pass (ref :64, N(0, 0.01) otherwise).
    """
    net = _net_of(network_fn)
    if net.training and torch.is_grad_enabled():
        # training step (BASELINE config 4): forward with saved intermediates + hand-written backward, train.py
        # every smooth_interval-th step adds the normal-smoothness terms (ref :60-79): mps-nerf_b200/smooth.py
        smooth_step = bool(global_args.smooth_loss) and bool(getattr(net, "smooth_loss", False)) and sp_input is not None \
            and "smooth_interval" in sp_input and "global_step" in sp_input \
            and int(sp_input["global_step"].reshape(-1)[0]) % int(sp_input["smooth_interval"].reshape(-1)[0]) == 0
        from .train import render_rays_train
        if _rays_ready is not None:
            torch.cuda.current_stream(ray_batch.device).wait_event(_rays_ready)
        return render_rays_train(net, ray_batch, sp_input, tp_input, N_samples, perturb, perturb_u, white_bkgd,
                                 bool(global_args.occupancy), _select, smooth_step=smooth_step, smooth_delta=smooth_delta)
    B, C = ray_batch.shape[:2]
    dev = ray_batch.device
    S = int(N_samples)
    t_vals = torch.linspace(0.0, 1.0, steps=S, device=dev)
    u = None
    if perturb > 0.0:
        u = (perturb_u if perturb_u is not None else torch.rand(B, C, S, device=dev)).float().contiguous()
    eng = net.engine()
    net.invalidate_frame_cache()      # like the reference, every render_rays call re-derives the per-frame state
    per = []
    for b in range(B):
        sp, tp = _select(sp_input, b), _select(tp_input, b)
        ctx = net.frame_context(sp, tp, n_points=C * S)
        if _rays_ready is not None:       # rays uploaded on the copy stream while the frame was being prepared
            torch.cuda.current_stream(dev).wait_event(_rays_ready)
            _rays_ready = None
        rays8 = ray_batch[b, :, :8].float().contiguous()
        per.append(eng.run(ctx, rays8=rays8, S=S, t_vals=t_vals, u=None if u is None else u[b],
                           occupancy=bool(global_args.occupancy)))

    def stack(key, *shape):
        if B == 1:      # a view: stacking would copy the 0.7 GB of per-sample extras of a full frame once more
            return per[0][key].reshape(1, C, *shape)
        return torch.stack([r[key].reshape(C, *shape) for r in per], 0)

    rgb = stack("rgb_map", 3)
    acc = stack("acc_map")
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    zeros3 = torch.zeros(1, 1, 1, 1, device=dev).expand(B, C, S, 3)
    return {
        "rgb_map": rgb, "disp_map": stack("disp_map"), "acc_map": acc,
        "smpl_query_pts": stack("smpl_query_pts", S, 3), "smpl_src_pts": stack("smpl_src_pts", S, 3),
        "correction_": zeros3, "other_loss": torch.zeros(1, 4, device=dev), "correction": zeros3,
        "pts_mask": stack("pts_mask", S, 1), "raw": stack("raw", S, 4),
    }


_COPY_STREAMS = {}
# what the render path reads from the input dicts (engine._prepare_frame, lib/skinnning_batch.frame_context)
HOT_KEYS_SP = ("gender", "params", "t_vertices", "img_all", "K_all", "R_all", "T_all")
# the training loop's step counters (ref :541-542): passed through on the host, they decide which path a step takes
STEP_KEYS = ("global_step", "smooth_interval")
HOT_KEYS_TP = ("gender", "params", "vertices")


def _upload_hot(d, keys, dev):
    """Device copy of the entries of a host input dict that the path reads (asynchronous for pinned tensors);
    everything else in the dataset dict -- the target view's images, masks, indices -- stays on the host."""
    def mv(v):
        if torch.is_tensor(v):
            return v.to(dev, non_blocking=True)
        if isinstance(v, dict):
            return {k: mv(x) for k, x in v.items()}
        return v
    # gender stays on the host: it only selects which SMPL tables to use, a host-side decision; so do pinned source
    # views: the engine uploads them on its trunk stream, beside the front and K1 (engine._prepare_frame)
    # (global_step / smooth_interval likewise: the training loop's step counters decide on the host which path runs)
    keep = lambda k, v: k in ("gender", "global_step", "smooth_interval") or \
        (k == "img_all" and torch.is_tensor(v) and v.is_pinned())
    return {k: (d[k] if keep(k, d[k]) else mv(d[k])) for k in keys if k in d}


def hot_input_bytes(sp_input, tp_input):
    """Bytes render() uploads when it is handed host dicts (bench.py reports them as h2d_bytes_per_step)."""
    def nbytes(v):
        if torch.is_tensor(v):
            return v.numel() * v.element_size()
        return sum(nbytes(x) for x in v.values()) if isinstance(v, dict) else 0
    return sum(nbytes(sp_input[k]) for k in HOT_KEYS_SP if k in sp_input) + \
        sum(nbytes(tp_input[k]) for k in HOT_KEYS_TP if k in tp_input)


def _copy_stream(dev):
    s = _COPY_STREAMS.get(dev)
    if s is None:
        s = _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


def batchify_rays(rays_flat, chunk=1024 * 32, sp_input=None, tp_input=None, **kwargs):
    """ref :85-97: render in chunks of ``chunk`` rays and concatenate along the ray dim."""
    all_ret = {}
    pu = kwargs.pop("perturb_u", None)          # supplied stratified-sampling uniforms follow their rays
    sd = kwargs.pop("smooth_delta", None)       # ... and so does a supplied smooth-loss perturbation (per sample point)
    S = int(kwargs.get("N_samples", 64))
    for i in range(0, rays_flat.shape[1], chunk):
        ret = render_rays(rays_flat[:, i:i + chunk], sp_input=sp_input, tp_input=tp_input,
                          perturb_u=None if pu is None else pu[:, i:i + chunk],
                          smooth_delta=None if sd is None else sd[:, i * S:(i + chunk) * S], **kwargs)
        if "_normal_fields" in ret:
            raise NotImplementedError("a smooth-loss step goes through render() in one pass (training batches are a few "
                                      "thousand rays); chunked ray sets are an inference feature")
        for k, v in ret.items():
            all_ret.setdefault(k, []).append(v)
    return {k: torch.cat(v, 1) for k, v in all_ret.items()}


def render(H=None, W=None, focal=None, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0.0, far=1.0,
           sp_input=None, tp_input=None, use_viewdirs=False, camera=None, **kwargs):
    """ref :100-135.  rays (B,2,N,3), near/far (B,N,1) -> [rgb_map, disp_map, acc_map, extras].

    The whole ray set goes through the kernels at once (they chunk internally by active
    points); ``chunk`` only determines the width of ``extras['other_loss']`` (4 per chunk).

    ``camera`` (extension): instead of ``rays`` / ``near`` / ``far`` pass the target view itself,
    ``dict(K=, R=, T=, bounds=, H=, W=[, rows=])`` -- the arguments of the reference's ``get_rays`` /
    ``get_near_far`` (lib/if_nerf_data_utils.py:11-25, 55-92), which the datasets run in numpy per view.  The rays
    are then generated on the device (csrc/raygen.cu; one subject, B = 1), nothing per-ray is uploaded, and
    ``extras['mask_at_box']`` (N,) bool is returned as well.  ``rows`` restricts the view to a list of image rows
    (how one frame is dealt out to several GPUs, parallel.interleaved_rows).
    """
    def pack(rays, near, far):
        rays_o, rays_d = rays[:, 0, ...], rays[:, 1, ...]
        B = rays_d.shape[0]
        return torch.cat([torch.reshape(rays_o, [B, -1, 3]).float(), torch.reshape(rays_d, [B, -1, 3]).float(),
                          torch.reshape(near, [B, -1, 1]).float(), torch.reshape(far, [B, -1, 1]).float()], -1)

    ready = None
    if not sp_input["K_all"].is_cuda:
        # host dicts (ideally pinned): upload only what the path reads, on the current stream (the frame
        # preparation needs them first); the rays follow on the copy stream below
        if not torch.cuda.is_available():
            raise RuntimeError("mpsnerf_b200 has no CPU path: a CUDA device is required")
        dev = torch.device("cuda", torch.cuda.current_device())
        sp_input, tp_input = _upload_hot(sp_input, HOT_KEYS_SP + STEP_KEYS, dev), _upload_hot(tp_input, HOT_KEYS_TP, dev)
    box = None
    if rays is None:
        if camera is None:
            raise ValueError("render() needs either rays / near / far or camera=dict(K, R, T, bounds, H, W)")
        from .lib.if_nerf_data_utils import gen_rays8
        rays8, box = gen_rays8(camera["H"], camera["W"], camera["K"], camera["R"], camera["T"], camera["bounds"],
                               device=sp_input["K_all"].device, rows=camera.get("rows"))
        packed = rays8[None]
        sh = (1, packed.shape[1], 3)
    elif not rays.is_cuda and sp_input["K_all"].is_cuda:
        # Host (ideally pinned) rays / near / far: uploaded and packed on a copy stream, so that the transfer
        # overlaps the per-frame preparation (trunk, K0, grids), which needs only sp_input / tp_input.
        sh = rays[:, 1, ...].shape
        dev = sp_input["K_all"].device
        main = torch.cuda.current_stream(dev)
        cs = _copy_stream(dev)
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            packed = pack(*(t.to(dev, non_blocking=True) for t in (rays, near, far)))
            ready = torch.cuda.Event()
            ready.record(cs)
        packed.record_stream(main)
    else:
        sh = rays[:, 1, ...].shape
        packed = pack(rays, near, far)
    n = packed.shape[1]
    S_ = int(kwargs.get("N_samples", 64))
    per_ray = S_ * (44 + 20 + 8)              # extras + active list + slack, bytes per ray
    # (cudaMemGetInfo costs ~0.4 ms per call: asked only for ray sets that could plausibly not fit)
    big = n * per_ray > (8 << 30)
    free = torch.cuda.mem_get_info(packed.device)[0] if (big and packed.is_cuda) else 1 << 62
    if n * S_ >= (1 << 31) or n * per_ray > 0.6 * free:
        # The kernels take the whole ray set at once (they chunk internally by ACTIVE points); only a set whose
        # per-sample outputs cannot be held -- 2^31 sample points, or more than the free device memory -- is cut into
        # ray chunks here (the reference's `chunk` semantics), the results concatenated as batchify_rays does.
        fit = max(1024, int(0.25 * free // per_ray))
        fit = min(fit, ((1 << 31) - 1) // S_)
        if ready is not None:
            torch.cuda.current_stream(packed.device).wait_event(ready)
        ret = batchify_rays(packed, fit, sp_input=sp_input, tp_input=tp_input, **kwargs)
    else:
        ret = render_rays(packed, sp_input=sp_input, tp_input=tp_input, _rays_ready=ready, **kwargs)
    nchunks = max(1, (n + chunk - 1) // chunk)
    fields = ret.pop("_normal_fields", None)
    if fields is None:
        ret["other_loss"] = torch.zeros(1, 4 * nchunks, device=packed.device)
    else:
        # smooth step: the reference evaluates the two terms once per ray chunk (means over that chunk's sample points)
        # and concatenates the (1,4) blocks, ref :62-78 + :85-97; the training loop reads block 0
        from .smooth import smooth_losses
        cut = lambda t, i: t[:, i * chunk * S_:(i + 1) * chunk * S_]
        ret["other_loss"] = torch.cat([smooth_losses(*(cut(t, i) for t in fields)) for i in range(nchunks)], 1)
    for k in ("rgb_map", "disp_map", "acc_map", "pts_mask", "raw"):
        ret[k] = torch.reshape(ret[k], list(sh[:-1]) + list(ret[k].shape[2:]))
    if box is not None:
        ret["mask_at_box"] = box
    keys = ("rgb_map", "disp_map", "acc_map")
    return [ret[k] for k in keys] + [{k: v for k, v in ret.items() if k not in keys}]


def create_nerf(args, device=None):
    """ref :301-366 -> (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    configure(args)
    device = device or torch.device("cuda")
    model = return_model(args)
    grad_vars = list(model.parameters())
    network_query_fn = lambda inputs, viewdirs, network_fn, sp_input=None, tp_input=None: run_network(
        inputs, viewdirs, network_fn, sp_input=sp_input, tp_input=tp_input)
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    start = 0
    ckpt_dir = os.path.join(args.basedir, args.expname or "")
    if args.ft_path is not None and args.ft_path != "None":
        ckpts = [os.path.join(ckpt_dir, args.ft_path)]
    elif os.path.isdir(ckpt_dir):
        ckpts = [os.path.join(ckpt_dir, f) for f in sorted(os.listdir(ckpt_dir)) if ".tar" in f]
    else:
        ckpts = []
    if ckpts and not args.no_reload:
        ckpt = torch.load(ckpts[-1], map_location="cpu")
        start = ckpt["global_step"]
        model.load_state_dict(ckpt["network_fn_state_dict"])
    model = NetworkHandle(model).to(device)
    render_kwargs_train = {
        "network_query_fn": network_query_fn, "perturb": args.perturb, "N_samples": args.N_samples,
        "network_fn": model, "use_viewdirs": args.use_viewdirs, "N_importance": args.N_importance,
    }
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test["perturb"] = False
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
