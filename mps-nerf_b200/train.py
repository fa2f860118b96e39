"""Training step through the CUDA hot path (BASELINE config 4; reference run_nerf_batch.py:522-573).

``render()`` in training mode (``network_fn.train()`` + grad enabled) lands here: the render forward runs on the
same kernels as inference -- K1 (sample / mask / argmin), K3 (deform / project), K4 (gather), K6 (composite) -- with the
dense stage in fp32 keeping its intermediates (csrc/dense_fp32.cu: mpsnerf_dense_train_fwd), and ``loss.backward()``
runs hand-written backward kernels: K6 backward, the dense backward (weight gradients of the transformer and the
MLP, token gradients) and K4 backward (scatter-add into the NHWC latent gradient), after which torch / cuDNN
backpropagates through the encoder trunk -- the boundary of the path, as in the forward.  Under the shipped configs
no parameter sits upstream of the canonical points (skinning_field = correction_field = 0), so K1 / K3 need no
backward.  Every ``smooth_interval``-th step adds the normal-smoothness terms (occupancy normals by double backward at
the sample points and at perturbed copies of them, ref run_nerf_batch.py:60-79): K1 / K3 locate both point sets
(mask, canonical points, nearest template vertices), the twice-differentiable chain on the active points is the
torch-autograd restatement of smooth.py, and its parameter gradients join the kernels' in the same bucket before the
all-reduce.

Data parallelism (``TrainStep``): one process per GPU, replicas of the network, ONE NCCL all-reduce of the
dense-stage gradients (46 tensors, 1.08 M floats, kept in one flat bucket the backward kernels accumulate into
directly) launched from inside the backward -- it overlaps the cuDNN trunk backward that follows -- and one of the
trunk gradients after it.  torch.distributed is the plumbing (gloo in the CPU tests of the bucket logic).
"""
import ctypes

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib
from .engine import DENSE_FP32_ORDER, RenderEngine, _stream


class DenseBucket:
    """The live parameters of the transformer + MLP in DENSE_FP32_ORDER, and one flat fp32 gradient buffer with a
    view per parameter: the backward kernels accumulate into the views, the all-reduce runs on the flat buffer."""

    def __init__(self, net):
        named = dict(net.named_parameters())
        self.params = [named[k] for k in DENSE_FP32_ORDER]
        for p in self.params:
            assert p.dtype == torch.float32 and p.is_contiguous()
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64          # 256-byte aligned views
        self.flat = torch.zeros(total, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        self.wtable = (ctypes.c_void_p * len(self.params))(*[p.data_ptr() for p in self.params])
        self.gtable = (ctypes.c_void_p * len(self.params))(*[v.data_ptr() for v in self.views])
        self.pending = 0            # autograd nodes of this step that have not run their backward yet
        self.work = None            # handle of the in-flight all-reduce
        self.world = 1
        self.deferred = False       # smooth step: autograd also writes p.grad; the all-reduce waits for absorb_autograd()
        # Two ways to drive a step.  EXPLICIT (TrainStep): begin_step() ... backward ... absorb_autograd() / finish(),
        # with the all-reduce launched from inside the backward.  AUTOMATIC (the reference's own loop, unchanged:
        # render -> loss.backward() -> optimizer.step(), run_nerf_batch.py:544-563): the first render node of a step
        # opens it, and a callback queued on the autograd engine closes it when the backward pass has finished.
        self.explicit = False
        self.open = False
        self.trunk = [p for k, p in named.items() if k.startswith("encoder_2d.")]
        self._callback_queued = False

    def begin_step(self, world):
        self.flat.zero_()
        self.pending, self.work, self.world, self.deferred = 0, None, world, False
        self.explicit, self.open = True, True

    def node_forward(self):
        """Called by every render node's forward.  Opens an automatic step if no step is open: the buffer is zeroed
        unless the parameters' .grad still ARE its views (the caller did not zero them: torch's accumulate-into-.grad
        semantics, then, also for the kernels' share)."""
        if not self.open:
            kept = self.params[0].grad is not None and self.params[0].grad.data_ptr() == self.views[0].data_ptr()
            if not kept:
                self.flat.zero_()
            self.pending, self.work, self.deferred = 0, None, False
            self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            self.explicit, self.open, self._callback_queued = False, True, False
        self.pending += 1

    def node_backward_begin(self):
        """Called at the start of every render node's backward: in an automatic step, have the autograd engine call
        auto_finalize() once this backward pass is complete (the hook DistributedDataParallel finalises with)."""
        if not self.explicit and not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self.auto_finalize)

    def node_done(self):
        """Called at the end of every render node's backward; in an explicit step the last one launches the all-reduce."""
        self.pending -= 1
        if self.explicit and self.pending == 0 and self.world > 1 and not self.deferred:
            self.work = dist.all_reduce(self.flat, async_op=True)

    def absorb_autograd(self):
        """Smooth step: the torch-autograd chain of smooth.py accumulated its share of the dense-stage gradients in
        ``p.grad``; add them to the kernels' share in the flat buffer, then launch the all-reduce that was held back."""
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.add_(p.grad)
                p.grad = None
        if self.deferred and self.world > 1 and self.work is None:
            self.work = dist.all_reduce(self.flat, async_op=True)

    def finish(self):
        """Wait for the all-reduce, average, hand the views to the parameters as .grad."""
        if self.work is not None:
            self.work.wait()
            self.work = None
        if self.world > 1:
            self.flat.div_(self.world)
        for p, v in zip(self.params, self.views):
            p.grad = v
        self.explicit, self.open = False, False

    def auto_finalize(self):
        """End of the backward pass of an automatic step: merge what autograd itself accumulated for the dense
        parameters (smooth steps), average over the ranks (dense bucket and encoder-trunk gradients), and leave every
        gradient in ``.grad`` -- ``optimizer.step()`` can follow, as in the reference's loop."""
        self._callback_queued = False
        if self.explicit or not self.open:
            return
        self.deferred = True
        self.absorb_autograd()
        if self.world > 1:
            if self.work is None:
                self.work = dist.all_reduce(self.flat, async_op=True)
            allreduce_mean_(self.trunk, self.world)
        self.finish()


def allreduce_mean_(params, world):
    """Average the gradients of ``params`` over the ranks in one flat all-reduce (the encoder trunk's, after cuDNN)."""
    ps = [p for p in params if p.grad is not None]
    if world <= 1 or not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat)
    flat.div_(world)
    o = 0
    for p in ps:
        p.grad.copy_(flat[o:o + p.numel()].view_as(p))
        o += p.numel()


class _RenderNode(torch.autograd.Function):
    """One subject's render: latent (V, Hf, Wf, 128) -> rgb_map, acc_map (+ the non-differentiable extras)."""

    @staticmethod
    def forward(ctx, latent, te, fctx, img4, rays8, S, t_vals, u, occupancy, located):
        eng, lib, dev = te.eng, te.eng.lib, latent.device
        N = rays8.shape[0]
        P, V = N * S, fctx.n_views
        raw = torch.empty(P, 4, device=dev)
        mask = torch.empty(P, device=dev)
        sq = torch.empty(P, 3, device=dev)
        ss = torch.empty(P, 3, device=dev)
        act_pid = torch.empty(P, dtype=torch.int32, device=dev)
        act_idx2 = torch.empty(P, dtype=torch.int32, device=dev)
        act_q = torch.empty(P, 3, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.mpsnerf_sample_knn(_lib.ptr(rays8), N, S, _lib.ptr(t_vals), _lib.ptr(u), None, _lib.ptr(fctx.frame_dev),
                                          _lib.ptr(fctx.grid_tp), _lib.ptr(raw), _lib.ptr(mask), _lib.ptr(sq), _lib.ptr(ss),
                                          _lib.ptr(act_pid), _lib.ptr(act_idx2), _lib.ptr(act_q), _lib.ptr(counter), _stream()),
                   "sample_knn")
        n = int(counter.item())
        act_pid = act_pid[:n].clone()
        xc = torch.empty(max(n, 1), 3, device=dev)
        uv = torch.empty(max(n, 1), V, 2, device=dev)
        tokens = torch.empty(max(n, 1), V, _lib.TOKEN_DIM, device=dev)
        out4 = torch.empty(max(n, 1), 4, device=dev)
        ws = torch.empty(max(lib.mpsnerf_dense_train_workspace(n, V), 256), dtype=torch.uint8, device=dev)
        idx3 = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if located is not None else None
        if n:
            fctx.wait_lbs()
            _lib.check(lib.mpsnerf_deform_project(_lib.ptr(act_pid), _lib.ptr(act_idx2), _lib.ptr(act_q), 0, n, _lib.ptr(fctx.skin_w),
                                                  _lib.ptr(fctx.frame_dev), _lib.ptr(fctx.grid_tv), _lib.ptr(xc), _lib.ptr(uv),
                                                  _lib.ptr(ss), _lib.ptr(idx3), None, 0, _stream()), "deform_project")
            _lib.check(lib.mpsnerf_gather_tokens(_lib.ptr(uv), n, V, _lib.ptr(fctx.frame_dev), _lib.ptr(latent), _lib.ptr(img4),
                                                 _lib.ptr(tokens), _lib.TOKEN_DIM, _stream()), "gather_tokens")
            _lib.check(lib.mpsnerf_dense_train_fwd(_lib.ptr(tokens), _lib.TOKEN_DIM, _lib.ptr(xc), n, V, te.bucket.wtable,
                                                   _lib.ptr(out4), _lib.ptr(ws), _stream()), "dense_train_fwd")
            _lib.check(lib.mpsnerf_rows4_scatter(_lib.ptr(out4), _lib.ptr(act_pid), n, _lib.ptr(raw), _stream()), "rows4_scatter")
        rgb = torch.empty(N, 3, device=dev)
        disp = torch.empty(N, device=dev)
        acc = torch.empty(N, device=dev)
        _lib.check(lib.mpsnerf_composite(_lib.ptr(raw), _lib.ptr(rays8), N, S, _lib.ptr(t_vals), _lib.ptr(u), None,
                                         1 if occupancy else 0, _lib.ptr(rgb), _lib.ptr(disp), _lib.ptr(acc), None, None, None,
                                         _stream()), "composite")
        _lib.count_launches(6 + 44)
        if located is not None:     # smooth step: the index stages of this pass, for smooth.normal_fields
            located.update(act_pid=act_pid, xc=xc[:n], idx3=idx3[:n])
        eng.last_active = n
        ctx.te, ctx.fctx, ctx.n, ctx.S, ctx.occupancy = te, fctx, n, S, occupancy
        ctx.keep = (rays8, t_vals, u, raw, act_pid, uv, ws, latent.shape)
        ctx.mark_non_differentiable(disp, raw, mask, sq, ss)
        te.bucket.node_forward()
        return rgb, acc, disp, raw, mask, sq, ss

    @staticmethod
    def backward(ctx, d_rgb, d_acc, *_):
        te, fctx, n, S = ctx.te, ctx.fctx, ctx.n, ctx.S
        lib = te.eng.lib
        te.bucket.node_backward_begin()
        rays8, t_vals, u, raw, act_pid, uv, ws, lshape = ctx.keep
        dev = raw.device
        N, V = rays8.shape[0], fctx.n_views
        d_latent = torch.zeros(lshape, device=dev)
        if n:
            d_rgb = (torch.zeros(N, 3, device=dev) if d_rgb is None else d_rgb).float().contiguous()
            d_acc = None if d_acc is None else d_acc.float().contiguous()
            d_raw = torch.empty(N * S, 4, device=dev)
            _lib.check(lib.mpsnerf_composite_bwd(_lib.ptr(raw), _lib.ptr(rays8), N, S, _lib.ptr(t_vals), _lib.ptr(u), None,
                                                 1 if ctx.occupancy else 0, _lib.ptr(d_rgb), _lib.ptr(d_acc), _lib.ptr(d_raw),
                                                 _stream()), "composite_bwd")
            d_out4 = torch.empty(n, 4, device=dev)
            _lib.check(lib.mpsnerf_rows4_gather(_lib.ptr(d_raw), _lib.ptr(act_pid), n, _lib.ptr(d_out4), _stream()), "rows4_gather")
            d_tokens = torch.empty(n, V, _lib.TOKEN_DIM, device=dev)
            _lib.check(lib.mpsnerf_dense_train_bwd(_lib.ptr(d_out4), n, V, te.bucket.wtable, te.bucket.gtable, _lib.ptr(d_tokens),
                                                   _lib.ptr(ws), _stream()), "dense_train_bwd")
            _lib.check(lib.mpsnerf_gather_tokens_bwd(_lib.ptr(uv), n, V, _lib.ptr(fctx.frame_dev), _lib.ptr(d_tokens),
                                                     _lib.TOKEN_DIM, _lib.ptr(d_latent), _stream()), "gather_tokens_bwd")
            _lib.count_launches(3 + 80)
        te.bucket.node_done()
        return (d_latent,) + (None,) * 9


class TrainEngine:
    """Per-network state of the training path: an fp32 RenderEngine for K0 / grids / K1 / K3 / K4 / K6 and the
    gradient bucket of the dense stage."""

    def __init__(self, net):
        self.net = net
        self.eng = RenderEngine(net, precision="fp32")
        self.bucket = DenseBucket(net)

    def render_subject(self, sp, tp, rays8, S, t_vals, u, occupancy, smooth_pts=None):
        """sp / tp: one subject's (squeezed) dicts on the device.  -> rgb (N,3), acc (N), disp, raw, mask, sq, ss and,
        on a smooth step (``smooth_pts`` = the perturbed sample points (N*S,3)), the three normal fields
        (occ0, smpl0, occ1), each (N*S,3); None otherwise."""
        net = self.net
        net._check_supported()
        fctx = self.eng.prepare_frame(sp, tp, net._smpl_for(sp["gender"]), trunk=False)
        img = sp["img_all"].to(tp["vertices"].device, non_blocking=True).float()
        latent_nchw = net.encoder_2d(img)                          # torch / cuDNN under autograd: the path's boundary
        latent = latent_nchw.permute(0, 2, 3, 1).contiguous().float()
        img4 = F.pad(img.permute(0, 2, 3, 1), (0, 1)).contiguous()
        located = {} if smooth_pts is not None else None
        out = _RenderNode.apply(latent, self, fctx, img4, rays8, S, t_vals, u, occupancy, located)
        if smooth_pts is None:
            return out, None
        from . import smooth
        P = rays8.shape[0] * S
        # (copies: the frame buffer belongs to the engine and is rewritten by the next subject's preparation, while
        # autograd keeps these constants for the backward)
        fctx.wait_lbs()
        fr = {k: v.clone() for k, v in smooth.frame_constants(fctx.frame_dev, fctx.n_views).items()}
        normals = smooth.vertex_normals(sp["t_vertices"].float(), net.faces.to(img.device))
        second = self.locate(fctx, smooth_pts)
        occ0, smpl0, occ1 = smooth.normal_fields(net, fr, latent_nchw, img, fctx.skin_w, normals, P, located, second)
        return out, (occ0, smpl0, occ1)

    def locate(self, fctx, points):
        """The index stages of a network pass on explicit world points (the perturbed pass of a smooth step): K1 in
        its points mode (mask, nearest posed vertex, compaction) and K3 (canonical point, nearest template vertex).
        -> dict(act_pid (n) int32, xc (n,3), idx3 (n) int32)."""
        lib, dev = self.eng.lib, points.device
        points = points.float().contiguous()
        P, V = points.shape[0], fctx.n_views
        raw = torch.empty(P, 4, device=dev)
        mask = torch.empty(P, device=dev)
        sq = torch.empty(P, 3, device=dev)
        ss = torch.empty(P, 3, device=dev)
        act_pid = torch.empty(P, dtype=torch.int32, device=dev)
        act_idx2 = torch.empty(P, dtype=torch.int32, device=dev)
        act_q = torch.empty(P, 3, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.mpsnerf_sample_knn(None, P, 1, None, None, _lib.ptr(points), _lib.ptr(fctx.frame_dev),
                                          _lib.ptr(fctx.grid_tp), _lib.ptr(raw), _lib.ptr(mask), _lib.ptr(sq), _lib.ptr(ss),
                                          _lib.ptr(act_pid), _lib.ptr(act_idx2), _lib.ptr(act_q), _lib.ptr(counter), _stream()),
                   "sample_knn")
        n = int(counter.item())
        act_pid = act_pid[:n].clone()
        xc = torch.empty(max(n, 1), 3, device=dev)
        uv = torch.empty(max(n, 1), V, 2, device=dev)
        idx3 = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        if n:
            fctx.wait_lbs()
            _lib.check(lib.mpsnerf_deform_project(_lib.ptr(act_pid), _lib.ptr(act_idx2), _lib.ptr(act_q), 0, n, _lib.ptr(fctx.skin_w),
                                                  _lib.ptr(fctx.frame_dev), _lib.ptr(fctx.grid_tv), _lib.ptr(xc), _lib.ptr(uv),
                                                  _lib.ptr(ss), _lib.ptr(idx3), None, 0, _stream()), "deform_project")
        _lib.count_launches(2)
        return {"act_pid": act_pid, "xc": xc[:n], "idx3": idx3[:n]}


def render_rays_train(net, ray_batch, sp_input, tp_input, N_samples, perturb, perturb_u, white_bkgd, occupancy, select,
                      smooth_step=False, smooth_delta=None):
    """Training-mode body of run_nerf_batch.render_rays: same dict of outputs, rgb_map / acc_map carry gradients.
    ``smooth_step``: also evaluate the normal-smoothness terms into ``other_loss`` (ref run_nerf_batch.py:60-79);
    ``smooth_delta`` (B, C*S, 3) supplies the perturbation of the second pass (N(0, 0.01) otherwise, :36, :64)."""
    te = net.train_engine()
    B, C = ray_batch.shape[:2]
    dev = ray_batch.device
    S = int(N_samples)
    t_vals = torch.linspace(0.0, 1.0, steps=S, device=dev)
    u = None
    if perturb > 0.0:
        u = (perturb_u if perturb_u is not None else torch.rand(B, C, S, device=dev)).float().contiguous()
    pts1 = None
    if smooth_step:
        from . import smooth
        te.bucket.deferred = True           # autograd will add to p.grad: TrainStep merges before the all-reduce
        delta = smooth_delta.to(dev).float() if smooth_delta is not None else 0.01 * torch.randn(B, C * S, 3, device=dev)
        pts1 = smooth.sample_points(ray_batch[..., :8].float(), t_vals, u).reshape(B, C * S, 3) + delta.reshape(B, C * S, 3)
    per, fields = [], []
    for b in range(B):
        sp, tp = select(sp_input, b), select(tp_input, b)
        out, nf = te.render_subject(sp, tp, ray_batch[b, :, :8].float().contiguous(), S, t_vals,
                                    None if u is None else u[b].contiguous(), occupancy,
                                    smooth_pts=None if pts1 is None else pts1[b])
        per.append(out)
        fields.append(nf)
    st = lambda i, *shape: torch.stack([r[i].reshape(C, *shape) for r in per], 0)
    rgb, acc = st(0, 3), st(1)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    out = {"rgb_map": rgb, "disp_map": st(2), "acc_map": acc, "smpl_query_pts": st(5, S, 3), "smpl_src_pts": st(6, S, 3),
           "other_loss": torch.zeros(1, 4, device=dev), "pts_mask": st(4, S, 1), "raw": st(3, S, 4)}
    out["correction"] = out["correction_"] = torch.zeros(1, 1, 1, 1, device=dev).expand(B, C, S, 3)
    if smooth_step:
        occ0, smpl0, occ1 = (torch.stack([f[i] for f in fields], 0) for i in range(3))      # (B, C*S, 3) each
        out["other_loss"] = smooth.smooth_losses(occ0, smpl0, occ1)
        out["_normal_fields"] = (occ0, smpl0, occ1)      # render() re-cuts the terms per ray chunk, as the reference does
    return out


def img2mse(x, y):
    return torch.mean((x - y) ** 2)


class TrainStep:
    """One optimisation step of the reference's loop (run_nerf_batch.py:544-570) with data-parallel gradient averaging.

    ``step(...)``: render in training mode -> img2mse(rgb, target) [+ img2mse(bkgd_msk, acc) when ``acc_loss``] ->
    backward (dense-stage all-reduce launched inside, overlapping the trunk backward) -> trunk all-reduce ->
    optimizer.step().  Returns the loss (a device scalar; no host sync is forced)."""

    def __init__(self, network_fn, optimizer, acc_loss=True):
        self.handle = network_fn
        self.net = network_fn.module if hasattr(network_fn, "module") else network_fn
        self.optimizer = optimizer
        self.acc_loss = acc_loss
        self.trunk_flat = None

    def step(self, render, rays, near, far, sp_input, tp_input, target_rgb, bkgd_msk=None, **render_kw):
        world = dist.get_world_size() if dist.is_initialized() else 1
        bucket = self.net.train_engine().bucket
        self.optimizer.zero_grad(set_to_none=True)
        bucket.begin_step(world)
        rgb, _, acc, extras = render(rays=rays, near=near, far=far, sp_input=sp_input, tp_input=tp_input,
                                     network_fn=self.handle, **render_kw)
        loss = img2mse(rgb, target_rgb)
        if self.acc_loss and bkgd_msk is not None:
            loss = loss + img2mse(bkgd_msk.squeeze(2), acc)
        loss = loss + extras["other_loss"][0][0]          # the smooth terms on interval steps, 0 otherwise (:558)
        loss.backward()
        bucket.absorb_autograd()
        self.allreduce_trunk(world)
        bucket.finish()
        self.optimizer.step()
        return loss.detach()

    def allreduce_trunk(self, world):
        """Average the encoder-trunk gradients over the ranks in one flat all-reduce."""
        allreduce_mean_(list(self.net.encoder_2d.parameters()), world)
