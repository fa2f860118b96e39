"""Render engine: per-frame preparation and kernel orchestration over the C ABI.

This is the host side of the hot path.  Per ``render()`` call it
  1. computes the per-frame constants (LBS transforms for target / big / source pose,
     cameras) once -- the reference recomputes them 4x per chunk
     (lib/skinnning_batch.py:206,243,266,289);
  2. runs the encoder trunk once and lays its output out NHWC -- the reference re-runs it
     per chunk (lib/skinnning_batch.py:350-351);
  3. builds the two exact nearest-vertex grids (posed vertices in SMPL space, template)
     -- steps 1-3 as three CUDA-graph branches on three streams, K1 starting behind the first;
  4. enqueues K1 (sample + mask + argmin + compaction) over all rays, K3 / K4 / K5 for a slab of active points whose
     count is read on the device, and K6 -- one C call (mpsnerf_render_rays_bf16) or stage by stage -- and looks at
     the active count only afterwards (overflow beyond the slab: remainder slabs, K6 again).
All per-point arithmetic happens in libmpsnerf_b200.so; torch is used for memory, streams
and the cuDNN encoder trunk only.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

GRID_CELL_TARGET = 0.0505      # >= 1.01 * 0.05 m: makes the 27-cell search exact for the mask radius
GRID_CELL_TEMPLATE = 0.06

# order of the host pointer table taken by mpsnerf_dense_fp32
DENSE_FP32_ORDER = (
    [f"transformer.layers.{l}.{k}" for l in range(2) for k in (
        "0.fn.norm.weight", "0.fn.norm.bias", "0.fn.fn.to_qkv.weight", "0.fn.fn.to_out.0.weight",
        "0.fn.fn.to_out.0.bias", "1.fn.norm.weight", "1.fn.norm.bias", "1.fn.fn.net.0.weight",
        "1.fn.fn.net.0.bias", "1.fn.fn.net.3.weight", "1.fn.fn.net.3.bias")]
    + [f"pts_linears.{i}.{k}" for i in range(8) for k in ("weight", "bias")]
    + ["alpha_linear.weight", "alpha_linear.bias", "feature_linear.weight", "feature_linear.bias",
       "views_linear.weight", "views_linear.bias", "rgb_linear.weight", "rgb_linear.bias"])


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_kmajor_sw128(w, n_pad=None, k_pad=None, dtype=torch.bfloat16):
    """Pack a (N,K) weight into the K-major SWIZZLE_128B chunk layout of csrc/umma.cuh.

    Returns a uint8 tensor of (K_pad/64) chunks x N_pad rows x 128 bytes: bf16, the 16-byte
    unit j of row r stored at unit (j ^ (r & 7)) of that row.  Pure index arithmetic, so it
    runs on any device; checked on CPU against the address formula in tests/test_packing.py.
    """
    N, K = w.shape
    n_pad = n_pad or ((N + 15) // 16) * 16
    k_pad = k_pad or ((K + 63) // 64) * 64
    wp = torch.zeros(n_pad, k_pad, dtype=dtype, device=w.device)      # bf16, or fp16 (the out-projection, see pack.py)
    wp[:N, :K] = w.to(dtype)
    t = wp.reshape(n_pad, k_pad // 64, 8, 8)                       # (row, chunk, unit, elem)
    r7 = (torch.arange(n_pad, device=w.device) & 7)[:, None]
    src_unit = torch.arange(8, device=w.device)[None, :] ^ r7      # stored unit u holds logical unit u ^ r7
    t = torch.gather(t, 2, src_unit[:, None, :, None].expand(n_pad, k_pad // 64, 8, 8))
    return t.permute(1, 0, 2, 3).contiguous().view(torch.uint8).reshape(-1)


_ROW_INDEX = {}


def assemble_latent_rows(gathered, Hf):
    """(world, V, mr, Wf, C) bands of latent rows, rank-major and padded to mr = ceil(Hf / world) rows each
    (parallel.ray_block split) -> (V, Hf, Wf, C)."""
    from .parallel import ray_block
    world, V, mr, Wf, C = gathered.shape
    x = gathered.permute(1, 0, 2, 3, 4).reshape(V, world * mr, Wf, C)
    if Hf == world * mr:
        return x.contiguous()
    key = (Hf, world, str(gathered.device))
    idx = _ROW_INDEX.get(key)
    if idx is None:
        idx = torch.cat([torch.arange(r * mr, r * mr + (e - s)) for r in range(world) for s, e in [ray_block(Hf, r, world)]])
        idx = _ROW_INDEX[key] = idx.to(gathered.device)
    return x.index_select(1, idx)


class FrameContext:
    """Device-resident state of one (source, target) pair.

    The preparation runs as three concurrent branches (engine._prepare_frame): the *front* (frame header + target-pose
    grid: all K1 needs) on the caller's stream, the *LBS* branch (transform sets + template grid: needed from K3 on)
    and the *trunk* branch (encoder latent + NHWC images: needed from K4 on) on side streams.  ``wait_lbs`` /
    ``wait_trunk`` make the current stream wait for a branch; reading ``latent`` / ``img4`` waits for the trunk."""
    __slots__ = ("frame_dev", "grid_tp", "grid_tv", "_latent", "_img4", "skin_w", "n_views", "keep", "ev_lbs", "ev_trunk",
                 "pending_trunk")

    def launch_pending(self):
        """Enqueue a trunk branch whose launch was deferred (sharded trunk: ~15 eager launches + a collective, whose
        host time would otherwise sit between the front and K1): the engine calls this right behind K1."""
        if getattr(self, "pending_trunk", None) is not None:
            fn, self.pending_trunk = self.pending_trunk, None
            fn()

    def wait_lbs(self):
        if self.ev_lbs is not None:
            torch.cuda.current_stream().wait_event(self.ev_lbs)

    def wait_trunk(self):
        self.launch_pending()
        if self.ev_trunk is not None:
            torch.cuda.current_stream().wait_event(self.ev_trunk)

    @property
    def latent(self):
        self.wait_trunk()
        return self._latent

    @property
    def img4(self):
        self.wait_trunk()
        return self._img4

    def frame_host(self):
        """Host copy of the device-side mpsnerf_frame (tests / debugging; synchronises)."""
        self.wait_lbs()
        return _lib.Frame.from_buffer_copy(bytes(self.frame_dev.cpu().numpy()))


class RenderEngine:
    def __init__(self, net, precision="fp32", slab=None):
        assert precision in ("fp32", "bf16")
        self.net = net
        self.precision = precision
        self.slab = slab or (32768 if precision == "fp32" else 1 << 21)
        self.lib = _lib.load()
        self._ws = {}
        self._packed = None
        self._packed_key = None
        self._dense_params = None
        self._trunk_params = None
        self._smpl_cache = {}
        self._trunk = None
        self._pinned = {}
        self._count_events = {}
        self.trunk_shard = None
        self._use_device_count = os.environ.get("MPSNERF_DEVICE_COUNT", "1") != "0"
        self._use_fused = os.environ.get("MPSNERF_FUSED_CALL", "1") != "0"
        self._side = {}
        self._prep_graphs = {}
        self._use_prep_graph = os.environ.get("MPSNERF_PREP_GRAPH", "1") != "0"
        self.debug = None            # set to a dict to capture per-stage tensors (tests)
        self.timers = None           # set to a dict to record CUDA-event pairs per stage (bench.py)
        self.last_active = 0

    class _Span:
        def __init__(self, eng, name):
            self.eng, self.name = eng, name

        def __enter__(self):
            if self.eng.timers is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *a):
            if self.eng.timers is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.eng.timers.setdefault(self.name, []).append((self.e0, e1))

    def span(self, name):
        return RenderEngine._Span(self, name)

    # ------------------------------------------------------------------ helpers
    def _buf(self, name, nbytes, device):
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes or t.device != device:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._ws[name] = t
        return t

    def _smpl_tables(self, smpl, dev):
        """Device copies of the SMPL tables of one gender (cached for the life of the engine)."""
        key = (id(smpl), str(dev))
        tab = self._smpl_cache.get(key)
        if tab is None:
            tab = {
                "v_template": smpl["v_template"].to(dev).float().contiguous(),
                "shapedirs": smpl["shapedirs"].to(dev).float().contiguous(),
                "J_regressor": smpl["J_regressor"].to(dev).float().contiguous(),
                "parents": smpl["kintree_table"][0].to(dev).to(torch.int32).contiguous(),
                "weights": smpl["weights"].to(dev).float().contiguous(),
            }
            self._smpl_cache[key] = tab
        return tab

    # ------------------------------------------------------------------ encoder trunk (boundary: cuDNN)
    def _trunk_ok(self, img):
        enc = self.net.encoder_2d
        m = enc.model
        return (enc.num_layers == 2 and not enc.use_first_pool and enc.feature_scale == 0.5 and not m.training
                and img.shape[-1] % 2 == 0 and img.shape[-2] % 2 == 0
                and all(type(b).__name__ == "BasicBlock" and b.downsample is None for b in m.layer1))

    def _folded_trunk(self, dev):
        """conv1 + layer1 with BatchNorm folded in (eval mode), channels_last, cached per parameter version."""
        m = self.net.encoder_2d.model
        key = tuple(t._version for t in self._trunk_state())
        if self._trunk is None or self._trunk[0] != key or self._trunk[1][0][0].device != dev:
            def fold(conv, bn):
                scale = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
                w = (conv.weight.double() * scale[:, None, None, None]).float().contiguous(memory_format=torch.channels_last)
                b = (bn.bias.double() - bn.running_mean.double() * scale).float().contiguous()
                return w, b
            layers = [fold(m.conv1, m.bn1)] + [f for blk in m.layer1 for f in (fold(blk.conv1, blk.bn1), fold(blk.conv2, blk.bn2))]
            self._trunk = (key, layers)
        return self._trunk[1]

    @staticmethod
    def _trunk_convs(x, L, nblocks):
        """half-resolution image (channels_last) -> (conv1 output, layer1 output)"""
        x = torch.cudnn_convolution_relu(x, L[0][0], L[0][1], (2, 2), (3, 3), (1, 1), 1)
        lat0 = x
        for i in range(nblocks):
            w1, b1 = L[1 + 2 * i]
            w2, b2 = L[2 + 2 * i]
            y = torch.cudnn_convolution_relu(x, w1, b1, (1, 1), (1, 1), (1, 1), 1)
            x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, (1, 1), (1, 1), (1, 1), 1)
        return lat0, x

    @torch.no_grad()
    def _encode(self, img):
        """SpatialEncoder.forward (lib/encoder.py:256-306) for the shipped trunk shape (num_layers = 2, no first
        pool, feature_scale = 0.5, eval mode), with BatchNorm folded into the convolutions and cuDNN's fused
        conv+bias+ReLU / conv+bias+add+ReLU kernels: ~9 launches instead of ~35.  Any other trunk configuration
        falls back to the module itself.  The folded weights are cached per parameter version."""
        if not self._trunk_ok(img):
            return self.net.encoder_2d(img)
        L = self._folded_trunk(img.device)
        # channels_last end to end: cuDNN's TF32 kernels are NHWC, and the NHWC view the gather wants is then free
        x = F.avg_pool2d(img.float(), 2).contiguous(memory_format=torch.channels_last)   # == area interpolation at scale 0.5 for even sizes
        lat0, x = self._trunk_convs(x, L, len(self.net.encoder_2d.model.layer1))
        return torch.cat([lat0, x], dim=1)

    @torch.no_grad()
    def _encode_rows(self, img, a, b):
        """Latent rows [a, b) only (NCHW, (V, 128, b - a, Wf)): the trunk on a horizontal band of the images.

        One frame split over N GPUs replicates the per-frame preparation; the trunk is its only large piece (0.6 ms of a
        2.6 ms frame at 3 x 1000 x 1000 on 8 GPUs), so each rank computes 1/N of the latent rows and the bands are
        all-gathered over NVLink (_prep_trunk_sharded).  A latent row depends on a finite band of input rows: conv1
        (7 x 7, stride 2, pad 3) output row y reads half-resolution rows 2y - 3 .. 2y + 3, and each of the six 3 x 3
        convolutions of layer1 widens the dependency by one latent row -- so the band is computed with a 6-row halo on
        each interior side (zero padding is only right at the true image border; at a cut it corrupts one more row per
        convolution, exactly the halo, which is cropped).  Same kernels and weights as _encode; results differ from the
        unsplit trunk only by cuDNN's choice of tiling per size (TF32 / fp32 re-association, ~1e-6)."""
        L = self._folded_trunk(img.device)
        nblk = len(self.net.encoder_2d.model.layer1)
        H = img.shape[-2]
        Hh = H // 2
        Hf = (Hh - 1) // 2 + 1
        halo = 2 * nblk
        la, lb = max(0, a - halo), min(Hf, b + halo)               # conv1 rows needed, incl. the layer1 halo
        r0 = max(0, 2 * la - 4)                                     # even: keeps the stride-2 phase of conv1
        r1 = min(Hh, 2 * (lb - 1) + 3 + 1)
        x = F.avg_pool2d(img[:, :, 2 * r0:2 * r1].float(), 2).contiguous(memory_format=torch.channels_last)
        x = torch.cudnn_convolution_relu(x, L[0][0], L[0][1], (2, 2), (3, 3), (1, 1), 1)
        off = la - r0 // 2                                          # absolute conv1 row of x[..., j, :] is r0 / 2 + j
        x = x[:, :, off:off + (lb - la)]
        lat0 = x
        for i in range(nblk):
            w1, b1 = L[1 + 2 * i]
            w2, b2 = L[2 + 2 * i]
            y = torch.cudnn_convolution_relu(x, w1, b1, (1, 1), (1, 1), (1, 1), 1)
            x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, (1, 1), (1, 1), (1, 1), 1)
        sl = slice(a - la, a - la + (b - a))
        return torch.cat([lat0[:, :, sl], x[:, :, sl]], dim=1)

    def _weights_fp32(self):
        sd = dict(self.net.named_parameters())
        tensors = [sd[k].detach() for k in DENSE_FP32_ORDER]
        for t in tensors:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        table = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return table, tensors

    # ------------------------------------------------------------------ per-frame preparation
    @torch.no_grad()
    def prepare_frame(self, sp, tp, smpl, trunk=True, n_points=None):
        """sp/tp: squeezed input dicts on the CUDA device; smpl: CPU tensor dict of the gender.
        ``trunk=False`` skips the encoder branch (the training step runs the trunk under autograd itself);
        ``n_points`` = sample points the frame will push through K1 (a hint for the branch order, see below)."""
        with self.span("prep"):
            return self._prepare_frame(sp, tp, smpl, trunk, n_points)

    def _prepare_frame(self, sp, tp, smpl, trunk=True, n_points=None):
        """~45 small launches (trunk, layout changes, K0, two grid builds) whose GPU time is ~0.45 ms but whose
        launch overhead is ~1 ms when issued one by one behind an idle GPU: they are captured once per
        (shape, device, weights) into THREE CUDA graphs over static input / output buffers -- front (frame header,
        target-pose grid), LBS (transform sets, template grid), trunk (encoder, NHWC layouts) -- and replayed per
        frame on three streams, so that K1 starts right behind the front and runs beside the trunk; K3 waits for the
        LBS branch and K4 for the trunk (FrameContext.wait_*).
        A FrameContext therefore stays valid until the next prepare_frame of this engine (the network keeps a
        single-entry frame cache, lib/skinnning_batch.py::frame_context).  MPSNERF_PREP_GRAPH=0 disables the graphs
        (same three branches, launched eagerly)."""
        # img_all may stay in PINNED HOST memory: it is then uploaded on the trunk stream, straight into the trunk's
        # static input, beside the front and K1 (which do not read it) -- the 9 MB of source views are the bulk of what an
        # end-to-end render() call has to bring over PCIe
        dev = tp["vertices"].device
        host_img = not sp["img_all"].is_cuda
        V = sp["img_all"].shape[0]
        vmax = _lib.MAX_VIEWS if self.precision == "fp32" else _lib.MAX_VIEWS_TC
        if not 2 <= V <= vmax:      # checked before anything is enqueued (the dense kernels would fail late and vaguely)
            raise ValueError(f"{V} input views: the {self.precision} path supports 2..{vmax} "
                             f"(tensor-core path 2..{_lib.MAX_VIEWS_TC}, fp32 path 2..{_lib.MAX_VIEWS})")
        f32 = lambda t: t.reshape(-1).float().contiguous()
        ins = [sp["img_all"].float().contiguous()] + \
              [f32(tp["params"][k]) for k in ("poses", "shapes", "R", "Th")] + \
              [f32(sp["params"][k]) for k in ("poses", "shapes", "R", "Th")] + \
              [f32(sp["R_all"]), f32(sp["T_all"]), f32(sp["K_all"]),
               tp["vertices"].float().contiguous(), sp["t_vertices"].float().contiguous()]
        main = torch.cuda.current_stream()
        sides = self._side.get(dev)
        if sides is None:
            sides = self._side[dev] = tuple(torch.cuda.Stream(device=dev) for _ in range(4))
        s_lbs, s_trunk = sides[0], sides[1]
        # a previous frame's branches may still be reading the static inputs / writing the static outputs
        main.wait_stream(s_lbs)
        main.wait_stream(s_trunk)
        if not self._use_prep_graph:
            if host_img:
                ins[0] = ins[0].to(dev, non_blocking=True)
            ctx = self._new_ctx(ins, smpl)
            self._prep_front(ins, ctx)
            for st in (s_lbs, s_trunk):          # behind the front: its single-CTA grid build runs alone (3x faster)
                st.wait_stream(main)
            with torch.cuda.stream(s_lbs):
                self._prep_lbs(ins, ctx, sides[2])
                ctx.ev_lbs = torch.cuda.Event()
                ctx.ev_lbs.record()
            if trunk:
                with torch.cuda.stream(s_trunk):
                    (self._prep_trunk if self.trunk_shard is None else self._prep_trunk_sharded)(ins, ctx)
                    ctx._latent.record_stream(main)
                    ctx._img4.record_stream(main)
                    ctx.ev_trunk = torch.cuda.Event()
                    ctx.ev_trunk.record()
            _lib.count_launches(4)
            return ctx
        m = self.net.encoder_2d.model
        key = (str(dev), self.precision, id(smpl), tuple(tuple(t.shape) for t in ins), bool(trunk),
               tuple(t._version for t in self._trunk_state()) if trunk else ())
        g = self._prep_graphs.get(key)
        if g is None:
            try:
                static = [torch.empty_like(t, device=dev) for t in ins]
                torch._foreach_copy_(static, ins)
                ctx = self._new_ctx(static, smpl)
                warm = torch.cuda.Stream(device=dev)
                warm.wait_stream(main)
                bodies = [lambda: self._prep_front(static, ctx), lambda: self._prep_lbs(static, ctx, sides[2])]
                if trunk:               # (the training step runs the trunk itself, under autograd, in training mode)
                    bodies.append(lambda: self._prep_trunk(static, ctx))
                with torch.cuda.stream(warm):          # warm-up outside the capture (cuDNN plans, lazy inits)
                    for body in bodies:
                        body()
                main.wait_stream(warm)
                torch.cuda.synchronize(dev)
                graphs = []
                for body in bodies:
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr):
                        body()
                    graphs.append(gr)
                if len(self._prep_graphs) >= 4:
                    self._prep_graphs.clear()
                g = self._prep_graphs[key] = (graphs, static, ctx)
            except RuntimeError as e:
                # Only a capture limitation (an op that cannot be captured on this build / allocator state) sends
                # the engine to the eager path, loudly; anything else -- bad inputs, OOM, a cuDNN failure -- is a
                # real error and propagates.
                msg = str(e)
                if "captur" not in msg.lower() or "out of memory" in msg.lower():
                    raise
                import warnings
                warnings.warn(f"mpsnerf_b200: CUDA-graph capture of the frame preparation failed ({msg.splitlines()[0]}); "
                              "running it eagerly from now on (MPSNERF_PREP_GRAPH=0 silences this)")
                self._use_prep_graph = False
                torch.cuda.synchronize(dev)
                return self._prepare_frame(sp, tp, smpl, trunk, n_points)
        graphs, static, ctx = g
        if host_img:
            torch._foreach_copy_(static[1:], ins[1:])
        else:
            torch._foreach_copy_(static, ins)
        # The side branches normally start BEHIND the front: its single-CTA grid build then runs alone (3x faster than
        # beside the trunk's convolutions), and the trunk / LBS kernels fill in beside K1 as its short-lived blocks
        # retire.  MPSNERF_PREP_ORDER=front_first | together forces one order (A/B).
        # When the trunk is the longer pole -- large views, or this GPU's share of the rays is small (one frame over 8
        # GPUs: trunk 0.6 ms vs K1 + K3 0.5 ms at 3 x 1000 x 1000) -- all branches start together instead: the grid
        # build slows down, but K4 no longer waits for the trunk (8 GPUs: 2.58 -> 2.51 ms; 1 GPU: the other way round).
        # Estimates: trunk 0.2 us per 1000 source pixels, K1 + K3 64 ns per 1000 sample points (measured, B200).
        order = os.environ.get("MPSNERF_PREP_ORDER", "auto")
        if order == "auto":
            V_, H_, W_ = ins[0].shape[0], ins[0].shape[-2], ins[0].shape[-1]
            trunk_long = n_points is not None and 0.2e-6 * V_ * H_ * W_ > 0.5 * 0.064e-6 * n_points
            order = "together" if (trunk and trunk_long) else "front_first"
        front_first = order != "together"
        if not front_first:
            for st in (s_lbs, s_trunk):
                st.wait_stream(main)
        graphs[0].replay()                       # front: on the caller's stream, K1 follows it
        if front_first:
            for st in (s_lbs, s_trunk):
                st.wait_stream(main)
        with torch.cuda.stream(s_lbs):
            graphs[1].replay()
            ctx.ev_lbs = torch.cuda.Event()
            ctx.ev_lbs.record()
        ctx.pending_trunk = None
        if host_img and trunk:
            with torch.cuda.stream(s_trunk):            # H2D on the copy engine, ordered in front of the trunk graph
                static[0].copy_(ins[0], non_blocking=True)
        if trunk and self.trunk_shard is None:
            with torch.cuda.stream(s_trunk):
                graphs[2].replay()
                ctx.ev_trunk = torch.cuda.Event()
                ctx.ev_trunk.record()
        elif trunk:
            def sharded():
                with torch.cuda.stream(s_trunk):
                    self._prep_trunk_sharded(static, ctx)
                    ctx._latent.record_stream(main)
                    ctx._img4.record_stream(main)
                    ctx.ev_trunk = torch.cuda.Event()
                    ctx.ev_trunk.record()
            ctx.ev_trunk = None
            ctx.pending_trunk = sharded          # launched by run() right behind K1 (FrameContext.launch_pending)
        _lib.count_launches(4)
        return ctx

    def _new_ctx(self, ins, smpl):
        """Allocate the outputs of the preparation (on the current stream) for the given (static) inputs."""
        img_all, verts = ins[0], ins[12]
        dev = img_all.device
        ctx = FrameContext()
        ctx.n_views = img_all.shape[0]
        ctx.skin_w = self._smpl_tables(smpl, dev)["weights"]
        ctx.frame_dev = torch.empty(ctypes.sizeof(_lib.Frame), dtype=torch.uint8, device=dev)
        gb = self.lib.mpsnerf_grid_bytes(verts.shape[0])
        ctx.grid_tp = torch.empty(gb, dtype=torch.uint8, device=dev)
        ctx.grid_tv = torch.empty(gb, dtype=torch.uint8, device=dev)
        ctx.keep = (list(ins), smpl)
        ctx._latent = ctx._img4 = ctx.ev_lbs = ctx.ev_trunk = ctx.pending_trunk = None
        return ctx

    @staticmethod
    def _feat_size(H, W):
        return (H // 2 - 1) // 2 + 1, (W // 2 - 1) // 2 + 1     # trunk output size: floor(H/2), then conv1 7x7 / 2 pad 3

    def _prep_front(self, ins, ctx):
        """What K1 needs: the copied fields of the frame (Th / R of the target pose, cameras, sizes) and the grid over
        the target-pose vertices in SMPL space (it reads the raw Th / R, not the frame)."""
        img_all, keep, verts = ins[0], ins[1:12], ins[12]
        lib = self.lib
        H, W = img_all.shape[-2:]
        Hf, Wf = self._feat_size(H, W)
        # keep = [poses, shapes, R, Th](target), [poses, shapes, R, Th](source), R_all, T_all, K_all
        _lib.check(lib.mpsnerf_frame_header(_lib.ptr(keep[2]), _lib.ptr(keep[3]), _lib.ptr(keep[6]), _lib.ptr(keep[7]),
                                            _lib.ptr(keep[8]), _lib.ptr(keep[9]), _lib.ptr(keep[10]), ctx.n_views, W, H, Wf, Hf,
                                            _lib.ptr(ctx.frame_dev), _stream()), "frame_header")
        nv = verts.shape[0]
        _lib.check(lib.mpsnerf_grid_build(_lib.ptr(verts), nv, _lib.ptr(keep[3]), _lib.ptr(keep[2]), GRID_CELL_TARGET,
                                          _lib.ptr(ctx.grid_tp), ctx.grid_tp.numel(), _stream()), "grid_build(target)")

    def _prep_lbs(self, ins, ctx, side):
        """What K3 needs: the four LBS transform sets (K0, float64 -> fp32) and, beside it, the template grid."""
        keep, tverts = ins[1:12], ins[13]
        lib = self.lib
        cur = torch.cuda.current_stream()
        tab = self._smpl_tables(ctx.keep[1], tverts.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            _lib.check(lib.mpsnerf_grid_build(_lib.ptr(tverts), tverts.shape[0], None, None, GRID_CELL_TEMPLATE,
                                              _lib.ptr(ctx.grid_tv), ctx.grid_tv.numel(), _stream()), "grid_build(template)")
        _lib.check(lib.mpsnerf_frame_transforms(_lib.ptr(keep[0]), _lib.ptr(keep[1]), _lib.ptr(keep[4]), _lib.ptr(keep[5]),
                                                _lib.ptr(tab["v_template"]), _lib.ptr(tab["shapedirs"]),
                                                _lib.ptr(tab["J_regressor"]), _lib.ptr(tab["parents"]),
                                                tab["v_template"].shape[0], _lib.ptr(ctx.frame_dev), _stream()),
                   "frame_transforms")
        cur.wait_stream(side)

    def set_trunk_shard(self, rank=None, world=None):
        """Split the encoder trunk of every following frame over `world` ranks (this rank computes 1/world of the
        latent rows, the bands are all-gathered with torch.distributed); None switches back to the replicated trunk."""
        self.trunk_shard = None if (rank is None or not world or world <= 1) else (int(rank), int(world))

    def _prep_trunk_sharded(self, ins, ctx):
        """The trunk branch when one frame is split over several GPUs: this rank's band of latent rows, an
        all-gather of the bands (NVLink; the only collective on the render path, and it runs beside K1 / K3 on the
        trunk stream), NHWC assembly.  Eager (no CUDA graph around the collective)."""
        import torch.distributed as dist
        from .parallel import ray_block
        rank, world = self.trunk_shard
        img_all = ins[0]
        V = img_all.shape[0]
        H, W = img_all.shape[-2:]
        Hf, Wf = self._feat_size(H, W)
        a, b = ray_block(Hf, rank, world)
        mr = (Hf + world - 1) // world
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=(self.precision != "fp32")):
            band = self._encode_rows(img_all, a, b)                                  # (V, 128, b - a, Wf)
        mine = torch.zeros(V, mr, Wf, 128, device=img_all.device)
        mine[:, :b - a] = band.permute(0, 2, 3, 1)
        gathered = torch.empty(world, V, mr, Wf, 128, device=img_all.device)
        dist.all_gather_into_tensor(gathered, mine)
        ctx._latent = assemble_latent_rows(gathered, Hf)
        ctx._img4 = F.pad(img_all.permute(0, 2, 3, 1), (0, 1)).contiguous().float()

    def _prep_trunk(self, ins, ctx):
        """What K4 needs: encoder trunk once per frame (cuDNN; boundary of the hot path), NHWC for the gather
        (cuDNN convolutions default to TF32; the fp32 precision mode keeps true fp32 end to end)."""
        img_all = ins[0]
        H, W = img_all.shape[-2:]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=(self.precision != "fp32")):
            latent = self._encode(img_all)
        assert latent.shape[-2:] == self._feat_size(H, W), (latent.shape, H, W)
        ctx._latent = latent.permute(0, 2, 3, 1).contiguous().float()
        ctx._img4 = F.pad(img_all.permute(0, 2, 3, 1), (0, 1)).contiguous().float()

    # ------------------------------------------------------------------ the hot path
    @torch.no_grad()
    def run(self, ctx, rays8=None, S=1, t_vals=None, u=None, points=None, composite=True, occupancy=False,
            all_active=False):
        """rays8 (N,8) [o,d,near,far] or points (P,3).  Returns a dict of flat tensors."""
        lib, dev = self.lib, ctx.frame_dev.device
        if points is not None:
            P, N = points.shape[0], points.shape[0]
            S = 1
        else:
            N = rays8.shape[0]
            P = N * S
        V = ctx.n_views
        raw = torch.empty(P, 4, device=dev)
        mask = torch.empty(P, device=dev)
        sq = torch.empty(P, 3, device=dev)
        ss = torch.empty(P, 3, device=dev)
        out = {"raw": raw, "pts_mask": mask, "smpl_query_pts": sq, "smpl_src_pts": ss}
        if P == 0:
            if composite and points is None:
                out.update(rgb_map=torch.empty(N, 3, device=dev), disp_map=torch.empty(N, device=dev),
                           acc_map=torch.empty(N, device=dev), depth_map=torch.empty(N, device=dev))
            out["n_active"] = 0
            return out
        counter = self._buf("counter", 256, dev).view(torch.int32)
        device_count = False
        k6_done = False
        # One C call for the whole frame (mpsnerf_render_rays_bf16: K1, K3, K4, T, M, K6 with the active count read on
        # the device) whenever nothing needs the stages one by one (per-stage timers, debug captures, direct point
        # queries); MPSNERF_FUSED_CALL=0 keeps the staged calls.
        fused = (self.precision == "bf16" and self.debug is None and self.timers is None and self._use_device_count
                 and self._use_fused and composite and points is None and not all_active
                 and ctx.pending_trunk is None)
        if fused:
            cap = int(min(self.slab, P))
            wsb = lib.mpsnerf_render_rays_workspace(N, S, V, cap)
            ws_all = self._buf("render_ws", wsb, dev)
            a4 = (4 * P + 255) // 256 * 256              # layout of the workspace head: act_pid | act_idx2 | act_q
            act_pid = ws_all[:4 * P].view(torch.int32)
            act_idx2 = ws_all[a4:a4 + 4 * P].view(torch.int32)
            act_q = ws_all[2 * a4:2 * a4 + 12 * P].view(torch.float32)
            out.update(rgb_map=torch.empty(N, 3, device=dev), disp_map=torch.empty(N, device=dev),
                       acc_map=torch.empty(N, device=dev), depth_map=torch.empty(N, device=dev))
            packed = self._packed_weights(dev)
            ev = lambda e: None if e is None else ctypes.c_void_p(e.cuda_event)
            pinned = self._pinned_count(dev)
            count_event = self._count_event(dev)
            _lib.check(lib.mpsnerf_render_rays_bf16(
                _lib.ptr(rays8), N, S, _lib.ptr(t_vals), _lib.ptr(u), _lib.ptr(ctx.frame_dev), _lib.ptr(ctx.grid_tp),
                _lib.ptr(ctx.grid_tv), _lib.ptr(ctx.skin_w), _lib.ptr(ctx._latent), _lib.ptr(ctx._img4), _lib.ptr(packed),
                packed.numel(), V, 1 if occupancy else 0, _lib.ptr(raw), _lib.ptr(mask), _lib.ptr(sq), _lib.ptr(ss),
                _lib.ptr(out["rgb_map"]), _lib.ptr(out["disp_map"]), _lib.ptr(out["acc_map"]), _lib.ptr(out["depth_map"]),
                _lib.ptr(counter), cap, _lib.ptr(ws_all), wsb, ev(ctx.ev_lbs), ev(ctx.ev_trunk),
                ctypes.c_void_p(pinned.data_ptr()), ev(count_event), _stream()), "render_rays_bf16")
            _lib.count_launches(6)
            device_count, k6_done, done_upto = True, True, cap
        else:
            act_pid = self._buf("act_pid", 4 * P, dev).view(torch.int32)
            act_idx2 = self._buf("act_idx2", 4 * P, dev).view(torch.int32)
            act_q = self._buf("act_q", 12 * P, dev).view(torch.float32)
        if fused:
            pass
        elif all_active:      # extract_mesh: every point is evaluated, canonical = the point itself
            act_pid[:P] = torch.arange(P, device=dev, dtype=torch.int32)
            act_q[:3 * P] = points.reshape(-1)
            mask.fill_(1.0)
            sq.zero_()
            ss.zero_()
            n_act = P
        else:
            counter[:1].zero_()
            with self.span("k1_sample_knn"):
              _lib.check(lib.mpsnerf_sample_knn(
                _lib.ptr(rays8), N, S, _lib.ptr(t_vals), _lib.ptr(u), _lib.ptr(points), _lib.ptr(ctx.frame_dev),
                _lib.ptr(ctx.grid_tp), _lib.ptr(raw), _lib.ptr(mask), _lib.ptr(sq), _lib.ptr(ss), _lib.ptr(act_pid),
                _lib.ptr(act_idx2), _lib.ptr(act_q), _lib.ptr(counter), _stream()), "sample_knn")
            _lib.count_launches(1)
            ctx.launch_pending()
            # Device-side count (tensor-core path): K3, K4, T and M are enqueued right behind K1 for the first
            # `cap` active points and read the real count on the device; the host reads it from pinned memory
            # only after everything is in the queue, so the GPU never waits for the CPU in the middle of a
            # frame.  Whatever exceeds `cap` goes through the host-count path below.
            device_count = (self.precision == "bf16" and self.debug is None and self._use_device_count)
            if device_count:
                cap = int(min(self.slab, P))
                pinned = self._pinned_count(dev)
                pinned.copy_(counter[:1], non_blocking=True)
                count_event = torch.cuda.Event()
                count_event.record()
                self._dense_slab_dc(ctx, V, cap, counter, act_pid, act_idx2, act_q, ss, raw, dev)
                done_upto = cap
            else:
                n_act = int(counter[0].item())          # host-count path: the one host sync of the frame
                done_upto = 0

        def composite_now():
            if "rgb_map" not in out:
                out.update(rgb_map=torch.empty(N, 3, device=dev), disp_map=torch.empty(N, device=dev),
                           acc_map=torch.empty(N, device=dev), depth_map=torch.empty(N, device=dev))
            with self.span("k6_composite"):
                _lib.check(lib.mpsnerf_composite(_lib.ptr(raw), _lib.ptr(rays8), N, S, _lib.ptr(t_vals), _lib.ptr(u), None,
                                                 1 if occupancy else 0, _lib.ptr(out["rgb_map"]), _lib.ptr(out["disp_map"]),
                                                 _lib.ptr(out["acc_map"]), _lib.ptr(out["depth_map"]), None, None, _stream()),
                           "composite")
            _lib.count_launches(1)

        if not all_active and device_count:
            # K6 goes into the queue BEFORE the host looks at the count: the frame is then enqueued without a single
            # host wait, and the (rare) overflow beyond `cap` is handled afterwards -- remainder slabs, K6 once more
            if composite and points is None and not k6_done:
                composite_now()
                k6_done = True
            count_event.synchronize()
            n_act = int(pinned[0])
        elif all_active:
            done_upto = 0
        out["n_active"] = n_act
        self.last_active = n_act
        dbg = self.debug
        if dbg is not None:
            dbg.update(act_pid=act_pid[:n_act].clone(), act_idx2=act_idx2[:n_act].clone(),
                       act_q=act_q[:3 * n_act].reshape(-1, 3).clone(), xc=[], idx3=[], xw=[], uv=[], tokens=[])
        ld = _lib.TOKEN_DIM if self.precision == "fp32" else _lib.TOKEN_LD
        slab = self.slab
        if n_act > done_upto:
            cap = min(slab, n_act - done_upto)
            xc = self._buf("xc", 12 * cap, dev).view(torch.float32)
            uv = self._buf("uv", 8 * V * cap, dev).view(torch.float32)
            if self.precision == "fp32":
                tokens = self._buf("tokens", 4 * ld * V * cap, dev).view(torch.float32)
            else:       # tensor-core path: fp16 tokens (halves the largest HBM stream of the frame)
                tokens = self._buf("tokens", 2 * ld * V * cap, dev).view(torch.float16)
            if self.precision == "fp32":
                wtable, keep = self._weights_fp32()
                ws = self._buf("dense", lib.mpsnerf_dense_fp32_workspace(cap, V), dev)
            else:
                packed = self._packed_weights(dev)
                ws = self._buf("dense", lib.mpsnerf_dense_bf16_workspace(cap, V), dev)
            idx3 = self._buf("idx3", 4 * cap, dev).view(torch.int32) if dbg is not None else None
            xw = self._buf("xw", 12 * cap, dev).view(torch.float32) if dbg is not None else None
        if n_act > done_upto:
            ctx.wait_lbs()
        for first in range(done_upto, n_act, slab):
            cnt = min(slab, n_act - first)
            with self.span("k3_deform"):
              _lib.check(lib.mpsnerf_deform_project(
                _lib.ptr(act_pid), _lib.ptr(act_idx2), _lib.ptr(act_q), first, cnt, _lib.ptr(ctx.skin_w),
                _lib.ptr(ctx.frame_dev), _lib.ptr(ctx.grid_tv), _lib.ptr(xc), _lib.ptr(uv), _lib.ptr(ss),
                _lib.ptr(idx3), _lib.ptr(xw), 1 if all_active else 0, _stream()), "deform_project")
            ctx.wait_trunk()
            with self.span("k4_gather"):
              if self.precision == "fp32":
                _lib.check(lib.mpsnerf_gather_tokens(_lib.ptr(uv), cnt, V, _lib.ptr(ctx.frame_dev), _lib.ptr(ctx.latent),
                                                     _lib.ptr(ctx.img4), _lib.ptr(tokens), ld, _stream()), "gather_tokens")
              else:
                _lib.check(lib.mpsnerf_gather_tokens_f16(_lib.ptr(uv), cnt, V, _lib.ptr(ctx.frame_dev), _lib.ptr(ctx.latent),
                                                         _lib.ptr(ctx.img4), _lib.ptr(tokens), _stream()), "gather_tokens_f16")
            if self.precision == "fp32":
              with self.span("dense"):
                _lib.check(lib.mpsnerf_dense_fp32(_lib.ptr(tokens), ld, _lib.ptr(xc), cnt, V, wtable, _lib.ptr(act_pid),
                                                  first, _lib.ptr(raw), _lib.ptr(ws), _stream()), "dense_fp32")
                _lib.count_launches(2 + 30)
            else:
              dense_args = (_lib.ptr(tokens), ld, _lib.ptr(xc), cnt, V, _lib.ptr(packed), packed.numel(), _lib.ptr(act_pid),
                            first, _lib.ptr(raw), _lib.ptr(ws), _stream())
              with self.span("dense_t"):        # cross-view transformer (tcgen05), tokens -> tok0 / tok1
                _lib.check(lib.mpsnerf_xformer_bf16(*dense_args), "xformer_bf16")
              with self.span("dense_m"):        # NeRF MLP (tcgen05), tok0 / tok1 / x_c -> raw
                _lib.check(lib.mpsnerf_mlp_bf16(*dense_args), "mlp_bf16")
              _lib.count_launches(2 + 2)
            if dbg is not None:
                dbg["xc"].append(xc[:3 * cnt].reshape(-1, 3).clone())
                dbg["idx3"].append(idx3[:cnt].clone())
                dbg["xw"].append(xw[:3 * cnt].reshape(-1, 3).clone())
                dbg["uv"].append(uv[:2 * V * cnt].reshape(-1, V, 2).clone())
                dbg["tokens"].append(tokens[:ld * V * cnt].reshape(-1, V, ld).float())
        if composite and points is None and (n_act > done_upto or not k6_done):
            composite_now()
        return out

    def _count_event(self, dev):
        """A (re-recordable) event per device for "the active count has reached the host"; recorded once here so that
        its CUDA handle exists before the C entry point records it."""
        e = self._count_events.get(dev)
        if e is None:
            e = self._count_events[dev] = torch.cuda.Event()
            e.record(torch.cuda.current_stream(dev))
        return e

    def _pinned_count(self, dev):
        t = self._pinned.get(dev)
        if t is None:
            t = self._pinned[dev] = torch.zeros(1, dtype=torch.int32).pin_memory()
        return t

    def _dense_slab_dc(self, ctx, V, cap, counter, act_pid, act_idx2, act_q, ss, raw, dev):
        """K3, K4, T, M for the active-list positions [0, min(count, cap)), count read on the device."""
        lib = self.lib
        ld = _lib.TOKEN_LD
        xc = self._buf("xc", 12 * cap, dev).view(torch.float32)
        uv = self._buf("uv", 8 * V * cap, dev).view(torch.float32)
        tokens = self._buf("tokens", 2 * ld * V * cap, dev).view(torch.float16)
        packed = self._packed_weights(dev)
        ws = self._buf("dense", lib.mpsnerf_dense_bf16_workspace(cap, V), dev)
        ctx.wait_lbs()
        with self.span("k3_deform"):
            _lib.check(lib.mpsnerf_deform_project_dc(_lib.ptr(act_pid), _lib.ptr(act_idx2), _lib.ptr(act_q), 0, cap, _lib.ptr(counter),
                                                     _lib.ptr(ctx.skin_w), _lib.ptr(ctx.frame_dev), _lib.ptr(ctx.grid_tv),
                                                     _lib.ptr(xc), _lib.ptr(uv), _lib.ptr(ss), _stream()), "deform_project_dc")
        ctx.wait_trunk()
        with self.span("k4_gather"):
            _lib.check(lib.mpsnerf_gather_tokens_f16_dc(_lib.ptr(uv), 0, cap, _lib.ptr(counter), V, _lib.ptr(ctx.frame_dev),
                                                        _lib.ptr(ctx.latent), _lib.ptr(ctx.img4), _lib.ptr(tokens), _stream()),
                       "gather_tokens_f16_dc")
        dense_args = (_lib.ptr(tokens), _lib.ptr(xc), 0, cap, _lib.ptr(counter), V, _lib.ptr(packed), packed.numel(),
                      _lib.ptr(act_pid), _lib.ptr(raw), _lib.ptr(ws), _stream())
        with self.span("dense_t"):
            _lib.check(lib.mpsnerf_xformer_bf16_dc(*dense_args), "xformer_bf16_dc")
        with self.span("dense_m"):
            _lib.check(lib.mpsnerf_mlp_bf16_dc(*dense_args), "mlp_bf16_dc")
        _lib.count_launches(4)

    def _live_dense(self):
        """The 46 live tensors of the transformer + MLP (cached: walking all ~370 parameters of the network to
        collect their versions cost 0.1 ms of host time per frame)."""
        if self._dense_params is None:
            named = dict(self.net.named_parameters())
            self._dense_params = [named[k] for k in DENSE_FP32_ORDER]
        return self._dense_params

    def _trunk_state(self):
        """Parameters and buffers of the live part of the encoder trunk (conv1, bn1, layer1)."""
        if self._trunk_params is None:
            m = self.net.encoder_2d.model
            mods = [m.conv1, m.bn1, m.layer1]
            self._trunk_params = [t for mod in mods for t in list(mod.parameters()) + list(mod.buffers())]
        return self._trunk_params

    def _packed_weights(self, dev):
        from .pack import pack_weights_bf16
        key = tuple(p._version for p in self._live_dense())
        if self._packed is None or key != self._packed_key or self._packed.device != dev:
            self._packed = pack_weights_bf16(self.net, dev)
            self._packed_key = key
        return self._packed
