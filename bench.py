#!/usr/bin/env python
"""Benchmark of the render hot path:  python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = one full novel-view render (BASELINE.json configs[1]: canonical_transformer network,
3 input views 512x512, 262 144 rays x 64 samples) through the public ``render()`` API.
``value`` times it with all inputs resident in HBM; ``e2e`` times the same call with HOST
(pinned) inputs, H2D copies and the D2H read of the rendered image inside the timed region.
Multi-GPU: one process per GPU (torchrun).  Default for N > 1 = the north-star split (BASELINE configs[2]): the rays
of ONE H36M-shaped 1000x1000 target view dealt out to the N GPUs in interleaved row groups, every rank generating
its own rays on the device (strong scaling, no data-path collective); --mode weak = one target view per GPU.
Time = max over ranks.

``--impl reference`` times the reference algorithm's CPU implementation (the oracle port,
oracle/oracle.py) on a bounded ray sample of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ACTIVE_POINT = 3514880          # algorithmic minimum, transformer + MLP
WORKLOAD = "thuman_512x512_fullframe_V3_S64"
# --workload h36m = BASELINE.json configs[2]: H36M-shaped 1000x1000 full frame, the rays of ONE target view
# split in contiguous blocks over the ranks (strong scaling, no data-path collective); not the bench line
WORKLOADS = {"thuman": ("thuman", "canonical_transformer.txt", WORKLOAD, "512x512"),
             "h36m": ("h36m", "h36m.txt", "h36m_1000x1000_fullframe_V3_S64", "1000x1000")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) == 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": float(np.median([int(r[0]) for r in rows])), "sm_max_mhz": float(rows[0][1]),
                "reasons": reasons, "samples": len(rows)}


def workload_config(workload, world, strong):
    """The ``config`` object of the JSON line: identical in both arms (ours / --impl reference) by construction."""
    w = WORKLOADS[workload]
    return {"workload": w[2], "samples_per_ray": 64, "input_views": 3, "image": w[3],
            "network": "configs/%s (skinning_batch), seeded random weights (synthetic.seeded_state_dict(0, 300))" % w[1],
            "extras": "full reference contract (raw, pts_mask, smpl_query_pts, smpl_src_pts)",
            "includes": "per-frame prep + encoder trunk + K1..K6", "l2": "flushed in front of every timed step (512 MiB fill, outside the step's event pair)",
            "parallelism": (f"ONE target view dealt out to {world} GPUs in interleaved groups of 2 image rows, rays generated "
                            f"on each GPU from the camera; per-frame preparation (encoder trunk, K0, grids) replicated; no "
                            f"data-path collective" if strong
                            else f"one target view per GPU x{world}")}


def resolve_mode(a):
    """N = 1: BASELINE configs[1] (thuman 512x512 frame).  N > 1: BASELINE configs[2] -- ONE H36M-shaped 1000x1000
    frame split over the N GPUs (strong scaling) -- unless --workload / --mode say otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = a.workload or ("thuman" if world == 1 else "h36m")
    strong = (world > 1) if a.mode == "auto" else (a.mode == "strong")
    return world, workload, strong


def build_scene_and_net(precision, target_view, workload="thuman"):
    from mpsnerf_b200 import synthetic
    from mpsnerf_b200.lib import skinnning_batch as SB
    from mpsnerf_b200.model_selection import return_model
    from mpsnerf_b200.parser_config import config_parser
    from mpsnerf_b200 import run_nerf_batch as R
    kind, cfg_file = WORKLOADS[workload][:2]
    scene = synthetic.make_scene(kind, seed=0)
    if target_view is not None and target_view != scene.target:          # other ranks render other target views of the same scene
        K, Rm, T = scene.cams[target_view % 24]
        ro, rd = synthetic.get_rays(scene.H, scene.W, K, Rm, T)
        ro, rd = ro.reshape(-1, 3).copy(), rd.reshape(-1, 3).copy()
        n, f, hit = synthetic.get_near_far(scene.bounds, ro, rd)
        scene.rays_o, scene.rays_d = ro, rd
        scene.near, scene.far = np.zeros(len(ro), np.float32), np.ones(len(ro), np.float32)
        scene.near[hit], scene.far[hit] = n, f
        scene.target = target_view % 24
    SB.set_default_smpl_models(scene.smpl)
    args = config_parser().parse_args(["--config", os.path.join(ROOT, "configs", cfg_file),
                                       "--N_samples", "64", "--precision", precision])
    R.configure(args)
    torch.manual_seed(0)
    net = return_model(args)
    net.load_state_dict(synthetic.seeded_state_dict(0, 300.0), strict=False)
    return scene, net, args


def run_ours(a):
    from mpsnerf_b200 import _lib, synthetic
    from mpsnerf_b200 import run_nerf_batch as R
    from mpsnerf_b200.lib.if_nerf_data_utils import gen_rays8
    from mpsnerf_b200.parallel import interleaved_rows
    import torch.distributed as dist
    world, workload, strong = resolve_mode(a)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL / c10d print their version banner there during start-up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local)
    _lib.check(_lib.load().mpsnerf_check_device(local), "check_device")
    scene, net, args = build_scene_and_net(a.precision, None if (strong or world == 1) else 1 + rank, workload)
    handle = R.NetworkHandle(net).to(dev).eval()
    if strong and world > 1 and os.environ.get("MPSNERF_TRUNK_SHARD", "0") == "1":
        # opt-in: each rank encodes 1/N of the latent rows + one NVLink all-gather -- measured SLOWER than the replicated
        # trunk (8 GPUs, 3 x 1000 x 1000: 2.69 vs 2.59 ms; the 96 MB all-gather costs what the convolutions save)
        net.engine().set_trunk_shard(rank, world)
    H, W = scene.H, scene.W
    n_total = H * W
    Kc, Rc, Tc = scene.cams[scene.target]
    # this rank's share of the target view: all of it, or interleaved groups of two image rows of the ONE frame
    rows = interleaved_rows(H, rank, world, 2).to(dev) if (strong and world > 1) else None
    camera = dict(K=Kc, R=Rc, T=Tc, bounds=scene.bounds, H=H, W=W, rows=rows)
    rays8, box = gen_rays8(H, W, Kc, Rc, Tc, scene.bounds, device=dev, rows=rows)
    n_rays = rays8.shape[0]

    def pin(d):
        return {k: (v.pin_memory() if torch.is_tensor(v) else pin(v) if isinstance(v, dict) else v) for k, v in d.items()}

    def to_dev(d):
        return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else to_dev(v) if isinstance(v, dict) else v)
                for k, v in d.items()}

    sp_h, tp_h = pin(scene.sp_input), pin(scene.tp_input)
    sp_d, tp_d = to_dev(sp_h), to_dev(tp_h)
    # resident inputs in the reference's own layout: rays (1,2,N,3), near / far (1,N,1)
    rays_d = torch.stack([rays8[:, 0:3], rays8[:, 3:6]])[None].contiguous()
    near_d, far_d = rays8[None, :, 6:7].contiguous(), rays8[None, :, 7:8].contiguous()
    rays_h, near_h, far_h = rays_d.cpu().pin_memory(), near_d.cpu().pin_memory(), far_d.cpu().pin_memory()
    kw = dict(network_fn=handle, N_samples=64, perturb=False, use_viewdirs=True, chunk=args.chunk)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    out_h = [torch.empty(1, n_rays, 3).pin_memory(), torch.empty(1, n_rays).pin_memory(), torch.empty(1, n_rays).pin_memory()]
    hot = R.hot_input_bytes(sp_h, tp_h)
    cam_bytes = (9 + 9 + 3 + 6) * 8 + (0 if rows is None else rows.numel() * 4)
    d2h = sum(t.numel() * 4 for t in out_h)

    def step_resident():
        return R.render(rays=rays_d, near=near_d, far=far_d, sp_input=sp_d, tp_input=tp_d, **kw)

    def read_back(rgb, disp, acc):
        out_h[0].copy_(rgb, non_blocking=True)
        out_h[1].copy_(disp, non_blocking=True)
        out_h[2].copy_(acc, non_blocking=True)

    def step_e2e():
        # everything starts on the host: pinned dicts (render() uploads the entries the path reads, R.HOT_KEYS_*) and
        # the target camera; the rays are generated on the device from it (csrc/raygen.cu), the image is read back
        rgb, disp, acc, _ = R.render(camera=camera, sp_input=sp_h, tp_input=tp_h, **kw)
        read_back(rgb, disp, acc)

    def step_e2e_host_rays():
        # round-1 form of the same call: rays / near / far built on the host and uploaded (32 B per ray)
        rgb, disp, acc, _ = R.render(rays=rays_h, near=near_h, far=far_h, sp_input=sp_h, tp_input=tp_h, **kw)
        read_back(rgb, disp, acc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, collective=True):
        """K steps, each bracketed by its own CUDA-event pair on the launch stream, the L2 flushed (512 MiB fill)
        in front of every step OUTSIDE its event pair.  The host does not synchronise between steps -- it enqueues
        ahead, as a renderer streaming frames does -- so a step's time is its device time, not the host's launch
        latency behind an idle GPU; sum over the K steps, max over ranks."""
        for _ in range(warmup):
            fn()
        if collective:
            barrier()
        else:
            torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = [e0.elapsed_time(e1) for e0, e1 in evs]
        if not collective:
            return sum(ms), ms
        barrier()
        t = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ms

    warm = max(a.warmup, 3)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.LAUNCHES
    total_ms, per = timed(step_resident, a.steps, warm)
    launches = (_lib.LAUNCHES - l0) * a.steps // (a.steps + warm)
    clocks = sampler.stop()
    e2e_ms, _ = timed(step_e2e, a.steps, 2)
    e2e_host_ms, _ = timed(step_e2e_host_rays, max(a.steps // 2, 2), 2)

    # per-stage device times of the same step (separate passes, CUDA events on the launch stream)
    eng = net.engine()
    eng.timers = {}
    for _ in range(3):
        flush.fill_(1)
        step_resident()
    torch.cuda.synchronize()
    stage_ms = {k: float(sum(s.elapsed_time(e) for s, e in v)) / 3 for k, v in eng.timers.items()}   # per step
    n_active = eng.last_active
    eng.timers = None

    # N > 1, strong: the same frame on ONE GPU (rank 0, the others wait), in the same run on the same box -- the
    # denominator a reader needs next to the N-GPU value (the driver's own N = 1 line is another workload)
    single = None
    if strong and world > 1:
        sharded = net.engine().trunk_shard
        net.engine().set_trunk_shard(None)                   # the one-GPU reference run cannot wait for the other ranks
        if rank == 0:
            r8, _ = gen_rays8(H, W, Kc, Rc, Tc, scene.bounds, device=dev)
            full = dict(rays=torch.stack([r8[:, 0:3], r8[:, 3:6]])[None].contiguous(), near=r8[None, :, 6:7].contiguous(),
                        far=r8[None, :, 7:8].contiguous())
            ms1, _ = timed(lambda: R.render(sp_input=sp_d, tp_input=tp_d, **full, **kw), max(a.steps // 2, 3), 2, collective=False)
            single = {"ms_per_step": ms1 / max(a.steps // 2, 3), "rays_per_s": n_total * max(a.steps // 2, 3) / (ms1 * 1e-3),
                      "what": "the whole frame on rank 0 alone, same run, resident inputs"}
            del full, r8
        barrier()
        if sharded is not None:
            net.engine().set_trunk_shard(*sharded)
    # load balance of the split: active points per rank
    act = torch.tensor([float(n_active)], device=dev, dtype=torch.float64)
    acts = [act.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(acts, act)
    acts = [int(x.item()) for x in acts]

    pk = peaks()
    FLOP_T, FLOP_M = FLOP_PER_ACTIVE_POINT - 1353728, 1353728       # algorithmic minimum per active point, each kernel
    if "dense" not in stage_ms and "dense_t" in stage_ms:
        stage_ms["dense"] = stage_ms["dense_t"] + stage_ms["dense_m"]
    dense_ms = stage_ms.get("dense", float("nan"))
    # DRAM bytes per launch from the committed `ncu --set full` capture of the same kernels (NOT measured in this run)
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and a.precision == "bf16" and workload == "thuman" and world == 1:
        tj = json.load(open(tpath))
        traffic, tsrc = tj.get("xformer_bf16"), "profiles/ncu_traffic.json: " + str(tj.get("xformer_bf16_source"))
    # HBM-side figures of the other kernels: algorithmic bytes / CUDA-event time
    n_pts, V = n_rays * 64, 3
    Hf, Wf = (H // 2 - 1) // 2 + 1, (W // 2 - 1) // 2 + 1
    alg_bytes = {"k1_sample_knn": n_pts * 44 + n_rays * 32, "k3_deform": n_active * (20 + 48),
                 # K4: compulsory HBM bytes = tokens written + uv read + one pass over the latent / image planes;
                 # its 4 taps x 128 channels per (point, view) are L2 hits (reported as l2_bytes below)
                 "k4_gather": n_active * V * (8 + 160 * (2 if a.precision == "bf16" else 4)) + V * (Hf * Wf * 128 + H * W * 4) * 4,
                 "k6_composite": n_rays * (64 * 16 + 52)}
    kernels = {k: {"bytes": b, "ms": stage_ms.get(k), "achieved_gbs": b / (stage_ms[k] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": b / (stage_ms[k] * 1e-3) / 1e9 / pk["hbm"]} for k, b in alg_bytes.items() if stage_ms.get(k)}
    for k, fl in (("dense_t", FLOP_T), ("dense_m", FLOP_M)):
        if stage_ms.get(k):
            t = n_active * fl / (stage_ms[k] * 1e-3) / 1e12
            kernels[k] = {"flop_per_active_point": fl, "ms": stage_ms[k], "achieved_tflops": t,
                          "frac_of_bf16_sustained_peak": t / pk["bf16_sustained"]}
    if "k4_gather" in kernels:
        kernels["k4_gather"]["l2_bytes"] = n_active * V * (4 * 128 + 16) * 4
        kernels["k4_gather"]["l2_gbs"] = kernels["k4_gather"]["l2_bytes"] / (stage_ms["k4_gather"] * 1e-3) / 1e9
    if a.precision == "bf16" and stage_ms.get("dense_t"):
        # the dominant kernel of the step: the cross-view transformer (xformer_tc_kernel)
        kname, kms, kflop = "xformer_tc_kernel (dense_t)", stage_ms["dense_t"], FLOP_T
    else:
        kname, kms, kflop = "dense_" + a.precision, dense_ms, FLOP_PER_ACTIVE_POINT
    ach = n_active * kflop / (kms * 1e-3) / 1e12
    stage_ach = n_active * FLOP_PER_ACTIVE_POINT / (dense_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_sustained"], "traffic": traffic, "traffic_source": tsrc,
                "peak_source": pk["src"] + " sustained bf16", "kernel_ms": kms, "active_points": n_active,
                "flop_per_active_point": kflop,
                "stage": {"what": "T + M (transformer + MLP), the dense stage", "ms": dense_ms, "achieved": stage_ach,
                          "frac": stage_ach / pk["bf16_sustained"], "flop_per_active_point": FLOP_PER_ACTIVE_POINT},
                "stage_ms": stage_ms, "kernels": kernels, "hbm_peak_gbs": pk["hbm"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline(budget_s=20.0, workload=workload)
    rays_per_step = n_total if strong else n_rays * world
    res = {
        "metric": "rays/sec (render fwd)", "value": rays_per_step * a.steps / (total_ms * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": workload_config(workload, world, strong),
        "e2e": {"value": rays_per_step * a.steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": hot + cam_bytes,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps,
                "what": "render(camera=...) with pinned host dicts: uploads + device ray generation + image read-back",
                "host_rays_variant": {"ms_per_step": e2e_host_ms / max(a.steps // 2, 2),
                                      "h2d_bytes_per_step": hot + sum(t.numel() * 4 for t in (rays_h, near_h, far_h)),
                                      "what": "rays / near / far built on the host and uploaded (round-1 e2e)"}},
        "rays_per_gpu_per_step": n_rays, "active_points_per_gpu": acts,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    if single is not None:
        res["single_gpu_same_workload"] = single
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def run_train(a):
    """--workload train = BASELINE configs[3]: one optimisation step (render forward in training mode, backward,
    gradient all-reduce, Adam) on 1024 rays per GPU x 64 samples x 3 views, data-parallel replicas (weak scaling)."""
    from mpsnerf_b200 import _lib, synthetic
    from mpsnerf_b200 import run_nerf_batch as R
    from mpsnerf_b200.train import TrainStep
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.check(_lib.load().mpsnerf_check_device(local), "check_device")
    scene, net, args = build_scene_and_net("fp32", None, "thuman")
    args.smooth_loss = 1            # shipped configs: every smooth_interval-th (4th) step adds the normal-smoothness terms
    R.configure(args)
    handle = R.NetworkHandle(net).to(dev).train()
    n_rays, S = 1024, 64
    box = np.nonzero(scene.mask_at_box)[0]
    ids = np.sort(np.random.RandomState(100 + rank).choice(box, n_rays, replace=False))      # every rank its own batch
    rays_h, near_h, far_h = (t.pin_memory() for t in synthetic.rays_tensor(scene, ids))
    rng = np.random.RandomState(7 + rank)
    tgt_h = torch.from_numpy(rng.uniform(0, 1, (1, n_rays, 3)).astype(np.float32)).pin_memory()
    msk_h = torch.from_numpy((rng.uniform(0, 1, (1, n_rays, 1)) > 0.5).astype(np.float32)).pin_memory()
    u_h = torch.from_numpy(rng.uniform(0, 1, (1, n_rays, S)).astype(np.float32)).pin_memory()
    pin = lambda d: {k: (v.pin_memory() if torch.is_tensor(v) else pin(v) if isinstance(v, dict) else v) for k, v in d.items()}
    to_dev = lambda d: {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else to_dev(v) if isinstance(v, dict) else v) for k, v in d.items()}
    sp_h, tp_h = pin(scene.sp_input), pin(scene.tp_input)
    sp_d, tp_d = to_dev(sp_h), to_dev(tp_h)

    def at_step(d, g):      # the training loop's counters (ref :541-542), host tensors
        return dict(d, global_step=torch.full((1,), g, dtype=torch.long), smooth_interval=torch.full((1,), 4, dtype=torch.long))

    sp_d, sp_smooth, sp_h = at_step(sp_d, 1), at_step(sp_d, 0), at_step(sp_h, 1)
    res_d = [t.to(dev) for t in (rays_h, near_h, far_h, tgt_h, msk_h, u_h)]
    opt = torch.optim.Adam(list(net.parameters()), lr=args.lrate, betas=(0.9, 0.999))
    ts = TrainStep(handle, opt, acc_loss=bool(args.acc_loss))
    kw = dict(N_samples=S, perturb=1.0, use_viewdirs=True, chunk=args.chunk)
    loss_h = torch.empty(1).pin_memory()

    def step_resident():
        rays, near, far, tgt, msk, u = res_d
        return ts.step(R.render, rays=rays, near=near, far=far, sp_input=sp_d, tp_input=tp_d, target_rgb=tgt, bkgd_msk=msk,
                       perturb_u=u, **kw)

    def step_smooth():      # an interval step: + second K1 / K3 pass on perturbed points, double backward on the active points
        rays, near, far, tgt, msk, u = res_d
        return ts.step(R.render, rays=rays, near=near, far=far, sp_input=sp_smooth, tp_input=tp_d, target_rgb=tgt,
                       bkgd_msk=msk, perturb_u=u, **kw)

    def step_e2e():
        rays, near, far, tgt, msk, u = (t.to(dev, non_blocking=True) for t in (rays_h, near_h, far_h, tgt_h, msk_h, u_h))
        sp, tp = R._upload_hot(sp_h, R.HOT_KEYS_SP + R.STEP_KEYS, dev), R._upload_hot(tp_h, R.HOT_KEYS_TP, dev)
        loss = ts.step(R.render, rays=rays, near=near, far=far, sp_input=sp, tp_input=tp, target_rgb=tgt, bkgd_msk=msk,
                       perturb_u=u, **kw)
        loss_h.copy_(loss.reshape(1), non_blocking=True)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warm = max(a.warmup, 3)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.LAUNCHES
    total_ms = timed(step_resident, a.steps, warm)
    launches = (_lib.LAUNCHES - l0) * a.steps // (a.steps + warm)
    clocks = sampler.stop()
    e2e_ms = timed(step_e2e, a.steps, 2)
    smooth_ms = timed(step_smooth, a.steps, 2)
    h2d = R.hot_input_bytes(sp_h, tp_h) + sum(t.numel() * 4 for t in (rays_h, near_h, far_h, tgt_h, msk_h, u_h))
    n_grad = sum(p.numel() for p in net.parameters() if p.grad is not None)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    print(json.dumps({
        "metric": "rays/sec (training step: fwd + bwd + grad all-reduce + Adam)", "value": n_rays * world * a.steps / (total_ms * 1e-3),
        "unit": "rays/s", "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": total_ms / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "train_thuman_512x512_V3_1024rays_S64", "rays_per_gpu_per_step": n_rays, "samples_per_ray": S,
                   "input_views": 3, "image": "512x512", "loss": "img2mse(rgb) + img2mse(acc), perturb = 1; the timed steps are off the smooth interval "
                                                             "(3 of 4 steps under the shipped configs), the interval step is timed beside them",
                   "optimizer": "Adam", "l2": "flushed in front of every timed step (512 MiB fill, outside the step's event pair)",
                   "parallelism": f"data-parallel replicas x{world}: one NCCL all-reduce of the dense-stage gradient bucket "
                                  f"(launched inside the backward, overlapping the cuDNN trunk backward) + one of the trunk gradients"},
        "e2e": {"value": n_rays * world * a.steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / a.steps},
        "smooth_interval_step": {"ms_per_step": smooth_ms / a.steps, "every": 4,
                                 "amortised_ms_per_step": (3 * total_ms + smooth_ms) / (4 * a.steps),
                                 "what": "step with the normal-smoothness terms (ref run_nerf_batch.py:60-79): second K1 / K3 pass "
                                         "on perturbed points + torch-autograd double backward on the active points (smooth.py)"},
        "gpu_launches": int(launches), "clocks": clocks, "live_gradient_floats": int(n_grad),
        "active_points": int(net.train_engine().eng.last_active)}))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(budget_s=20.0, n_rays=None, steps=1, warmup=0, workload="thuman"):
    """Time the reference algorithm on the host cores on in-box rays of the bench scene (BASELINE config 1 shape):
    the UNMODIFIED reference under import shims when its tree is present ($MPSNERF_REF, /root/reference,
    baseline/_ref; kind "reference"), else the oracle port (kind "port": the GPU box never has the tree)."""
    from mpsnerf_b200 import synthetic
    from oracle import oracle as O
    from oracle import run_reference as RR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scene = synthetic.make_scene(WORKLOADS[workload][0], seed=0)
    sd = synthetic.seeded_state_dict(0, 300.0)
    kind = "port"
    if RR.find_reference() is not None and os.environ.get("MPSNERF_CPU_ARM", "auto") != "port":
        try:
            R, wrapped = RR.load(scene, sd)
            kind = "reference"
        except Exception as e:          # a tree that does not import here: say so and time the port
            sys.stderr.write(f"bench: reference tree found but not runnable ({type(e).__name__}: {e}); timing the port\n")
    smpl = O.smpl_tensors(scene.smpl)

    def run(n):
        ids = synthetic.inbox_ray_subset(scene, n)
        t0 = time.perf_counter()
        if kind == "reference":
            RR.render(R, wrapped, scene, ids, S=64)
        else:
            O.render(smpl, sd, scene.sp_input, scene.tp_input, scene.rays_o[ids], scene.rays_d[ids], scene.near[ids],
                     scene.far[ids], S=64)
        return time.perf_counter() - t0

    if n_rays is None:
        t = run(256)                                    # calibration (also warms the thread pools)
        n_rays = int(np.clip(256 * budget_s / max(t, 1e-3) / max(steps + warmup, 1), 256, 4096))
    for _ in range(warmup):
        run(n_rays)
    ts = [run(n_rays) for _ in range(steps)]
    what = "the unmodified reference (oracle/ref_shims.py)" if kind == "reference" else "oracle port (oracle/oracle.py)"
    return {"value": n_rays * len(ts) / sum(ts), "unit": "rays/s", "cores": cores, "kind": kind,
            "sample": f"{n_rays} in-box rays x 64 samples of the same scene and weights, {len(ts)} pass(es), {what}, "
                      f"torch CPU fp32, {cores} threads, encoder included once per pass", "seconds": sum(ts)}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world, workload, strong = resolve_mode(a)
    cpu = cpu_baseline(budget_s=float(os.environ.get("MPSNERF_CPU_ARM_BUDGET_S", "150")), steps=a.steps, warmup=a.warmup,
                       workload=workload)
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec (render fwd)", "value": cpu["value"], "unit": "rays/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * cpu["seconds"] / a.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, world, strong),
        "note": "each step = a bounded ray sample of the workload on the host cores (cpu_baseline.sample)",
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MPSNERF_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["train"],
                    help="default: thuman (BASELINE configs[1]) on 1 GPU, h36m (configs[2]) on several")
    ap.add_argument("--mode", default="auto", choices=["auto", "strong", "weak"],
                    help="N > 1: strong = ONE frame split over the GPUs (default), weak = one target view per GPU")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "train":
        run_train(a)
    else:
        run_ours(a)
