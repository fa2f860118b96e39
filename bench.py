#!/usr/bin/env python
"""Benchmark of the render hot path:  python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = one full novel-view render (BASELINE.json configs[1]: canonical_transformer network,
3 input views 512x512, 262 144 rays x 64 samples) through the public ``render()`` API.
``value`` times it with all inputs resident in HBM; ``e2e`` times the same call with HOST
(pinned) inputs, H2D copies and the D2H read of the rendered image inside the timed region.
Multi-GPU: one process per GPU (torchrun), every rank renders its own target view of the same
scene (weak scaling, no data-path collective); time = max over ranks.

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
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
    strong = a.workload != "thuman"
    scene, net, args = build_scene_and_net(a.precision, None if strong else 1 + rank, a.workload)
    handle = R.NetworkHandle(net).to(dev).eval()
    rays_h, near_h, far_h = synthetic.rays_tensor(scene, None)
    n_total = rays_h.shape[2]
    if strong:                               # this rank's contiguous block of the one frame
        # Balance by measured work, not by ray count: one untimed profiling render of the whole frame
        # gives the active samples per ray; a ray costs ~0.6 active-point equivalents on its own (K1 + K6).
        from mpsnerf_b200.parallel import balanced_ray_block
        if world > 1:
            dd = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else dd(v) if isinstance(v, dict) else v) for k, v in d.items()}
            ex = R.render(rays=rays_h.to(dev), near=near_h.to(dev), far=far_h.to(dev), sp_input=dd(scene.sp_input),
                          tp_input=dd(scene.tp_input), network_fn=handle, N_samples=64, perturb=False, use_viewdirs=True)[3]
            weights = ex["pts_mask"][0, ..., 0].sum(-1).double().cpu() + 0.6
            del ex
            torch.cuda.empty_cache()
        else:
            weights = scene.mask_at_box
        s0, s1 = balanced_ray_block(weights, rank, world)
        rays_h, near_h, far_h = rays_h[:, :, s0:s1].contiguous(), near_h[:, s0:s1].contiguous(), far_h[:, s0:s1].contiguous()
    n_rays = rays_h.shape[2]

    def pin(d):
        return {k: (v.pin_memory() if torch.is_tensor(v) else pin(v) if isinstance(v, dict) else v) for k, v in d.items()}

    def to_dev(d):
        return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else to_dev(v) if isinstance(v, dict) else v)
                for k, v in d.items()}

    sp_h, tp_h = pin(scene.sp_input), pin(scene.tp_input)
    rays_h, near_h, far_h = rays_h.pin_memory(), near_h.pin_memory(), far_h.pin_memory()
    sp_d, tp_d = to_dev(sp_h), to_dev(tp_h)
    rays_d, near_d, far_d = rays_h.to(dev), near_h.to(dev), far_h.to(dev)
    kw = dict(network_fn=handle, N_samples=64, perturb=False, use_viewdirs=True, chunk=args.chunk)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    out_h = [torch.empty(1, n_rays, 3).pin_memory(), torch.empty(1, n_rays).pin_memory(), torch.empty(1, n_rays).pin_memory()]

    def tensor_bytes(d):
        return sum(v.numel() * v.element_size() if torch.is_tensor(v) else tensor_bytes(v) if isinstance(v, dict) else 0
                   for v in d.values())

    h2d = R.hot_input_bytes(sp_h, tp_h) + sum(t.numel() * 4 for t in (rays_h, near_h, far_h))
    d2h = sum(t.numel() * 4 for t in out_h)

    def step_resident():
        return R.render(rays=rays_d, near=near_d, far=far_d, sp_input=sp_d, tp_input=tp_d, **kw)

    def step_e2e():
        # everything starts in pinned host memory: render() uploads the entries of sp / tp that the path reads
        # (R.HOT_KEYS_*; not, e.g., the target view's own images) and, on a copy stream under the frame
        # preparation, rays / near / far
        rgb, disp, acc, _ = R.render(rays=rays_h, near=near_h, far=far_h, sp_input=sp_h, tp_input=tp_h, **kw)
        out_h[0].copy_(rgb, non_blocking=True)
        out_h[1].copy_(disp, non_blocking=True)
        out_h[2].copy_(acc, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ms = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
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
    e2e_ms, _ = timed(step_e2e, a.steps, 1)

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
    pk = peaks()
    if "dense" not in stage_ms and "dense_t" in stage_ms:
        stage_ms["dense"] = stage_ms["dense_t"] + stage_ms["dense_m"]
    dense_ms = stage_ms.get("dense", float("nan"))
    ach = n_active * FLOP_PER_ACTIVE_POINT / (dense_ms * 1e-3) / 1e12
    # DRAM bytes of the dense stage (T + M kernels) per step from the committed `ncu --set full` capture
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dense_" + a.precision)
    # HBM-side figures of the other kernels: algorithmic bytes / CUDA-event time
    n_pts, V = n_rays * 64, 3
    alg_bytes = {"k1_sample_knn": n_pts * 44 + n_rays * 32, "k3_deform": n_active * (20 + 48),
                 # K4: compulsory HBM bytes = tokens written + uv read + one pass over the latent / image planes;
                 # its 4 taps x 128 channels per (point, view) are L2 hits (reported as l2_bytes below)
                 "k4_gather": n_active * V * (8 + 160 * (2 if a.precision == "bf16" else 4)) + V * (128 * 128 * 128 + 512 * 512 * 4) * 4,
                 "k6_composite": n_rays * (64 * 16 + 52)}
    kernels = {k: {"bytes": b, "ms": stage_ms.get(k), "achieved_gbs": b / (stage_ms[k] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": b / (stage_ms[k] * 1e-3) / 1e9 / pk["hbm"]} for k, b in alg_bytes.items() if stage_ms.get(k)}
    # the two fused tensor-core kernels on their own (algorithmic minimum FLOP per active point each)
    for k, fl in (("dense_t", FLOP_PER_ACTIVE_POINT - 1353728), ("dense_m", 1353728)):
        if stage_ms.get(k):
            t = n_active * fl / (stage_ms[k] * 1e-3) / 1e12
            kernels[k] = {"flop_per_active_point": fl, "ms": stage_ms[k], "achieved_tflops": t,
                          "frac_of_bf16_sustained_peak": t / pk["bf16_sustained"]}
    if "k4_gather" in kernels:
        kernels["k4_gather"]["l2_bytes"] = n_active * V * (4 * 128 + 16) * 4
        kernels["k4_gather"]["l2_gbs"] = kernels["k4_gather"]["l2_bytes"] / (stage_ms["k4_gather"] * 1e-3) / 1e9
    roofline = {"bound": "tensor", "kernel": "dense_" + a.precision, "achieved": ach, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"], "traffic": traffic, "peak_source": pk["src"] + " sustained bf16",
                "kernel_ms": dense_ms, "active_points": n_active, "flop_per_active_point": FLOP_PER_ACTIVE_POINT,
                "stage_ms": stage_ms, "kernels": kernels, "hbm_peak_gbs": pk["hbm"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline(budget_s=20.0)
    rays_per_step = n_total if strong else n_rays * world
    res = {
        "metric": "rays/sec (render fwd)", "value": rays_per_step * a.steps / (total_ms * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
        "config": {"workload": WORKLOADS[a.workload][2], "rays_per_gpu_per_step": n_rays, "samples_per_ray": 64, "input_views": 3,
                   "image": WORKLOADS[a.workload][3], "network": "configs/%s (skinning_batch), seeded random weights" % WORKLOADS[a.workload][1],
                   "extras": "full reference contract (raw, pts_mask, smpl_query_pts, smpl_src_pts)",
                   "includes": "per-frame prep + encoder trunk + K1..K6", "l2": "flushed between timed steps (512 MiB fill)",
                   "parallelism": (f"rays: {world} contiguous blocks of one target view, balanced by active samples per ray "
                                   f"(one untimed profiling render)" if strong
                                   else f"rays: one target view per GPU x{world}")},
        "e2e": {"value": rays_per_step * a.steps / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(budget_s=20.0, n_rays=None, steps=1, warmup=0):
    """Time the oracle port (reference algorithm on CPU) on in-box rays of the same scene."""
    from mpsnerf_b200 import synthetic
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scene = synthetic.make_scene("thuman", seed=0)
    sd = synthetic.seeded_state_dict(0, 300.0)
    smpl = O.smpl_tensors(scene.smpl)

    def run(n):
        ids = synthetic.inbox_ray_subset(scene, n)
        t0 = time.perf_counter()
        O.render(smpl, sd, scene.sp_input, scene.tp_input, scene.rays_o[ids], scene.rays_d[ids], scene.near[ids],
                 scene.far[ids], S=64)
        return time.perf_counter() - t0

    if n_rays is None:
        t = run(256)                                    # calibration (also warms the thread pools)
        n_rays = int(np.clip(256 * budget_s / max(t, 1e-3) / max(steps + warmup, 1), 256, 4096))
    for _ in range(warmup):
        run(n_rays)
    ts = [run(n_rays) for _ in range(steps)]
    return {"value": n_rays * len(ts) / sum(ts), "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{n_rays} in-box rays x 64 samples of the same scene, {len(ts)} pass(es), torch CPU fp32, "
                      f"{cores} threads, encoder included once per pass", "seconds": sum(ts)}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cpu = cpu_baseline(budget_s=150.0, steps=a.steps, warmup=a.warmup)
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec (render fwd)", "value": cpu["value"], "unit": "rays/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * cpu["seconds"] / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "each step = a bounded ray sample of the workload on host cores"},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MPSNERF_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--workload", default="thuman", choices=sorted(WORKLOADS))
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
