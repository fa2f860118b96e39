"""Ray generation on the device (mirror of the ray half of ``lib/if_nerf_data_utils.py`` of the reference).

The reference builds the rays of a target view in numpy on the CPU (``get_rays`` :11-25, ``get_near_far``
:55-92) and ships 32 bytes per ray to the GPU.  Here one CUDA kernel (csrc/raygen.cu) writes the (N, 8)
``[o, d, near, far]`` rows that ``render`` / K1 consume straight into device memory.  Same names and
argument meaning as the reference; results are CUDA tensors instead of numpy arrays, and there is no CPU
fallback.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


def _host3(a, shape):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    return np.ascontiguousarray(a, dtype=np.float64).reshape(shape)


def gen_rays8(H, W, K, R, T, bounds, device="cuda", rows=None):
    """-> rays8 (n, 8) float32 [o, d, near, far] and mask_at_box (n,) bool, on ``device``; n = H*W, or
    len(rows)*W when ``rows`` (int32 CUDA tensor / sequence of image-row indices) selects a subset of the view.

    Rays that do not cross the (0.01-widened) box exactly twice keep near = 0, far = 1
    (the full-frame convention of ``sample_ray_THuman``, ref :719-724)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mpsnerf_b200 ray generation runs on CUDA devices only")
    K, R, T, b = _host3(K, (3, 3)), _host3(R, (3, 3)), _host3(T, (3,)), _host3(np.asarray(bounds, dtype=np.float32), (2, 3))
    if rows is not None:
        rows = torch.as_tensor(rows, dtype=torch.int32, device=dev).contiguous()
    n = (int(H) if rows is None else int(rows.numel())) * int(W)
    rays8 = torch.empty(n, 8, device=dev)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    cam = (K.ctypes.data_as(ctypes.c_void_p), R.ctypes.data_as(ctypes.c_void_p), T.ctypes.data_as(ctypes.c_void_p),
           b.ctypes.data_as(ctypes.c_void_p))
    with torch.cuda.device(dev):
        if rows is None:
            _lib.check(_lib.load().mpsnerf_gen_rays(*cam, int(H), int(W), _lib.ptr(rays8), _lib.ptr(mask), st), "gen_rays")
        else:
            _lib.check(_lib.load().mpsnerf_gen_rays_rows(*cam, int(H), int(W), _lib.ptr(rows), int(rows.numel()),
                                                         _lib.ptr(rays8), _lib.ptr(mask), st), "gen_rays_rows")
    _lib.count_launches(1)
    return rays8, mask.bool()


def get_rays(H, W, K, R, T, device="cuda"):
    """ref :11-25 -> (rays_o, rays_d), each (H, W, 3) float32 CUDA tensors."""
    inf = np.array([[-1e30] * 3, [1e30] * 3])
    rays8, _ = gen_rays8(H, W, K, R, T, inf, device)
    return rays8[:, 0:3].reshape(H, W, 3), rays8[:, 3:6].reshape(H, W, 3)


def get_near_far_full(H, W, K, R, T, bounds, device="cuda"):
    """Rays and box hits of a whole view: (rays_o, rays_d, near, far, mask_at_box), near/far compacted to the hits
    like ``get_near_far`` (ref :55-92)."""
    rays8, hit = gen_rays8(H, W, K, R, T, bounds, device)
    return rays8[:, 0:3], rays8[:, 3:6], rays8[hit, 6], rays8[hit, 7], hit
