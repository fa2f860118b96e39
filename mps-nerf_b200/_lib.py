"""ctypes binding of libmpsnerf_b200.so (the C ABI declared in include/mpsnerf.h).

There is no CPU fallback: if the library is missing or was built from other sources than the
ones in the tree (source hash, ``build.py``) it is rebuilt with nvcc; if that fails, or a call
returns an error code, a RuntimeError is raised.
"""
import ctypes
import os

from . import build as _build

c_void_p, c_int, c_int32, c_int64, c_size_t, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int32,
                                                       ctypes.c_int64, ctypes.c_size_t, ctypes.c_float)

MAX_VIEWS = 8
MAX_VIEWS_TC = 4        # tensor-core (bf16) path: a point's V tokens must fit one 128-row tile several times over
NUM_JOINTS = 24
TOKEN_DIM = 155
TOKEN_LD = 160


class Frame(ctypes.Structure):
    """mpsnerf_frame (include/mpsnerf.h)."""
    _fields_ = [
        ("Th_tp", c_float * 3), ("R_tp", c_float * 9), ("Rinv_sp", c_float * 9), ("Th_sp", c_float * 3),
        ("A_tp", c_float * 288), ("A_big_tp", c_float * 288), ("A_big_sp", c_float * 288), ("A_sp", c_float * 288),
        ("cam_R", c_float * (MAX_VIEWS * 9)), ("cam_T", c_float * (MAX_VIEWS * 3)), ("cam_K", c_float * (MAX_VIEWS * 9)),
        ("n_views", c_int32), ("img_w", c_int32), ("img_h", c_int32), ("feat_w", c_int32), ("feat_h", c_int32),
        ("reserved", c_int32 * 3),
    ]


# name -> (restype, argtypes); must list every symbol of include/mpsnerf.h
SIGNATURES = {
    "mpsnerf_last_error": (ctypes.c_char_p, []),
    "mpsnerf_abi_version": (c_int, []),
    "mpsnerf_check_device": (c_int, [c_int]),
    "mpsnerf_frame_prepare": (c_int, [c_void_p] * 11 + [c_int] * 5 + [c_void_p] * 4 + [c_int, c_void_p, c_void_p]),
    "mpsnerf_frame_header": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p, c_void_p]),
    "mpsnerf_frame_transforms": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_void_p]),
    "mpsnerf_grid_bytes": (c_size_t, [c_int]),
    "mpsnerf_grid_build": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_size_t, c_void_p]),
    "mpsnerf_knn1": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_occupancy_fix": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_sample_knn": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "mpsnerf_deform_project": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mpsnerf_gather_tokens": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                      c_void_p]),
    "mpsnerf_deform_project_dc": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_gather_tokens_f16_dc": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p]),
    "mpsnerf_xformer_bf16_dc": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "mpsnerf_mlp_bf16_dc": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "mpsnerf_gen_rays": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_gen_rays_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p,
                                      c_void_p, c_void_p]),
    "mpsnerf_gather_tokens_f16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_dense_fp32_workspace": (c_size_t, [c_int64, c_int]),
    "mpsnerf_dense_fp32": (c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_void_p]),
    "mpsnerf_dense_train_workspace": (c_size_t, [c_int64, c_int]),
    "mpsnerf_dense_train_fwd": (c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_dense_train_bwd": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_composite_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "mpsnerf_gather_tokens_bwd": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "mpsnerf_rows4_gather": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "mpsnerf_rows4_scatter": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "mpsnerf_allreduce_mean": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "mpsnerf_dense_bf16_workspace": (c_size_t, [c_int64, c_int]),
    "mpsnerf_dense_bf16": (c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p,
                                   c_int64, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_xformer_bf16": (c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p,
                                   c_int64, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_mlp_bf16": (c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p,
                                   c_int64, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_render_rays_workspace": (c_size_t, [c_int64, c_int32, c_int, c_int64]),
    "mpsnerf_render_rays_bf16": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int] + [c_void_p] * 9 +
                                 [c_int64, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_render_rays_active_list": (c_int, [c_void_p, c_int64, c_int32, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_composite": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpsnerf_selftest_umma": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mpsnerf_selftest_umma_ts": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mpsnerf_debug_read_prof": (c_int, [c_void_p]),
    "mpsnerf_debug_read_trace": (c_int, [c_void_p]),
}

_lib = None
LAUNCHES = 0          # kernels launched through the ABI (bench.py reports it as gpu_launches)


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) and type every exported symbol."""
    global _lib
    if _lib is None:
        # always go through build(): a cheap source-hash check that rebuilds a stale or missing library (under a
        # file lock, atomically replaced) -- a leftover .so from an older checkout is never loaded silently
        _build.build()
        lib = ctypes.CDLL(_build.LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.mpsnerf_abi_version() != 1:
            raise RuntimeError("libmpsnerf_b200.so: ABI version mismatch")
        _lib = lib
    return _lib


def check(code, what):
    if code != 0:
        msg = load().mpsnerf_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({code}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def count_launches(n):
    global LAUNCHES
    LAUNCHES += n
