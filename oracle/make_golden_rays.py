"""Generate tests/golden/rays.npz by executing the reference's own get_rays / get_near_far.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python oracle/make_golden_rays.py
The reference module is imported unmodified; trimesh / imageio / lib.base_utils are stubbed because
they are import-time dependencies only (neither function touches them).
"""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MPSNERF_REF", "/root/reference")


def load_reference():
    for name in ("trimesh", "imageio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    try:
        import cv2  # noqa: F401
    except Exception:
        sys.modules["cv2"] = types.ModuleType("cv2")
    sys.path.insert(0, REF)
    lib = types.ModuleType("lib")
    lib.__path__ = [os.path.join(REF, "lib")]
    sys.modules["lib"] = lib
    sys.modules["lib.base_utils"] = types.ModuleType("lib.base_utils")
    return importlib.import_module("lib.if_nerf_data_utils")


def main():
    from mpsnerf_b200 import synthetic
    U = load_reference()
    out = {}
    for name, kw in (("thuman", dict(kind="thuman", seed=0, H=96, W=96)), ("h36m", dict(kind="h36m", seed=2, H=80, W=120))):
        scene = synthetic.make_scene(**kw)
        K, R, T = (np.asarray(a, dtype=np.float64) for a in scene.cams[scene.target])
        bounds = np.asarray(scene.bounds, dtype=np.float32)
        H, W = scene.H, scene.W
        ray_o, ray_d = U.get_rays(H, W, K, R, T.reshape(3, 1))
        ray_o = ray_o.reshape(-1, 3).astype(np.float32)      # evaluation branch, ref :716-718
        ray_d = ray_d.reshape(-1, 3).astype(np.float32)
        near, far, hit = U.get_near_far(bounds, ray_o, ray_d.copy())
        out.update({f"{name}_K": K, f"{name}_R": R, f"{name}_T": T.ravel(), f"{name}_bounds": bounds,
                    f"{name}_HW": np.array([H, W]), f"{name}_ray_o": ray_o, f"{name}_ray_d": ray_d,
                    f"{name}_near": near.astype(np.float32), f"{name}_far": far.astype(np.float32), f"{name}_hit": hit})
        print(name, H, W, "hits", int(hit.sum()), "of", len(hit))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rays.npz"), **out)


if __name__ == "__main__":
    main()
