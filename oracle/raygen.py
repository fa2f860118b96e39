"""CPU restatement of the reference's ray generation (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.py).

Follows ``lib/if_nerf_data_utils.py`` of the reference:
  * ``get_rays``      :11-25  pinhole rays, unnormalised directions, camera centre -R^T T;
  * ``get_near_far``  :55-92  six plane intersections with the box widened by 0.01, a ray counts when
                              exactly two of them lie on the box (slack 1e-6), near/far = |p - o| / |d|;
  * the evaluation branch of ``sample_ray_THuman`` :719-724: float32 rays, misses keep near = 0, far = 1.
Pinned by tests/golden/rays.npz, which oracle/make_golden_rays.py produced by executing the reference's own
functions (tests/test_oracle_vs_golden.py::test_raygen_oracle_matches_reference).
"""
import numpy as np


def get_rays(H, W, K, R, T):
    K, R, T = (np.asarray(a, dtype=np.float64) for a in (K, R, T))
    o = -np.dot(R.T, T.reshape(3, 1)).ravel()
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    xy1 = np.stack([i, j, np.ones_like(i)], axis=2)
    cam = np.dot(xy1, np.linalg.inv(K).T)
    world = np.dot(cam - T.ravel(), R)
    d = world - o[None, None]
    return np.broadcast_to(o, d.shape), d


def get_near_far(bounds, ray_o, ray_d):
    b = np.asarray(bounds) + np.array([-0.01, 0.01])[:, None]
    d = ray_d.copy()
    d[d == 0.0] = 1e-8
    t6 = ((b[None] - ray_o[:, None]) / d[:, None]).reshape(-1, 6)
    p = t6[..., None] * d[:, None] + ray_o[:, None]
    eps = 1e-6
    inside = np.all((p >= b.ravel()[:3] - eps) & (p <= b.ravel()[3:] + eps), axis=-1)
    hit = inside.sum(-1) == 2
    dist = np.linalg.norm(p - ray_o[:, None], axis=2) / np.linalg.norm(d, axis=1)[:, None]
    near = np.where(inside, dist, np.inf).min(1)[hit]
    far = np.where(inside, dist, -np.inf).max(1)[hit]
    return near, far, hit


def rays8(H, W, K, R, T, bounds):
    """(H*W, 8) float32 [o, d, near, far] + mask_at_box, the evaluation-branch convention."""
    o, d = get_rays(H, W, K, R, T)
    o = o.reshape(-1, 3).astype(np.float32)
    d = d.reshape(-1, 3).astype(np.float32)
    near, far, hit = get_near_far(np.asarray(bounds, dtype=np.float32), o, d)
    n_all, f_all = np.zeros(len(o), np.float32), np.ones(len(o), np.float32)
    n_all[hit], f_all[hit] = near.astype(np.float32), far.astype(np.float32)
    return np.concatenate([o, d, n_all[:, None], f_all[:, None]], 1), hit
