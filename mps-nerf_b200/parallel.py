"""Ray sharding across the GPUs of one node.

Rays (and density-grid points) are independent given replicated read-only state, so the render
path needs no exchange step: every rank renders a contiguous block of the ray set.  The only
collectives are optional: assembling the frame on every rank (all_gather of 20 B/ray) and the
max-over-ranks timing used by bench.py.  Works with any torch.distributed backend (NCCL on the
B200 box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def ray_block(n_rays, rank, world):
    """[start, stop) of rank's contiguous block; blocks differ by at most one ray."""
    base, rem = divmod(n_rays, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def balanced_ray_block(weights, rank, world):
    """[start, stop) of rank's contiguous block such that the blocks carry (nearly) equal total ``weights``.

    ``weights`` (n_rays,) is a per-ray work estimate known before rendering -- e.g. the box mask of the ray
    generator (rays that miss the body's box cost almost nothing) -- SURVEY 8e: "balance by in-box / active
    counts, not raw rays".  Falls back to equal counts when all weights are zero."""
    w = torch.as_tensor(weights, dtype=torch.float64).reshape(-1).cpu()
    n = w.numel()
    total = float(w.sum())
    if total <= 0.0 or world == 1:
        return ray_block(n, rank, world)
    c = torch.cumsum(w, 0)
    cuts = [0] + [int(torch.searchsorted(c, torch.tensor(total * r / world, dtype=torch.float64)).item()) for r in range(1, world)] + [n]
    for i in range(1, len(cuts)):            # monotone, in range
        cuts[i] = min(max(cuts[i], cuts[i - 1]), n)
    return cuts[rank], cuts[rank + 1]


def interleaved_rows(H, rank, world, group=2):
    """Image rows of ``rank`` when one H-row target view is dealt out to ``world`` GPUs in groups of ``group`` rows:
    row r belongs to rank (r // group) % world.  Every rank then samples the whole image evenly, so the active
    points -- which cluster on the body's silhouette -- balance by themselves (no profiling pass, no work
    estimate), and each rank generates its own rays on the device from the camera (render(camera=dict(...,
    rows=...))).  Returns an int32 tensor, ascending."""
    r = torch.arange(int(H), dtype=torch.int32)
    return r[((r // int(group)) % int(world)) == int(rank)].contiguous()


def gather_rows_frame(block, H, W, group=2):
    """all_gather per-ray outputs ``block (n_local_rows * W, ...)`` of an ``interleaved_rows`` split into the full
    ``(H * W, ...)`` frame, rows back in image order (20 B/ray for rgb + disp + acc; optional -- the render path
    itself needs no exchange)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return block
    world = dist.get_world_size()
    rows = [interleaved_rows(H, r, world, group) for r in range(world)]
    width = max(len(r) for r in rows) * W
    pad = torch.zeros(width, *block.shape[1:], dtype=block.dtype, device=block.device)
    pad[:block.shape[0]] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty(H * W, *block.shape[1:], dtype=block.dtype, device=block.device)
    out = out.reshape(H, W, *block.shape[1:])
    for p, r in zip(parts, rows):
        out[r.long().to(block.device)] = p[:len(r) * W].reshape(len(r), W, *block.shape[1:])
    return out.reshape(H * W, *block.shape[1:])


def grid_slab(n, rank, world):
    """[z0, z1) of rank's slab when an n^3 density grid (extract_thuman_mesh.py:107-125) is split along its first
    axis; slabs differ by at most one plane.  Points and the post-step are per point: no exchange."""
    return ray_block(int(n), rank, world)


def render_sharded(render_fn, rays, near, far, **kw):
    """Render this rank's block of ``rays (B,2,N,3)`` with ``render_fn`` (= run_nerf_batch.render).

    Returns ([rgb, disp, acc, extras] of the block, (start, stop)).
    """
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    s, e = ray_block(rays.shape[2], rank, world)
    return render_fn(rays=rays[:, :, s:e], near=near[:, s:e], far=far[:, s:e], **kw), (s, e)


def gather_frame(block, n_rays):
    """all_gather per-ray outputs ``block (B, n_local, ...)`` into the full ``(B, n_rays, ...)`` frame."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return block
    world = dist.get_world_size()
    sizes = [ray_block(n_rays, r, world) for r in range(world)]
    width = max(e - s for s, e in sizes)
    pad = torch.zeros(block.shape[0], width, *block.shape[2:], dtype=block.dtype, device=block.device)
    pad[:, :block.shape[1]] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:, :e - s] for p, (s, e) in zip(parts, sizes)], dim=1)


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
