"""World-size-2 gloo test of the ray-sharding host logic (no GPU needed: the render function is
a stand-in, the product render has no CPU path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(rays, near, far, **kw):
    o, d = rays[:, 0], rays[:, 1]
    rgb = o + d * near
    acc = (d * far).sum(-1)
    return [rgb, acc * 2, acc, {}]


def _worker(rank, world, port, n_rays, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpsnerf_b200 import parallel
    g = torch.Generator().manual_seed(0)
    rays = torch.randn(1, 2, n_rays, 3, generator=g)
    near, far = torch.rand(1, n_rays, 1, generator=g), torch.rand(1, n_rays, 1, generator=g) + 1
    (rgb, disp, acc, _), (s, e) = parallel.render_sharded(_fake_render, rays, near, far)
    assert rgb.shape[1] == e - s
    full_rgb = parallel.gather_frame(rgb, n_rays)
    full_acc = parallel.gather_frame(acc, n_rays)
    ref = _fake_render(rays, near, far)
    assert torch.equal(full_rgb, ref[0]) and torch.equal(full_acc, ref[2])
    assert parallel.max_over_ranks(rank + 1.5, "cpu") == world + 0.5
    # one frame dealt out by interleaved row groups: every rank fills its own rows, the gathered frame is whole
    H, W = 37, 5
    frame = torch.arange(H * W * 3, dtype=torch.float32).reshape(H * W, 3)
    rows = parallel.interleaved_rows(H, rank, world, group=2)
    mine = frame.reshape(H, W, 3)[rows.long()].reshape(-1, 3)
    assert torch.equal(parallel.gather_rows_frame(mine, H, W, group=2), frame)
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_ray_blocks_partition():
    from mpsnerf_b200.parallel import ray_block
    for n in (0, 1, 7, 262144, 1000001):
        for w in (1, 2, 3, 8):
            blocks = [ray_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_rows_partition_and_grid_slabs():
    from mpsnerf_b200.parallel import grid_slab, interleaved_rows
    for H in (1, 2, 7, 512, 1000):
        for w in (1, 2, 4, 8):
            for group in (1, 2, 4):
                parts = [interleaved_rows(H, r, w, group) for r in range(w)]
                allr = torch.cat(parts)
                assert allr.dtype == torch.int32 and sorted(allr.tolist()) == list(range(H))
                if H >= 64 * w:            # even shares to within one row group
                    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= group
    for n in (1, 5, 256):
        for w in (1, 2, 8):
            slabs = [grid_slab(n, r, w) for r in range(w)]
            assert slabs[0][0] == 0 and slabs[-1][1] == n and all(slabs[i][1] == slabs[i + 1][0] for i in range(w - 1))


def test_balanced_blocks_partition():
    from mpsnerf_b200.parallel import balanced_ray_block
    g = torch.Generator().manual_seed(1)
    for n in (1, 9, 1000):
        for w in (1, 2, 4, 8):
            for weights in (torch.zeros(n), torch.ones(n), (torch.rand(n, generator=g) < 0.3).float(),
                            torch.cat([torch.zeros(n // 2), torch.ones(n - n // 2)])):
                blocks = [balanced_ray_block(weights, r, w) for r in range(w)]
                assert blocks[0][0] == 0 and blocks[-1][1] == n
                assert all(blocks[i][1] == blocks[i + 1][0] and blocks[i][0] <= blocks[i][1] for i in range(w - 1))
                if weights.sum() >= 8 * w:           # every block within one ray's weight of the fair share
                    share = float(weights.sum()) / w
                    assert all(abs(float(weights[s:e].sum()) - share) <= 1.0 + 1e-9 for s, e in blocks)


def test_sharded_render_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpsnerf_b200.engine import DENSE_FP32_ORDER
    from mpsnerf_b200.train import DenseBucket, TrainStep

    class Tiny(torch.nn.Module):          # parameters with the names / shapes the bucket walks, nothing else
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(0)
            self._names = {}
            for i, k in enumerate(DENSE_FP32_ORDER):
                p = torch.nn.Parameter(torch.randn(3 + i % 5, 2 + i % 3, generator=g))
                self.register_parameter("p%d" % i, p)
                self._names[k] = p
            self.encoder_2d = torch.nn.Linear(4, 3)

        def named_parameters(self, *a, **k):
            return list(self._names.items()) + [("encoder_2d." + n, p) for n, p in self.encoder_2d.named_parameters()]

    net = Tiny()
    b = DenseBucket(net)
    assert len(b.params) == 46 and b.flat.numel() % 64 == 0
    b.begin_step(world)
    b.pending = 2                          # two render nodes in this step
    for v in b.views:                      # what the backward kernels do: accumulate into the views
        v += float(rank + 1)
    b.node_done()
    assert b.work is None                  # the all-reduce waits for the last node
    for v in b.views:
        v += 10.0 * (rank + 1)
    b.node_done()
    assert b.work is not None
    b.finish()
    want = sum(11.0 * (r + 1) for r in range(world)) / world
    for p in b.params:
        assert p.grad is not None and torch.allclose(p.grad, torch.full_like(p, want))
    # smooth step: autograd adds its share to p.grad, the all-reduce is held back until both shares are merged
    for p in b.params:
        p.grad = None
    b.begin_step(world)
    b.deferred = True
    b.pending = 1
    for v in b.views:
        v += float(rank + 1)               # the kernels' share
    b.node_done()
    assert b.work is None                  # held back
    for p in b.params:
        p.grad = torch.full_like(p, 100.0 * (rank + 1))      # the torch-autograd share (smooth.py)
    b.absorb_autograd()
    assert b.work is not None and all(p.grad is None for p in b.params)
    b.finish()
    want = sum(101.0 * (r + 1) for r in range(world)) / world
    for p in b.params:
        assert torch.allclose(p.grad, torch.full_like(p, want))
    # trunk gradients: one flat all-reduce, averaged
    ts = TrainStep(net, optimizer=None)
    for p in net.encoder_2d.parameters():
        p.grad = torch.full_like(p, float(rank))
    ts.allreduce_trunk(world)
    for p in net.encoder_2d.parameters():
        assert torch.allclose(p.grad, torch.full_like(p, sum(range(world)) / world))
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_gradient_bucket_allreduce_two_ranks_gloo():
    """DenseBucket / TrainStep.allreduce_trunk (the data-parallel plumbing of the training step) on two gloo ranks."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]
