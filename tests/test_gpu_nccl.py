"""mpsnerf_allreduce_mean (include/mpsnerf.h): the C-ABI gradient all-reduce of data-parallel training, on a raw
ncclComm_t created the way a C / C++ host would (ncclGetUniqueId on rank 0, ncclCommInitRank on every rank) -- two
processes, one per GPU.  Needs two GPUs: skipped on a one-GPU box (`gpurun --gpus 2` runs it)."""
import ctypes
import multiprocessing as mp

import pytest
import torch

pytestmark = pytest.mark.gpu


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]


def _worker(rank, world, ids, out):
    try:
        torch.cuda.set_device(rank)
        from mpsnerf_b200 import _lib
        lib = _lib.load()
        nccl = ctypes.CDLL("libnccl.so.2")            # the copy torch already loaded (matched by soname)
        uid = _UniqueId()
        if rank == 0:
            assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
            for _ in range(world - 1):
                ids.put(ctypes.string_at(ctypes.byref(uid), 128))
        else:
            ctypes.memmove(ctypes.byref(uid), ids.get(timeout=120), 128)
        comm = ctypes.c_void_p()
        nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        assert nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
        stream = torch.cuda.Stream()
        for n in (1, 1000, 1_080_064):                # 1 080 064 floats = the dense-stage gradient bucket
            g = torch.Generator().manual_seed(n)
            per_rank = [torch.randn(n, generator=g) for _ in range(world)]
            buf = per_rank[rank].cuda()
            stream.wait_stream(torch.cuda.current_stream())
            _lib.check(lib.mpsnerf_allreduce_mean(comm, _lib.ptr(buf), n, ctypes.c_void_p(stream.cuda_stream)), "allreduce_mean")
            stream.synchronize()
            want = torch.stack(per_rank).double().mean(0).float()
            assert torch.allclose(buf.cpu(), want, atol=1e-6), (n, float((buf.cpu() - want).abs().max()))
        assert lib.mpsnerf_allreduce_mean(None, _lib.ptr(buf), 4, None) == -1          # MPSNERF_EINVAL: no communicator
        assert lib.mpsnerf_allreduce_mean(comm, None, 0, None) == 0                     # empty bucket: nothing to do
        nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        nccl.ncclCommDestroy(comm)
        out.put((rank, "ok"))
    except Exception as e:          # report instead of hanging the peer
        import traceback
        out.put((rank, traceback.format_exc()))
        raise e


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allreduce_mean_on_a_raw_nccl_communicator():
    ctx = mp.get_context("spawn")
    ids, out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, ids, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, "ok"), (1, "ok")], res
