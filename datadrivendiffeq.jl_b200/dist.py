"""Row-sharded multi-GPU path: one process per GPU, ``torch.distributed`` for the plumbing.

The samples (columns of X / Y) are cut into contiguous ranges, one per rank.
The path has exactly two exchange steps, both plain sum all-reduces of small buffers:

1. the packed normal-equation blocks ``[G | R | yy]`` (``8 (k^2 + k n_t + n_t)`` bytes, 37.9 MB at
   k = 2145, n_t = 64) after each rank's fused Theta+Gram pass;
2. the candidate-rss table (one double per recorded candidate) after each rank's streaming rss pass.

The sweep itself runs replicated on every rank: it is deterministic and all ranks hold bit-identical
blocks after the all-reduce, so no further communication is needed.

Both collectives run INSIDE the library (``csrc/comm.cu``: ``ncclAllReduce`` on the context's compute stream, no
host synchronisation) once the context has a communicator: :func:`init_comm` creates the NCCL unique id on rank
0, carries its 128 bytes to the other ranks over ``torch.distributed`` (the only job left to torch) and calls
``ddb_comm_init_rank``.  Contexts without a communicator (the CPU test double of ``tests/test_dist_gloo.py``)
keep the ``torch.distributed`` all-reduces.
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def shard_range(m: int, rank: int, world: int, align: int = 16) -> Tuple[int, int]:
    """Contiguous sample range ``[lo, hi)`` of ``rank``.  Interior boundaries are multiples of
    ``align`` samples (the Gram kernel's stage size), so only the last rank can own a ragged stage."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = (m + align - 1) // align
    lo = (blocks * rank // world) * align
    hi = (blocks * (rank + 1) // world) * align
    return min(lo, m), min(hi, m)


def _torch():
    import torch
    import torch.distributed as dist

    return torch, dist


def allreduce_sum(t, group=None):
    """Sum all-reduce of a torch tensor over ``group`` (NCCL for CUDA tensors, gloo for CPU)."""
    torch, dist = _torch()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()
    return t


def total_samples(m_local: int, group=None, device="cpu") -> int:
    torch, dist = _torch()
    t = torch.tensor([int(m_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def init_comm(ctx, group=None):
    """Attach an NCCL communicator to ``ctx`` whose ranks are those of the ``torch.distributed`` ``group``."""
    torch, dist = _torch()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    dev = f"cuda:{ctx.device}"
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(uid, src=src, group=group)
    ctx.comm_init_rank(world, rank, bytes(uid.cpu().numpy().tobytes()))


def _has_comm(ctx) -> bool:
    info = getattr(ctx, "comm_info", None)
    return bool(info) and info()[0] > 1


def reduce_blocks_and_sweep(ctx, prm, lambdas: Sequence[float], m_total: int, group=None, paths: bool = False,
                            tensor_device=None, host_data=None):
    """Steps after the data of this rank is bound to ``ctx``: Gram -> all-reduce -> sweep -> rss ->
    all-reduce -> select.  ``ctx`` is a :class:`~.backend.Context` (or a test double with the same
    methods that works on CPU tensors)."""
    if _has_comm(ctx):  # collectives inside the library, on the compute stream
        if host_data is not None and not paths:
            return ctx.solve_sparse_sharded(host_data[0], host_data[1], prm, lambdas, m_total)
        if host_data is not None:
            ctx.upload_gram(host_data[0], host_data[1])
        else:
            ctx.gram()
        ctx.allreduce_gram()
        ctx.sweep(prm, lambdas, m_total)
        ctx.rss()
        ctx.allreduce_rss()
        return ctx.select(paths=paths)
    torch, dist = _torch()
    dev = tensor_device if tensor_device is not None else f"cuda:{ctx.device}"
    if host_data is not None:  # (X_local, Y_local) host arrays: upload overlapped with the Gram kernels
        Xl, Yl = host_data
        ctx.n_t = Yl.shape[0]
        packed = torch.empty(ctx.k * ctx.k + ctx.k * ctx.n_t + ctx.n_t, dtype=torch.float64, device=dev)
        ctx.upload_gram(Xl, Yl, packed.data_ptr())
    else:
        packed = torch.empty(ctx.gram_count(), dtype=torch.float64, device=dev)
        ctx.gram(packed.data_ptr())
    allreduce_sum(packed, group)
    k, n_t = ctx.k, ctx.n_t
    base = packed.data_ptr()
    ctx.gram_upload_ptr(base, base + 8 * k * k, base + 8 * (k * k + k * n_t), n_t)
    ctx.sweep(prm, lambdas, m_total)
    rss = torch.empty(max(1, ctx.rss_count()), dtype=torch.float64, device=dev)
    ctx.rss(rss.data_ptr())
    allreduce_sum(rss, group)
    out = ctx.select(paths=paths)
    del rss, packed
    return out


def solve_sharded(ctx, X_local: np.ndarray, Y_local: np.ndarray, prm, lambdas, group=None, paths: bool = False):
    """Every rank passes its own shard (n x m_local, n_t x m_local); returns ``(SweepOutput, m_total)``."""
    m_total = total_samples(X_local.shape[1], group, device=f"cuda:{ctx.device}")
    if not _has_comm(ctx) and hasattr(ctx, "comm_init_rank"):
        init_comm(ctx, group)
    return reduce_blocks_and_sweep(ctx, prm, lambdas, m_total, group, paths, host_data=(X_local, Y_local)), m_total


def allreduce_host_sum(ctx, values, group=None):
    """Sum over the ranks of a small host array (statistics, counts): through the library's communicator when the
    context has one, else ``torch.distributed``."""
    torch, dist = _torch()
    a = np.ascontiguousarray(values, dtype=np.float64).ravel()
    dev = f"cuda:{ctx.device}" if hasattr(ctx, "comm_info") else "cpu"
    t = torch.from_numpy(a.copy()).to(dev)
    if _has_comm(ctx):
        ctx.allreduce_sum(t.data_ptr(), t.numel())
        ctx.synchronize()
    else:
        allreduce_sum(t, group)
    return t.cpu().numpy().reshape(np.shape(values))


def shard_offsets(m_local: int, group=None, device="cpu"):
    """``(first global sample of this rank, total samples)`` for contiguous row shards in rank order."""
    torch, dist = _torch()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.zeros(world, dtype=torch.int64, device=device)
    t[rank] = int(m_local)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    counts = t.cpu().numpy()
    return int(counts[:rank].sum()), int(counts.sum())

