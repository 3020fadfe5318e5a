"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) carries the one exchange
step of the path -- the sum of the per-rank partial reduced camera systems [S_upper | b] and of the scalar partial
sums (chi2, step dot-products, initial damping). Everything numeric stays in libspp_b200.so; this module only turns
the library's device pointer into a tensor view and calls all_reduce on it.

The reference has no counterpart (single process, OpenMP only).
"""
from __future__ import annotations

import ctypes

import numpy as np


class _CudaView:
    """Zero-copy view of n doubles at a device pointer (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def make_torch_allreduce(ctx_stream: int, device=None, group=None):
    """Returns fn(ptr, n) summing n device doubles over the ranks with torch.distributed.

    With the NCCL backend the collective is enqueued with the context's stream as torch's current stream: NCCL
    orders itself after the kernels that produced the buffer and the library's next kernels after the collective,
    no host synchronisation involved."""
    import torch
    import torch.distributed as dist

    ext = torch.cuda.ExternalStream(ctx_stream, device=device) if ctx_stream else None

    def fn(ptr: int, n: int):
        t = torch.as_tensor(_CudaView(ptr, n), device=device)
        if ext is not None:
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return fn


def make_host_allreduce(group=None):
    """The same hook for HOST memory (gloo backend): used by the CPU tests of the multi-rank logic."""
    import torch
    import torch.distributed as dist

    def fn(ptr: int, n: int):
        buf = (ctypes.c_double * n).from_address(ptr)
        a = np.frombuffer(buf, dtype=np.float64)
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return fn


def attach_torch_allreduce(ctx, rank: int, world: int, group=None):
    """Installs the NCCL sum hook on a capi.Context; call BEFORE ba_set_graph (the landmark slice of this rank is
    chosen there)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    ctx.set_allreduce(make_torch_allreduce(ctx.stream, dev, group), rank, world)


def attach_nccl(ctx, rank: int, world: int, group=None):
    """The library's own NCCL path (spp_set_nccl): rank 0 creates the unique id, torch.distributed only carries those
    128 bytes to the other ranks (any transport would do); from then on every collective of the path is an
    ncclAllReduce issued by libspp_b200.so itself on the context's stream. Call BEFORE ba_set_graph."""
    import torch.distributed as dist
    from . import capi
    box = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.set_nccl(box[0], rank, world)


def gather_points(ctx, pts_full: np.ndarray, group=None) -> np.ndarray:
    """All ranks receive the full landmark array: every rank owns the slice returned by ctx.ba_get_partition()."""
    import torch
    import torch.distributed as dist
    b, e = ctx.ba_get_partition()
    out = np.zeros_like(pts_full)
    out[b:e] = pts_full[b:e]
    t = torch.from_numpy(out)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
