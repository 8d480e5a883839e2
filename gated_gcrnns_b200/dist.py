"""Batch data-parallelism for the recurrence: shard sequences across ranks, sum parameter gradients.

The reference is single-process (no collective call sites).  Every sequence of the batch
is independent through the whole recurrence, so the path shards by batch with ONE collective per step:
a sum of the flat fp32 parameter-gradient bucket that the backward kernels accumulate into.

Two transports for that one all-reduce:
  * ``torch.distributed`` (NCCL on GPUs over NVLink/NVSwitch, gloo in the CPU tests) — default;
  * the library's own ``gcrnn_allreduce_sum`` (ncclAllReduce enqueued on the backward stream through
    the C ABI, communicator bootstrapped from a unique id broadcast over the torch store).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

_group = None          # torch.distributed process group (or True for the default group)
_native = None         # gcrnn_comm* when the C-ABI transport is active
_enabled = False
launches = 0           # all-reduces issued (for bench accounting)


def enable(group=None, native: bool = False, device: Optional[int] = None):
    """Turn on the gradient all-reduce inside the cell's backward.  Call after init_process_group."""
    global _group, _enabled, _native
    import torch.distributed as dist
    assert dist.is_initialized(), 'call torch.distributed.init_process_group first'
    _group, _enabled = group, True
    if native:
        from . import _lib
        L = _lib.lib()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(L.gcrnn_comm_unique_id(C.cast(buf, C.c_void_p)), 'comm_unique_id')
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        dev = torch.cuda.current_device() if device is None else device
        uid_dev = uid.to(f'cuda:{dev}') if dist.get_backend(group) == 'nccl' else uid
        dist.broadcast(uid_dev, src=0, group=group)
        raw = bytes(uid_dev.cpu().numpy().tobytes())
        out = C.c_void_p()
        _lib.check(L.gcrnn_comm_create(C.byref(out), C.c_char_p(raw), rank, world, dev), 'comm_create')
        _native = out.value


def disable():
    global _group, _enabled, _native
    if _native is not None:
        from . import _lib
        _lib.lib().gcrnn_comm_destroy(C.c_void_p(_native))
    _group, _enabled, _native = None, False, None


def is_enabled() -> bool:
    return _enabled


def allreduce_bucket(bucket: torch.Tensor):
    """Sum ``bucket`` (flat fp32) over the ranks, in place, on the current stream.  No-op when disabled."""
    global launches
    if not _enabled:
        return bucket
    launches += 1
    if _native is not None and bucket.is_cuda:
        from . import _lib
        st = C.c_void_p(torch.cuda.current_stream(bucket.device).cuda_stream)
        _lib.check(_lib.lib().gcrnn_allreduce_sum(C.c_void_p(_native), C.c_void_p(bucket.data_ptr()), bucket.numel(), st),
                   'allreduce_sum')
        return bucket
    import torch.distributed as dist
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=_group)
    return bucket


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of a batch of B sequences for ``rank`` (first B % world ranks get one more)."""
    assert 0 <= rank < world and B >= 0
    q, r = divmod(B, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def flatten_grads(grads: Sequence[Optional[torch.Tensor]], like: Sequence[torch.Tensor]) -> torch.Tensor:
    """Flat fp32 bucket with a FIXED layout on every rank: ``None`` gradients are zero-filled."""
    parts = [(g if g is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for g, p in zip(grads, like)]
    return torch.cat(parts) if parts else torch.zeros(0)


def unflatten(bucket: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    sizes = [p.numel() for p in like]
    return [v.reshape(p.shape) for v, p in zip(torch.split(bucket, sizes), like)]
