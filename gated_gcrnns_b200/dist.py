"""Batch data-parallelism for the recurrence: shard sequences across ranks, reduce parameter gradients once per step.

The reference is single-process (no collective call sites).  Every sequence of the batch is independent
through the whole recurrence, so the path shards by batch with ONE collective per step: a reduction of a flat fp32 bucket
holding every parameter gradient.

What reduces what (pick ONE per model; they do not stack):

  * ``allreduce_gradients(params, op)`` — explicit, end of step: flattens the ``.grad`` of EVERY given parameter (cell,
    readout, anything upstream) into one bucket, all-reduces it once (``'sum'`` or ``'mean'``) and writes the result back.
    This is what a loop with gradient accumulation over micro-batches needs (bench.py uses it).
  * ``attach(model, op)`` — for training loops that cannot be edited (``Modules/train_rnn.py:273-276`` has no hook point
    between ``backward()`` and ``optim.step()``): registers autograd hooks so that the same one-bucket reduction over ALL of
    the model's parameters runs automatically at the end of every ``backward()``.
  * ``enable(reduce_in_backward=True)`` — the cell's own backward sums ITS flat gradient bucket (the buffer the CUDA kernels
    accumulate into) before handing gradients to autograd.  It covers the cell's parameters ONLY and is a SUM; readout layers
    or anything else in the model are NOT reduced by it.  Use it for a bare cell; it is off by default and is suspended while
    ``attach`` is active, so gradients are never reduced twice.

Two transports for the collective:
  * ``torch.distributed`` (NCCL on GPUs over NVLink/NVSwitch, gloo in the CPU tests) — default;
  * the library's own ``gcrnn_allreduce_sum`` (ncclAllReduce enqueued on the current stream through the C ABI, communicator
    bootstrapped from a unique id broadcast over the torch process group): ``enable(native=True)``.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

_group = None          # torch.distributed process group (None = default group)
_native = None         # gcrnn_comm* when the C-ABI transport is active
_enabled = False
_in_backward = False   # the cell's backward reduces its own bucket (cell parameters only, SUM)
_attached = 0          # number of live attach() registrations (suspends _in_backward)
launches = 0           # all-reduces issued (for bench accounting)


def enable(group=None, native: bool = False, device: Optional[int] = None, reduce_in_backward: bool = False):
    """Set up the process group / transport for the gradient reduction.  Call after ``init_process_group``.

    ``reduce_in_backward=True`` additionally makes every cell backward SUM its own gradient bucket over the ranks (cell
    parameters only — see the module docstring)."""
    global _group, _enabled, _native, _in_backward
    import torch.distributed as dist
    assert dist.is_initialized(), 'call torch.distributed.init_process_group first'
    _group, _enabled, _in_backward = group, True, bool(reduce_in_backward)
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
    global _group, _enabled, _native, _in_backward
    if _native is not None:
        from . import _lib
        _lib.lib().gcrnn_comm_destroy(C.c_void_p(_native))
    _group, _enabled, _native, _in_backward = None, False, None, False


def is_enabled() -> bool:
    return _enabled


def world_size() -> int:
    if not _enabled:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(_group)


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


def reduce_cell_bucket(bucket: torch.Tensor):
    """Called by the cell's backward on its flat gradient bucket: reduces only in ``reduce_in_backward`` mode and never while
    an ``attach`` registration owns the model's gradients."""
    if _enabled and _in_backward and _attached == 0:
        allreduce_bucket(bucket)
    return bucket


def allreduce_gradients(params: Iterable[torch.nn.Parameter], op: str = 'sum') -> Optional[torch.Tensor]:
    """Reduce the ``.grad`` of every parameter in ``params`` over the ranks with ONE collective on one flat fp32 bucket.

    ``op='sum'``: gradient of the summed per-rank losses; ``op='mean'``: divided by the world size (what a mean-reduced loss
    over the global batch needs when every rank holds ``B / world`` samples).  Parameters whose ``.grad`` is ``None`` (never used
    in forward: ``GFL_out.*`` / ``MLP_out.*``) contribute zeros so that the bucket layout is identical on every rank, and keep
    ``None``.  Returns the reduced bucket (a view source of nothing: the results are copied back into each ``.grad``)."""
    assert op in ('sum', 'mean')
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return None
    bucket = flatten_grads([p.grad for p in ps], ps)
    allreduce_bucket(bucket)
    if op == 'mean':
        bucket.div_(world_size())
    for p, v in zip(ps, unflatten(bucket, ps)):
        if p.grad is not None:
            p.grad.copy_(v.to(p.grad.dtype))
    return bucket


class _Attachment:
    """Autograd hooks that run ``allreduce_gradients(model.parameters(), op)`` once at the end of every backward pass."""

    def __init__(self, model: torch.nn.Module, op: str):
        global _attached
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.op = op
        self._queued = False
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        _attached += 1

    def _hook(self, _param):
        if not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)

    def _finish(self):
        self._queued = False
        allreduce_gradients(self.params, self.op)

    def detach(self):
        global _attached
        for h in self.handles:
            h.remove()
        if self.handles:
            _attached -= 1
        self.handles = []


def attach(model: torch.nn.Module, op: str = 'mean') -> _Attachment:
    """Make every ``backward()`` through ``model`` end with one all-reduce of ALL its parameter gradients (see the module
    docstring).  ``op='mean'`` matches the reference's mean-reduced losses with the batch sharded evenly over the ranks.
    Returns a handle with ``.detach()``."""
    assert _enabled, 'call gated_gcrnns_b200.dist.enable() first'
    assert op in ('sum', 'mean')
    return _Attachment(model, op)


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
