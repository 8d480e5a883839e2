"""CPU, world_size = 2 over gloo: the data-parallel host logic (batch sharding, fixed-layout gradient bucket,
one all-reduce per step).  The per-rank gradients come from the fp64 oracle (no GPU here); the check is that
sum over ranks of shard gradients == gradients of the whole batch, including parameters whose gradient is None."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gated_gcrnns_b200 import dist as gdist
from oracle import gcrnn_oracle as orc


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _problem():
    torch.manual_seed(0)
    N, G, F, K, T, B = 10, 1, 3, 3, 4, 5            # B = 5 over 2 ranks: ragged shards (3 + 2)
    S = (torch.rand(1, N, N) * (torch.rand(1, N, N) < 0.4)).double()
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
    try:
        p = orc.init_cell_params(G, F, K, K, N, True, 'node', 1, True)
    finally:
        torch.set_default_dtype(prev)
    X, h0, dH = torch.randn(B, T, G, N).double(), torch.randn(B, F, N).double(), torch.randn(B, T, F, N).double()
    return p, S, X, h0, dH, B


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        p, S, X, h0, dH, B = _problem()
        lo, hi = gdist.shard_range(B, rank, world)
        _, g = orc.cell_forward_backward(p, S, X[lo:hi], h0[lo:hi], dH[lo:hi], True, 'node')
        names = list(p.keys())
        bucket = gdist.flatten_grads([g[n] for n in names], [p[n] for n in names])      # None -> zeros, fixed layout
        gdist.enable()
        gdist.allreduce_bucket(bucket)
        gdist.disable()
        if rank == 0:
            q.put((names, bucket))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for B in (0, 1, 5, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [gdist.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_sum_equals_full_batch():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    names, bucket = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, S, X, h0, dH, B = _problem()
    _, g = orc.cell_forward_backward(p, S, X, h0, dH, True, 'node')
    full = gdist.flatten_grads([g[n] for n in names], [p[n] for n in names])
    assert bucket.shape == full.shape
    assert (bucket.double() - full.double()).abs().max() <= 1e-5 * full.abs().max()
    # the never-used GFL_out / MLP_out parameters occupy zero-filled slots of the bucket (layout identical on all ranks)
    parts = gdist.unflatten(bucket, [p[n] for n in names])
    for n, v in zip(names, parts):
        if g[n] is None:
            assert float(v.abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------
# whole-model reduction: attach(model, 'mean') / allreduce_gradients reduce EVERY parameter once per backward
# (the cell's own in-backward all-reduce covers the cell's parameters only)
# ---------------------------------------------------------------------------------------------------------------
def _model_and_data():
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    unused = torch.nn.Linear(3, 3)                     # registered but never used in forward: its .grad stays None
    model.add_module('never_used', unused)
    X, Y = torch.randn(10, 6), torch.randn(10, 2)
    return model, X, Y


def _fwd(model, x):
    return model[2](model[1](model[0](x)))


def _attach_worker(rank, world, port, q, mode):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model, X, Y = _model_and_data()
        lo, hi = gdist.shard_range(X.shape[0], rank, world)       # 5 + 5
        gdist.enable()
        n0 = gdist.launches
        if mode == 'attach':
            att = gdist.attach(model, op='mean')
            for _ in range(2):                                     # two steps: the hook re-arms after every backward
                model.zero_grad()
                torch.nn.functional.mse_loss(_fwd(model, X[lo:hi]), Y[lo:hi]).backward()
            att.detach()
            used = gdist.launches - n0
        else:
            model.zero_grad()
            for a, b in ((lo, lo + 2), (lo + 2, hi)):              # gradient accumulation over two micro-batches
                (torch.nn.functional.mse_loss(_fwd(model, X[a:b]), Y[a:b], reduction='sum') / (X.shape[0] * Y.shape[1])).backward()
            gdist.allreduce_gradients(model.parameters(), op='sum')
            used = gdist.launches - n0
        gdist.disable()
        if rank == 0:
            q.put(([None if p.grad is None else p.grad.clone() for p in model.parameters()], used))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize('mode', ['attach', 'explicit'])
def test_whole_model_gradients_match_single_process(mode):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_attach_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for pr in procs:
        pr.start()
    grads, used = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    model, X, Y = _model_and_data()
    torch.nn.functional.mse_loss(_fwd(model, X), Y).backward()     # single process, whole batch, mean loss
    assert used == (2 if mode == 'attach' else 1)                  # ONE collective per backward / per step
    for g, p in zip(grads, model.parameters()):
        if p.grad is None:
            assert g is None                                       # never-used parameters keep None, as in the reference
        else:
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7)
