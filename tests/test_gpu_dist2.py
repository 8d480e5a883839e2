"""GPU, world_size = 2 over NCCL (skipped when the box has fewer than 2 GPUs; run with `gpurun --gpus 2`).

The data-parallel ARITHMETIC on hardware: two ranks each run the fused recurrence + a readout on half of the batch and
reduce every parameter gradient with ONE collective; the result must equal the single-rank gradients of the whole batch.
Covers the three product paths of gated_gcrnns_b200.dist: allreduce_gradients (explicit), attach (hooks for unchanged
training loops) and the native transport (gcrnn_allreduce_sum through the C ABI)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _build(dev, prec):
    import gated_gcrnns_b200 as gg
    gg.set_precision(prec)
    N, F, K = (256, 16, 3) if prec != 'fp32' else (48, 8, 3)
    S = gg.graphs.dense_random(N, 0.3, seed=2) if prec != 'fp32' else gg.graphs.sbm(N, 4, 0.7, 0.1, seed=2)
    torch.manual_seed(0)
    cell = gg.GGCRNNCell(1, F, K, K, torch.tanh, True, None, 1, True)
    cell.addGSO(S)
    readout = torch.nn.Linear(F, 1)
    model = torch.nn.ModuleDict(dict(cell=cell, readout=readout)).to(dev)
    B, T = 8, 4
    g = torch.Generator().manual_seed(7)
    X = torch.randn(B, T, 1, N, generator=g).to(dev)
    Y = torch.randn(B, T, N, generator=g).to(dev)
    h0 = 0.2 * torch.randn(B, F, N, generator=g).to(dev)
    return model, X, Y, h0


def _loss(model, X, Y, h0):
    H = model['cell'](X, h0)                                   # [B,T,F,N]
    y = model['readout'](H.permute(0, 1, 3, 2)).squeeze(-1)    # per-node readout (architectures.py:1613-1636)
    return (y - Y).abs().mean()                                # mean-reduced L1 loss, as miscTools.batchTimeL1Loss


def _worker(rank, world, port, mode, prec, q):
    import torch.distributed as dist
    import gated_gcrnns_b200 as gg
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        model, X, Y, h0 = _build(dev, prec)
        lo, hi = gg.dist.shard_range(X.shape[0], rank, world)
        gg.dist.enable(native=(mode == 'native'))
        n0 = gg.dist.launches
        if mode == 'attach':
            att = gg.dist.attach(model, op='mean')
            _loss(model, X[lo:hi], Y[lo:hi], h0[lo:hi]).backward()
            att.detach()
        else:
            _loss(model, X[lo:hi], Y[lo:hi], h0[lo:hi]).backward()
            gg.dist.allreduce_gradients(model.parameters(), op='mean')
        torch.cuda.synchronize()
        used = gg.dist.launches - n0
        gg.dist.disable()
        if rank == 0:
            q.put(({k: (None if p.grad is None else p.grad.cpu()) for k, p in model.named_parameters()}, used))
    finally:
        gg.set_precision('fp32')
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('mode,prec', [('explicit', 'fp32'), ('attach', 'fp32'), ('native', 'fp32'), ('explicit', 'bf16x2')])
def test_two_gpu_gradients_match_single_gpu(mode, prec):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    import gated_gcrnns_b200 as gg
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, prec, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    grads, used = q.get(timeout=500)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert used == 1, used                                      # ONE collective per step
    try:
        model, X, Y, h0 = _build(torch.device('cuda', 0), prec)
        _loss(model, X, Y, h0).backward()
    finally:
        gg.set_precision('fp32')
    tol = 1e-4 if prec == 'fp32' else 2e-3                      # fp32: float-atomic summation order; bf16x2: its stated one-step bound
    for k, p in model.named_parameters():
        if p.grad is None:
            assert grads[k] is None, k
            continue
        ref = p.grad.cpu()
        err = (grads[k] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err < tol, (k, err)
