"""GPU: the library's own NCCL transport (gcrnn_comm_* / gcrnn_allreduce_sum, resolved from the process's libnccl) on a
single-rank communicator, and the opt-in in-backward gradient all-reduce of the cell (identity at world size 1, but it runs the
whole path: unique id -> ncclCommInitRank -> ncclAllReduce on the backward stream).  The 2-rank arithmetic is covered on
CPU by tests/test_dist_gloo.py and on 2 GPUs by tests/test_gpu_dist2.py (skipped on a 1-GPU box); bench.py prints `grad_check`."""
import os

import pytest
import torch

import gated_gcrnns_b200 as gg

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_native_allreduce_single_rank_and_cell_hook():
    import torch.distributed as dist
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29577')
    torch.cuda.set_device(0)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        gg.dist.enable(native=True, device=0, reduce_in_backward=True)
        b = torch.arange(1000, dtype=torch.float32, device=DEV)
        out = gg.dist.allreduce_bucket(b.clone())
        torch.cuda.synchronize()
        assert torch.equal(out, b)
        # the cell's backward issues exactly one all-reduce per step on its flat gradient bucket
        gg.set_precision('fp32')
        S = gg.graphs.sbm(40, 4, 0.7, 0.1, seed=3)
        torch.manual_seed(0)
        cell = gg.GGCRNNCell(1, 8, 3, 3, torch.tanh, True, None, 1, True)
        cell.addGSO(S)
        cell = cell.to(DEV)
        X, h0 = torch.randn(4, 3, 1, 40, device=DEV), torch.zeros(4, 8, 40, device=DEV)
        n0 = gg.dist.launches
        cell(X, h0).sum().backward()
        torch.cuda.synchronize()
        assert gg.dist.launches == n0 + 1
        g_on = cell.weight_B.grad.clone()
        gg.dist.disable()
        cell.zero_grad()
        cell(X, h0).sum().backward()
        assert torch.allclose(cell.weight_B.grad, g_on, rtol=1e-4, atol=1e-4)      # float atomics: summation order varies
    finally:
        gg.dist.disable()
        if created:
            dist.destroy_process_group()
