"""GPU: the tcgen05 dense path.  Tolerances for bf16 operands / fp32 accumulation are stated per test."""
import ctypes as C

import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg
from gated_gcrnns_b200 import _lib, graph as ggraph

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _gemm(g, A_bf16, backward, want_f32=True, want_bf16=True):
    M, N = A_bf16.shape
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=DEV) if want_bf16 else None
    of = torch.empty(M, N, dtype=torch.float32, device=DEV) if want_f32 else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().gcrnn_debug_shift_gemm(g.ptr, int(backward), C.c_void_p(A_bf16.data_ptr()), M,
                                                 C.c_void_p(ob.data_ptr() if ob is not None else 0),
                                                 C.c_void_p(of.data_ptr() if of is not None else 0), st), 'debug_shift_gemm')
    torch.cuda.synchronize()
    return ob, of


@pytest.mark.parametrize('N,M', [(256, 128), (128, 300), (1024, 128 * 5 + 17), (512, 4096)])
def test_shift_gemm_matches_torch(N, M):
    """Operands are exactly representable bf16 values, so the only difference to an fp32 matmul of the same
    operands is accumulation order: tolerance 1e-5 relative to max|ref| (fp32 out), 2^-8 for the bf16 copy."""
    torch.manual_seed(N + M)
    S = (torch.randn(N, N) * (torch.rand(N, N) < 0.3)).to(torch.bfloat16).float() / 16
    g = ggraph.from_dense(S.reshape(1, N, N), DEV, keep_dense=True)
    A = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    for backward in (False, True):
        ob, of = _gemm(g, A, backward)
        Sd = S.to(DEV)
        ref = A.float() @ (Sd.t() if backward else Sd)
        scale = ref.abs().max().item()
        assert (of - ref).abs().max().item() / scale < 1e-5, (N, M, backward)
        assert (ob.float() - ref).abs().max().item() / scale < 2 ** -8
