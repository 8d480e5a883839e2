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


def _set_opt(name, value):
    return _lib.lib().gcrnn_debug_set_option(name.encode(), int(value))


@pytest.mark.parametrize('pair', [0, 1])
@pytest.mark.parametrize('N,M', [(256, 128), (128, 300), (1024, 128 * 5 + 17), (512, 4096), (1024, 256 * 75 + 130)])
def test_shift_gemm_matches_torch(N, M, pair):
    """The library keeps the operator as bf16(S / max|S|) and multiplies by max|S| in the fp32 epilogue (exact for
    unweighted graphs).  Against an fp32 matmul with that same effective operator the only difference is accumulation
    order: tolerance 1e-5 relative to max|ref| (fp32 out), 2^-8 for the bf16 copy."""
    torch.manual_seed(N + M)
    S = torch.randn(N, N) * (torch.rand(N, N) < 0.3) / 16
    g = ggraph.from_dense(S.reshape(1, N, N), DEV, keep_dense=True)
    mx = S.abs().max()
    S = (S / mx).to(torch.bfloat16).float() * mx
    A = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    old = _set_opt('gemm_pair', pair)          # 1: CTA-pair (cta_group::2) kernel where the shape allows it
    try:
        for backward in (False, True):
            ob, of = _gemm(g, A, backward)
            Sd = S.to(DEV)
            ref = A.float() @ (Sd.t() if backward else Sd)
            scale = ref.abs().max().item()
            assert (of - ref).abs().max().item() / scale < 1e-5, (N, M, backward)
            assert (ob.float() - ref).abs().max().item() / scale < 2 ** -8
    finally:
        _set_opt('gemm_pair', old)


# ---------------------------------------------------------------------------------------------------------
# the tensor-core cell path.  Stated bound for bf16 operands (8-bit mantissa) with fp32 accumulation, fp32
# state and fp32 gates, max-norm error relative to max|ref|:
#   one step (T = 1, "teacher forced")        : H <= 1e-2, parameter gradients <= 3e-2
#   short horizons (T <= 6), reference init   : H <= 1e-1, parameter gradients <= 6e-2
# With the reference initialisation the state map has gain > 1 (weight_B ~ U(+-1/sqrt(G*Kin)) over F inputs and an
# eigenvalue-1 common mode of S), so any rounding difference is amplified step by step; see also the contractive
# long-horizon test below.  The fp32 path's 1e-5 / 1e-4 bound does NOT apply here.
# ---------------------------------------------------------------------------------------------------------
TC_TOL_H, TC_TOL_G = 1e-1, 6e-2
TC_TOL_H1, TC_TOL_G1 = 1e-2, 3e-2


def _log(*a):
    import os
    line = ' '.join(str(x) for x in a)
    print(line)
    if os.path.isdir('gpurun_out'):
        with open('gpurun_out/tc_errors.log', 'a') as f:
            f.write(line + '\n')


def _relerr(a, b, scale=None):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max() if scale is None else torch.maximum(b.abs().max(), scale.detach().double().cpu().abs().max())
    return ((a - b).abs().max() / den.clamp_min(1e-30)).item()


def _grad_errs(got, ref):
    """max-norm relative errors per parameter.  A 1-element gradient (the gate MLP bias: one signed sum over all
    (b, t), heavily cancelling) has no scale of its own: it is measured against the largest entry of the same Linear
    layer's weight gradient."""
    errs = {}
    for k in ref:
        if ref[k] is None:
            assert got[k] is None, k
            continue
        scale = ref.get(k.replace('.bias', '.weight')) if (ref[k].numel() == 1 and k.endswith('.0.bias')) else None
        errs[k] = _relerr(got[k], ref[k], scale)
    return errs


def _make_cell(S, G, F, K, tg, prec, seed=0):
    torch.manual_seed(seed)
    cell = gg.GGCRNNCell(G, F, K, K, torch.tanh, tg, None, 1, True)
    cell.addGSO(S)
    gg.set_precision(prec)
    return cell.to(DEV)


@pytest.mark.parametrize('tg', [False, True])
@pytest.mark.parametrize('N,F,K,T,B,G', [(256, 32, 3, 5, 8, 1), (128, 16, 4, 3, 5, 2), (512, 64, 5, 4, 6, 1),
                                         (256, 32, 3, 1, 8, 1), (512, 64, 5, 1, 6, 1)])
def test_tc_cell_matches_fp32_path(tg, N, F, K, T, B, G):
    S = gg.graphs.dense_random(N, 0.3, seed=1)
    torch.manual_seed(3)
    X, h0, dH = torch.randn(B, T, G, N, device=DEV), 0.3 * torch.randn(B, F, N, device=DEV), torch.randn(B, T, F, N, device=DEV)
    out = {}
    try:
        for prec in ('fp32', 'bf16'):
            cell = _make_cell(S, G, F, K, tg, prec)
            hh = h0.clone().requires_grad_(True)
            H = cell(X, hh)
            (H * dH).sum().backward()
            out[prec] = (H.detach(), {k: v.grad for k, v in cell.named_parameters()}, hh.grad)
    finally:
        gg.set_precision('fp32')
    H32, g32, dh32 = out['fp32']
    Hb, gb, dhb = out['bf16']
    errs = {'H': _relerr(Hb, H32), 'dh0': _relerr(dhb, dh32)}
    errs.update(_grad_errs(gb, g32))
    _log('tc-vs-fp32', dict(tg=tg, N=N, F=F, K=K, T=T, B=B, G=G), {k: f'{v:.2e}' for k, v in errs.items()})
    tol_h, tol_g = (TC_TOL_H1, TC_TOL_G1) if T == 1 else (TC_TOL_H, TC_TOL_G)
    assert errs['H'] < tol_h, errs
    bad = {k: v for k, v in errs.items() if k != 'H' and v > tol_g}
    assert not bad, errs


@pytest.mark.parametrize('tg', [False, True])
@pytest.mark.parametrize('N,K,T,B,G', [(512, 5, 4, 6, 1), (256, 4, 3, 5, 2), (1024, 2, 2, 3, 1)])
def test_tc_fused_backward_step_matches_unfused(tg, N, K, T, B, G):
    """F = 64: the fused reverse-time kernel (tensor-core weight/tap gradients, tc_bwd.cuh) against the separate
    tap-contraction + wgrad kernels.  Both are bf16-operand paths with the same forward, so they agree much more tightly
    than either does with fp32: 1e-2 of max|ref| on every gradient."""
    F = 64
    S = gg.graphs.dense_random(N, 0.3, seed=4)
    torch.manual_seed(6)
    X, h0, dH = torch.randn(B, T, G, N, device=DEV), 0.3 * torch.randn(B, F, N, device=DEV), torch.randn(B, T, F, N, device=DEV)
    out = {}
    old = _set_opt('bwd_fused', 1)
    try:
        for fused in (0, 1):
            _set_opt('bwd_fused', fused)
            cell = _make_cell(S, G, F, K, tg, 'bf16')
            hh = h0.clone().requires_grad_(True)
            H = cell(X, hh)
            (H * dH).sum().backward()
            out[fused] = ({k: v.grad for k, v in cell.named_parameters()}, hh.grad)
    finally:
        _set_opt('bwd_fused', old)
        gg.set_precision('fp32')
    errs = {'dh0': _relerr(out[1][1], out[0][1])}
    errs.update(_grad_errs(out[1][0], out[0][0]))
    _log('tc-fused-vs-unfused', dict(tg=tg, N=N, K=K, T=T, B=B, G=G), {k: f'{v:.2e}' for k, v in errs.items()})
    assert all(v < 1e-2 for v in errs.values()), errs


def test_tc_cell_vs_fp64_oracle_reduced_cfg3():
    """cfg3's shapes with a small batch: N=1024, F=64, K=5, G=1, time-gated, against the fp64 oracle."""
    from oracle import gcrnn_oracle as orc
    N, F, K, T, B = 1024, 64, 5, 6, 2
    S = gg.graphs.dense_random(N, 0.3, seed=0)
    try:
        cell = _make_cell(S, 1, F, K, True, 'bf16')
        torch.manual_seed(5)
        X, h0, dH = torch.randn(B, T, 1, N), torch.zeros(B, F, N), torch.randn(B, T, F, N)
        p = {k: v.detach().double().cpu() for k, v in cell.state_dict().items()}
        Href, gref = orc.cell_forward_backward(p, S.double(), X.double(), h0.double(), dH.double(), True, None)
        H = cell(X.to(DEV), h0.to(DEV))
        (H * dH.to(DEV)).sum().backward()
    finally:
        gg.set_precision('fp32')
    errs = {'H': _relerr(H, Href)}
    errs.update(_grad_errs({k: v.grad for k, v in cell.named_parameters()}, {k: gref[k] for k, _ in cell.named_parameters()}))
    _log('tc-vs-fp64', {k: f'{v:.2e}' for k, v in errs.items()})
    assert errs['H'] < TC_TOL_H, errs
    assert all(v < TC_TOL_G for k, v in errs.items() if k != 'H'), errs


def test_tc_cell_long_horizon_contractive():
    """With the reference init the state map has gain > 1 (weight_B ~ U(+-1/sqrt(G*Kin)) over F inputs), so ANY
    rounding difference grows with T.  With a contractive recurrence (weight_B scaled by 0.2) the bf16 path must
    stay within 1e-2 of the fp32 path over T = 48 steps."""
    N, F, K, T, B = 256, 32, 4, 48, 4
    S = gg.graphs.dense_random(N, 0.3, seed=2)
    torch.manual_seed(11)
    X, h0 = torch.randn(B, T, 1, N, device=DEV), torch.zeros(B, F, N, device=DEV)
    out = {}
    try:
        for prec in ('fp32', 'bf16'):
            cell = _make_cell(S, 1, F, K, True, prec, seed=9)
            with torch.no_grad():
                cell.weight_B.mul_(0.2)
            H = cell(X, h0)
            H.square().sum().backward()
            out[prec] = (H.detach(), cell.weight_B.grad.clone(), cell.weight_A.grad.clone())
    finally:
        gg.set_precision('fp32')
    e = [_relerr(a, b) for a, b in zip(out['bf16'], out['fp32'])]
    _log('tc-long-horizon', [f'{v:.2e}' for v in e])
    assert e[0] < 1e-2 and e[1] < 3e-2 and e[2] < 3e-2, e


def test_tc_full_size_batch_properties():
    """cfg3 at FULL per-sequence size (N=1024, F=64, K=5, T=64, time-gated) through size-independent properties of the
    recurrence: every sequence is independent of its batch mates (a sample run in a batch of 48 equals the same sample
    in a batch of 16), gradients are additive over a partition of the batch, and |h| <= 1.  The recurrence is made
    contractive (weight_B x 0.2) so that reduction-order differences of the gate sums are not amplified over 64 steps.
    Tolerance 2^-8 = one bf16 ulp: a 1e-7 difference in a gate (summation order depends on the batch split) occasionally flips
    the bf16 rounding of a state element that feeds the next shift GEMM."""
    N, F, K, T, B = 1024, 64, 5, 64, 48
    S = gg.graphs.dense_random(N, 0.3, seed=0)
    try:
        cell = _make_cell(S, 1, F, K, True, 'bf16')
        with torch.no_grad():
            cell.weight_B.mul_(0.2)
        torch.manual_seed(17)
        X, h0 = torch.randn(B, T, 1, N, device=DEV), torch.zeros(B, F, N, device=DEV)
        dH = torch.randn(B, T, F, N, device=DEV) / (T * N)
        res = []
        for lo, hi in ((0, B), (0, 16), (16, B)):
            cell.zero_grad()
            H = cell(X[lo:hi], h0[lo:hi])
            (H * dH[lo:hi]).sum().backward()
            res.append((H.detach(), {k: v.grad.clone() for k, v in cell.named_parameters() if v.grad is not None}))
    finally:
        gg.set_precision('fp32')
    Hall, gall = res[0]
    assert torch.isfinite(Hall).all() and Hall.abs().max() <= 1.0
    eh = (_relerr(res[1][0], Hall[:16]), _relerr(res[2][0], Hall[16:]))
    errs = {k: _relerr(res[1][1][k] + res[2][1][k], gall[k]) for k in gall}
    _log('tc-full-size-additivity', eh, {k: f'{v:.2e}' for k, v in errs.items()})
    assert max(eh) < 2 ** -8, eh
    assert all(v < 2 ** -8 for v in errs.values()), errs
