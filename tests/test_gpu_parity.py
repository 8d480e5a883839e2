"""GPU: the CUDA path (through the C ABI) against the reference's golden vectors and the fp64 oracle.

Tolerances (fp32 exact path vs fp64 reference), max-norm relative to max|ref|:
  outputs H / y : 1e-5      parameter and input gradients : 1e-4
"""
import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg
from oracle import gcrnn_oracle as orc
from tests import _golden as G
from gated_gcrnns_b200 import _lib

pytestmark = pytest.mark.gpu
TOL_OUT, TOL_GRAD = 1e-5, 1e-4
DEV = 'cuda:0'


def relerr(a, b):
    a = np.asarray(a.detach().cpu().double() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu().double() if torch.is_tensor(b) else b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def f32(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device=DEV)


def test_lsigf_golden_e2():
    c = G.load('lsigf_e2')
    S = torch.tensor(c['S'])                                  # fp64 CPU tensor, as addGSO receives it
    x, w, b = f32(c['x']).requires_grad_(True), f32(c['weight']).requires_grad_(True), f32(c['bias']).requires_grad_(True)
    y = gg.LSIGF(w, S, x, b)
    (y * f32(c['dy'])).sum().backward()
    assert relerr(y, c['y']) < TOL_OUT
    assert relerr(x.grad, c['dx']) < TOL_GRAD
    assert relerr(w.grad, c['dweight']) < TOL_GRAD
    assert relerr(b.grad, c['dbias']) < TOL_GRAD


def test_graph_filter_padding_golden():
    c = G.load('lsigf_e2')
    gf = gg.GraphFilter(3, 4, 3, 2, True).to(DEV)
    gf.addGSO(torch.tensor(c['S']))
    with torch.no_grad():
        gf.weight.copy_(f32(c['weight'])); gf.bias.copy_(f32(c['bias']))
    assert relerr(gf(f32(c['x_short'])), c['y_short']) < TOL_OUT       # Nin < N zero padding (graphML.py:1181-1194)


def test_gat_golden():
    c = G.load('gat')
    ga = gg.GraphAttentional(5, 5, 1).to(DEV)
    ga.addGSO(torch.tensor(c['S']))
    with torch.no_grad():
        ga.mixer.copy_(f32(c['mixer'])); ga.weight.copy_(f32(c['weight']))
    x = f32(c['x']).requires_grad_(True)
    y = ga(x)
    (y * f32(c['dy'])).sum().backward()
    assert relerr(y, c['y']) < TOL_OUT
    assert relerr(x.grad, c['dx']) < TOL_GRAD
    assert relerr(ga.mixer.grad, c['dmixer']) < TOL_GRAD
    assert relerr(ga.weight.grad, c['dweight']) < TOL_GRAD


def build_cell(m, S, params, dtype=torch.float32):
    cell = gg.GGCRNNCell(m['G'], m['F'], m['Kin'], m['Kst'], torch.tanh, m['time_gating'], m['spatial_gating'], m['E'],
                         m['bias'])
    cell.addGSO(S)
    cell.load_state_dict({k: torch.tensor(v) for k, v in params.items()})
    return cell.to(device=DEV, dtype=dtype)


@pytest.mark.parametrize('name', G.names('cell_'))
def test_cell_golden(name):
    c = G.load(name)
    m = G.cell_meta(c)
    cell = build_cell(m, torch.tensor(c['S']), c['param'])
    X, h0 = f32(c['X']).requires_grad_(True), f32(c['h0']).requires_grad_(True)
    H = cell(X, h0)
    assert H.shape == c['H'].shape and H.is_contiguous()
    (H * f32(c['dH'])).sum().backward()
    errs = {'H': relerr(H, c['H']), 'dX': relerr(X.grad, c['dX']), 'dh0': relerr(h0.grad, c['dh0'])}
    named = dict(cell.named_parameters())
    for k, ref in c['grad'].items():
        if ref.size == 0:
            assert named[k].grad is None, f'{k}: the reference leaves this gradient None'
        else:
            errs[k] = relerr(named[k].grad, ref)
    bad = {k: v for k, v in errs.items() if v > (TOL_OUT if k == 'H' else TOL_GRAD)}
    assert not bad, f'{name}: {bad}'


def test_cell_fp64_module_roundtrip_dtype():
    """Scripts run in float64 (kStepPredGRNNs.py:44): fp64 params/inputs in, fp64 out, fp64 grads."""
    c = G.load('cell_small_t1_node')
    m = G.cell_meta(c)
    cell = build_cell(m, torch.tensor(c['S']), c['param'], torch.float64)
    X = torch.tensor(c['X'], device=DEV)
    h0 = torch.tensor(c['h0'])                                  # CPU h0, like train_rnn.py:256
    H = cell(X, h0)
    assert H.dtype == torch.float64 and H.device.type == 'cuda'
    H.sum().backward()
    assert cell.weight_B.grad.dtype == torch.float64
    assert relerr(H, c['H']) < TOL_OUT


@pytest.mark.parametrize('tg,sg', [(True, None), (False, 'node'), (True, 'edge')])
def test_cell_vs_oracle_seeded(tg, sg):
    """Seeded random case (sizes the oracle finishes in seconds), sparse-ish non-symmetric S, G=3."""
    torch.manual_seed(7)
    N, G_, F_, Kin, Kst, T, B = 37, 3, 6, 4, 3, 6, 5
    S = (torch.rand(1, N, N) * (torch.rand(1, N, N) < 0.15)).double()
    S = S / torch.linalg.eigvals(S[0]).abs().max()
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
    try:
        p = orc.init_cell_params(G_, F_, Kin, Kst, N, tg, sg, 1, True)
    finally:
        torch.set_default_dtype(prev)
    X, h0, dH = torch.randn(B, T, G_, N).double(), 0.3 * torch.randn(B, F_, N).double(), torch.randn(B, T, F_, N).double()
    Href, gref = orc.cell_forward_backward(p, S, X, h0, dH, tg, sg, input_grads=True)
    m = dict(G=G_, F=F_, Kin=Kin, Kst=Kst, time_gating=tg, spatial_gating=sg, E=1, bias=True)
    cell = build_cell(m, S, {k: v.numpy() for k, v in p.items()})
    Xg, hg = X.float().to(DEV).requires_grad_(True), h0.float().to(DEV).requires_grad_(True)
    H = cell(Xg, hg)
    (H * dH.float().to(DEV)).sum().backward()
    assert relerr(H, Href) < TOL_OUT
    assert relerr(Xg.grad, gref['__X']) < TOL_GRAD and relerr(hg.grad, gref['__h0']) < TOL_GRAD
    for k, v in cell.named_parameters():
        if gref[k] is None:
            assert v.grad is None
        else:
            assert relerr(v.grad, gref[k]) < TOL_GRAD, k


def test_sparse_gso_equals_dense_gso():
    """The same graph handed over as a torch sparse tensor (large-graph entry) gives identical results."""
    c = G.load('cell_cfg2_edge')
    m = G.cell_meta(c)
    S = torch.tensor(c['S'])
    a = build_cell(m, S, c['param'])
    b = build_cell(m, S[0].float().to_sparse_csr(), c['param'])
    X, h0 = f32(c['X']), f32(c['h0'])
    with torch.no_grad():
        assert torch.equal(a(X, h0), b(X, h0))


def test_linearity_of_lsigf_large():
    """Size-independent property at a larger size: LSIGF is linear in x and in the taps."""
    torch.manual_seed(0)
    N, B, G_, F_, K = 2000, 16, 8, 16, 4
    idx = torch.randint(0, N, (2, 12 * N))
    S = torch.sparse_coo_tensor(idx, torch.rand(12 * N) / 12, (N, N)).coalesce()
    g = gg.graph.from_sparse_tensor(S, DEV)
    h = torch.randn(F_, 1, K, G_, device=DEV)
    x1, x2 = torch.randn(B, G_, N, device=DEV), torch.randn(B, G_, N, device=DEV)
    y = gg.LSIGF(h, g, 2 * x1 - 3 * x2)
    y12 = 2 * gg.LSIGF(h, g, x1) - 3 * gg.LSIGF(h, g, x2)
    assert relerr(y, y12) < 1e-5
    # against dense torch fp64 on the same operator
    Sd = S.to_dense().double()
    yref = orc.lsigf(h.cpu().double(), Sd.reshape(1, N, N), x1.cpu().double())
    assert relerr(gg.LSIGF(h, g, x1), yref) < TOL_OUT


def test_graph_replay_matches_direct_launches():
    """Small fp32 calls are captured into CUDA graphs keyed by their pointer set (csrc/api.cu): replays must launch the same
    kernels and give bit-identical results to direct launches, also when the data behind the pointers changes."""
    from gated_gcrnns_b200 import _lib
    L = _lib.lib()
    c = G.load('cell_cfg2_edge')
    m = G.cell_meta(c)
    cell = build_cell(m, torch.tensor(c['S']), c['param'])
    X, h0, dH = f32(c['X']), f32(c['h0']), f32(c['dH'])

    def run(x):
        cell.zero_grad()
        l0 = L.gcrnn_debug_launch_count()
        H = cell(x, h0)
        (H * dH).sum().backward()
        torch.cuda.synchronize()
        return H.detach().clone(), cell.weight_B.grad.clone(), L.gcrnn_debug_launch_count() - l0

    old = gg.options.set('graph_capture', 0)
    try:
        Hd, gd, ld = run(X)
        Hd2, gd2, _ = run(2 * X)
    finally:
        gg.options.set('graph_capture', old)
    assert gg.options.set('graph_capture', 1) in (0, 1)
    outs = [run(X), run(2 * X), run(X), run(X)]
    assert all(o[2] == ld for o in outs), [o[2] for o in outs]
    assert torch.equal(outs[0][0], Hd) and torch.equal(outs[2][0], Hd) and torch.equal(outs[3][0], Hd)
    assert torch.equal(outs[1][0], Hd2)
    assert relerr(outs[3][1], gd) < 1e-6 and relerr(outs[1][1], gd2) < 1e-6      # float atomics: order may differ in the last bits
    gg.options.set('graph_capture', old)


# ---------------------------------------------------------------------------------------------------------------
# persistent fused recurrence for small graphs (csrc/persist_f32.cuh, GCRNN_PATH_PERSIST): one launch per direction
# ---------------------------------------------------------------------------------------------------------------
PATH_PERSIST = 2


def _last_path(cell):
    h = next(iter(cell._handles.values()))
    return h.get_option('last_path')


@pytest.mark.parametrize('name', G.names('cell_'))
def test_persistent_path_golden(name):
    """The reference's own outputs and gradients (fixtures generated from the unmodified reference) on the persistent path:
    X does not require grad here (as in the reference's training loops), so the library picks it for every gating mode (none,
    time, node, edge, time + node / edge) with one edge feature."""
    c = G.load(name)
    m = G.cell_meta(c)
    if m['E'] != 1:
        pytest.skip('E > 1 takes the per-op kernels')
    cell = build_cell(m, torch.tensor(c['S']), c['param'])
    X, h0 = f32(c['X']), f32(c['h0']).requires_grad_(True)
    l0 = _lib.lib().gcrnn_debug_launch_count()
    H = cell(X, h0)
    assert _last_path(cell) == PATH_PERSIST, 'expected the persistent small-graph path'
    (H * f32(c['dH'])).sum().backward()
    torch.cuda.synchronize()
    assert _lib.lib().gcrnn_debug_launch_count() - l0 == 2          # ONE launch forward, ONE launch backward
    errs = {'H': relerr(H, c['H']), 'dh0': relerr(h0.grad, c['dh0'])}
    named = dict(cell.named_parameters())
    for k, ref in c['grad'].items():
        if ref.size == 0:
            assert named[k].grad is None, f'{k}: the reference leaves this gradient None'
        else:
            errs[k] = relerr(named[k].grad, ref)
    bad = {k: v for k, v in errs.items() if v > (TOL_OUT if k == 'H' else TOL_GRAD)}
    assert not bad, f'{name}: {bad}'


@pytest.mark.parametrize('tg,sg', [(False, None), (True, None), (False, 'node'), (True, 'node'), (False, 'edge'), (True, 'edge')])
@pytest.mark.parametrize('N,G_,F_,Kin,Kst,T,B,bias', [(37, 3, 6, 4, 3, 6, 5, True), (80, 1, 20, 5, 5, 5, 7, True), (59, 2, 8, 1, 4, 3, 4, False),
                                                     (16, 1, 4, 2, 1, 2, 3, True)])
def test_persistent_path_matches_per_op_kernels(tg, sg, N, G_, F_, Kin, Kst, T, B, bias):
    """Same cell, same inputs: persistent kernels vs the generic per-op kernels (option persist = 0), including a non-symmetric
    operator, G > 1, Kin != Kst, K = 1, no bias, h0 != 0, and the last-state-only gradient."""
    torch.manual_seed(N + T)
    S = torch.rand(1, N, N) * (torch.rand(1, N, N) < 0.2)
    S = S / torch.linalg.eigvals(S[0]).abs().max()
    X, h0 = torch.randn(B, T, G_, N, device=DEV), 0.3 * torch.randn(B, F_, N, device=DEV)
    dH, dHl = torch.randn(B, T, F_, N, device=DEV), torch.randn(B, F_, N, device=DEV)
    out = {}
    old = gg.options.set('persist', 1)
    try:
        for persist in (0, 1):
            gg.options.set('persist', persist)
            torch.manual_seed(1)
            cell = gg.GGCRNNCell(G_, F_, Kin, Kst, torch.tanh, tg, sg, 1, bias)
            cell.addGSO(S)
            cell = cell.to(DEV)
            hh = h0.clone().requires_grad_(True)
            H = cell(X, hh)
            fits = not (tg and sg and N == 80)          # time + node gates at N = 80, F = 20, K = 5 exceed one SM's 227 KB: per-op kernels
            assert (_last_path(cell) == PATH_PERSIST) == bool(persist and fits)
            (H * dH).sum().backward()
            res = [H.detach(), hh.grad.clone()] + [p.grad.clone() for p in cell.parameters() if p.grad is not None]
            cell.zero_grad()
            cell.last_state_only = True
            hh2 = h0.clone().requires_grad_(True)
            Hl = cell(X, hh2)
            (Hl.select(1, -1) * dHl).sum().backward()
            res += [hh2.grad.clone()] + [p.grad.clone() for p in cell.parameters() if p.grad is not None]
            out[persist] = res
    finally:
        gg.options.set('persist', old)
    assert len(out[0]) == len(out[1])
    for i, (a, b) in enumerate(zip(out[1], out[0])):
        assert relerr(a, b) < (TOL_OUT if i == 0 else TOL_GRAD), (i, relerr(a, b))
