"""GPU: the fused F == 32 edge-gated sparse kernels (csrc/sp32_kernels.cuh, the cfg5 path) against the fp64 oracle and
against the generic per-op kernels of the same library.

Tolerances as in test_gpu_parity.py (fp32 exact path vs fp64 reference, max-norm relative to max|ref|):
  H : 1e-5      parameter gradients and dh0 : 1e-4
"""
import ctypes as C

import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg
from gated_gcrnns_b200 import _lib
from oracle import gcrnn_oracle as orc

pytestmark = pytest.mark.gpu
TOL_OUT, TOL_GRAD = 1e-5, 1e-4
DEV = 'cuda:0'


def relerr(a, b):
    a = a.detach().cpu().double().numpy()
    b = b.detach().cpu().double().numpy()
    assert a.shape == b.shape
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def ragged_graph(N, seed, max_out=31, hub=True):
    """Non-symmetric sparse S with out-degrees 0..max_out (row degree of S+I <= 32), one hub column of in-degree > 32,
    a node whose only edge is the added self loop, and one explicit diagonal entry."""
    rng = np.random.RandomState(seed)
    S = np.zeros((N, N))
    for i in range(N):
        deg = rng.randint(0, max_out + 1) if i > 2 else (0, max_out, 16)[i]
        cols = rng.choice(np.delete(np.arange(N), i), size=deg, replace=False)
        S[i, cols] = rng.rand(deg) + 0.1
    if hub:
        rows = rng.choice(np.arange(3, N), size=min(N - 3, 45), replace=False)
        for i in rows:
            if S[i, 7] == 0 and (S[i] != 0).sum() < max_out:
                S[i, 7] = rng.rand() + 0.1
    if (S[5] != 0).sum() < max_out:
        S[5, 5] = 0.37
    S = S / np.abs(np.linalg.eigvals(S)).max()
    return torch.tensor(S).reshape(1, N, N)


PATH_GENERIC, PATH_NODE32 = 0, 1


def run_case(N, G_, Kin, Kst, T, B, bias, seed, need_x=False, expect_path=None, reorder=None):
    torch.manual_seed(seed)
    F_ = 32
    S = ragged_graph(N, seed)
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
    try:
        p = orc.init_cell_params(G_, F_, Kin, Kst, N, False, 'edge', 1, bias)
    finally:
        torch.set_default_dtype(prev)
    X, h0, dH = torch.randn(B, T, G_, N).double(), 0.3 * torch.randn(B, F_, N).double(), torch.randn(B, T, F_, N).double()
    Href, gref = orc.cell_forward_backward(p, S, X, h0, dH, False, 'edge', input_grads=True)
    cell = gg.GGCRNNCell(G_, F_, Kin, Kst, torch.tanh, False, 'edge', 1, bias)
    cell.addGSO(S)
    cell.load_state_dict({k: v for k, v in p.items()})
    cell = cell.to(device=DEV, dtype=torch.float32)
    if reorder is not None:
        gg.graph.get(cell.S, DEV, keep_dense=False).set_option('reorder', reorder)
    Xg = X.float().to(DEV).requires_grad_(need_x)
    hg = h0.float().to(DEV).requires_grad_(True)
    L = _lib.lib()
    l0 = L.gcrnn_debug_launch_count()
    H = cell(Xg, hg)
    (H * dH.float().to(DEV)).sum().backward()
    torch.cuda.synchronize()
    launches = L.gcrnn_debug_launch_count() - l0
    if reorder is not None:
        assert gg.graph.get(cell.S, DEV, keep_dense=False).get_option('reordered') == int(reorder == 2)
    if expect_path is not None:
        took = cell._handle(torch.device(DEV)).get_option('last_path')
        assert took == expect_path, f'forward took path {took}, expected {expect_path}'
    errs = {'H': relerr(H, Href), 'dh0': relerr(hg.grad, gref['__h0'])}
    if need_x:
        errs['dX'] = relerr(Xg.grad, gref['__X'])
    for k, v in cell.named_parameters():
        if gref[k] is None:
            assert v.grad is None
        else:
            errs[k] = relerr(v.grad, gref[k])
    bad = {k: v for k, v in errs.items() if v > (TOL_OUT if k == 'H' else TOL_GRAD)}
    assert not bad, bad
    return launches, H.detach(), {k: v.grad.detach().clone() for k, v in cell.named_parameters() if v.grad is not None}


@pytest.mark.parametrize('G_,Kin,Kst,bias', [(1, 3, 3, True), (2, 4, 2, True), (1, 2, 4, False), (3, 5, 3, True)])
def test_fused_edge32_vs_oracle(G_, Kin, Kst, bias):
    launches, _, _ = run_case(N=150, G_=G_, Kin=Kin, Kst=Kst, T=5, B=3, bias=bias, seed=11 + Kst, expect_path=PATH_NODE32)
    # 4 + (Kst - 2) kernels per forward step and 4 + (Kst - 2) + 1 memset-free kernels per backward step, plus set-up
    assert launches < 5 * (10 + 2 * (Kst - 2)) + 12 + 2 * Kin, f'the fused kernels did not run ({launches} launches)'


@pytest.mark.parametrize('mask', [0, 1, 2, 12, 16, 32])
def test_fused_stage_generations_vs_oracle(mask):
    """The fused path has two generations of every stage (sp32_kernels.cuh: warp per node; sp32_tile.cuh: 8 lanes per node +
    shared-memory tile contractions).  The default (all second generation, mask 63) is what every other test here runs; this one
    holds each stage's second-generation kernel ALONE (bits: 1 spmm, 2 filter, 4 + 8 aggregate and bwd_rows - they share the saved softmax statistics -, 16 bwd_node, 32 dh) and the
    all-first-generation path (0) to the same oracle bounds, on a graph whose N is not a multiple of the tile size."""
    L = _lib.lib()
    old = gg.options.set('sparse_v2', mask)
    try:
        run_case(N=203, G_=2, Kin=3, Kst=3, T=4, B=3, bias=True, seed=21, expect_path=PATH_NODE32)
        if mask in (2, 16, 32):
            run_case(N=97, G_=1, Kin=2, Kst=4, T=3, B=2, bias=True, seed=22, expect_path=PATH_NODE32)
            run_case(N=131, G_=3, Kin=4, Kst=2, T=3, B=2, bias=False, seed=23, expect_path=PATH_NODE32)
    finally:
        gg.options.set('sparse_v2', old)


@pytest.mark.parametrize('tc,bps', [(1, 1), (1, 2), (0, 1), (0, 2)])
def test_fused_tile_kernels_contraction_modes(tc, bps):
    """The tile kernels contract either on tensor cores (3xTF32 mma.sync with error compensation, the default) or with packed
    fp32 FFMA2, at any number of resident blocks per SM (grid size / shared-memory carve-out): both hold the fp32 path's bounds
    (1e-5 on H, 1e-4 on gradients vs the fp64 oracle) for every supported tap count."""
    L = _lib.lib()
    old_bps = gg.options.set('sparse_v2_bps', bps)
    old_tc = gg.options.set('sparse_v2_tc', tc)
    try:
        run_case(N=300, G_=1, Kin=3, Kst=3, T=3, B=2, bias=True, seed=31, expect_path=PATH_NODE32)
        run_case(N=200, G_=2, Kin=2, Kst=4, T=2, B=2, bias=True, seed=32, expect_path=PATH_NODE32)
        run_case(N=170, G_=4, Kin=4, Kst=2, T=2, B=2, bias=True, seed=33, expect_path=PATH_NODE32)
    finally:
        gg.options.set('sparse_v2_bps', old_bps)
        gg.options.set('sparse_v2_tc', old_tc)


def test_fused_dpre_epilogue_matches_separate_kernel():
    """By default the dh kernel of reverse step t finishes step t-1's dpre (dH + dh, tanh', relu masks) in its epilogue; with the
    option off a separate dpre_k runs from a stored dh.  Same oracle bounds either way, and the two agree to rounding."""
    L = _lib.lib()
    _, H1, g1 = run_case(N=260, G_=1, Kin=3, Kst=3, T=6, B=3, bias=True, seed=41, expect_path=PATH_NODE32)
    old = gg.options.set('sparse_v2_fuse_dpre', 0)
    try:
        _, H0, g0 = run_case(N=260, G_=1, Kin=3, Kst=3, T=6, B=3, bias=True, seed=41, expect_path=PATH_NODE32)
    finally:
        gg.options.set('sparse_v2_fuse_dpre', old)
    assert torch.equal(H0, H1)
    for k in g0:
        assert relerr(g1[k], g0[k]) < 1e-5, k


def test_fused_matches_generic_kernels_and_falls_back_for_dX():
    L = _lib.lib()
    old = gg.options.set('sparse_fused', 0)
    try:
        lg, Hg, gg_ = run_case(N=130, G_=1, Kin=3, Kst=3, T=4, B=2, bias=True, seed=5, expect_path=PATH_GENERIC)
    finally:
        gg.options.set('sparse_fused', old)
    lf, Hf, gf = run_case(N=130, G_=1, Kin=3, Kst=3, T=4, B=2, bias=True, seed=5)
    assert lf < lg
    assert relerr(Hf, Hg) < 2e-6
    for k in gg_:
        assert relerr(gf[k], gg_[k]) < 2e-5, k
    # dX requested: fused forward, generic reverse sweep on the prefix of the saved state
    run_case(N=130, G_=2, Kin=3, Kst=3, T=4, B=2, bias=True, seed=6, need_x=True)


@pytest.mark.parametrize('G_,Kin,Kst,bias', [(1, 3, 3, True), (2, 4, 2, True), (1, 2, 4, False)])
def test_library_reorder_vs_oracle(G_, Kin, Kst, bias):
    """The library-owned node renumbering (graph option 'reorder' = 2: forced, the graph is far below the automatic threshold):
    X / h0 / dH gathered, H / dh0 scattered through the permutation; outputs and every gradient against the fp64 oracle in the
    CALLER's node order, on a non-symmetric ragged graph (hub column, isolated node, explicit diagonal)."""
    run_case(N=200, G_=G_, Kin=Kin, Kst=Kst, T=4, B=3, bias=bias, seed=11, expect_path=PATH_NODE32, reorder=2)


def test_library_reorder_decision_and_equivalence():
    """N = 8192 16-NN graph: in the generator's random node order the library renumbers on its own ('reorder' = 1, the default)
    and the results equal those of the un-renumbered run ('reorder' = 0), also for the last-state-only gradient; the same graph
    numbered along a Hilbert curve is left alone."""
    N, F_, K, T, B = 8192, 32, 3, 3, 2
    res = {}
    for mode in (0, 1):
        rp, ci, va, _ = gg.graphs.knn_csr_gpu(N, 16, seed=3, device=DEV, reorder=False)
        S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
        torch.manual_seed(0)
        cell = gg.GGCRNNCell(1, F_, K, K, torch.tanh, False, 'edge', 1, True)
        cell.addGSO(S)
        cell = cell.to(DEV)
        gh = gg.graph.get(cell.S, DEV, keep_dense=False)
        gh.set_option('reorder', mode)
        assert gh.get_option('reordered') == mode
        if mode:
            assert gh.get_option('tile_rows_before_x100') > 1.5 * gh.get_option('tile_rows_after_x100')
        torch.manual_seed(1)
        X = torch.randn(B, T, 1, N, device=DEV)
        h0 = (0.1 * torch.randn(B, F_, N, device=DEV)).requires_grad_(True)
        dH = torch.randn(B, T, F_, N, device=DEV)
        H = cell(X, h0)
        (H * dH).sum().backward()
        out = [H.detach(), h0.grad.clone()] + [p.grad.clone() for p in cell.parameters() if p.grad is not None]
        cell.zero_grad(); cell.last_state_only = True
        h1 = h0.detach().clone().requires_grad_(True)
        Hl = cell(X, h1)
        (Hl.select(1, -1) * dH[:, -1]).sum().backward()
        out += [h1.grad.clone()] + [p.grad.clone() for p in cell.parameters() if p.grad is not None]
        res[mode] = out
    for i, (a, b) in enumerate(zip(res[1], res[0])):
        assert relerr(a, b) < (5e-6 if i == 0 else TOL_GRAD), (i, relerr(a, b))
    rp, ci, va, _ = gg.graphs.knn_csr_gpu(N, 16, seed=3, device=DEV, reorder=True)
    S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
    assert gg.graph.get(S, DEV, keep_dense=False).get_option('reordered') == 0, 'a Hilbert-numbered graph needs no renumbering'


def test_library_reorder_with_input_gradients():
    """A caller that wants dX: the forward then stays in the caller's numbering (dX comes from the generic reverse sweep, which reads
    the forward's saved state), so the results must equal those of a graph whose renumbering is switched off."""
    N, F_, K, T, B = 8192, 32, 3, 2, 2
    res = {}
    for mode in (0, 1):
        rp, ci, va, _ = gg.graphs.knn_csr_gpu(N, 16, seed=5, device=DEV, reorder=False)
        S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
        torch.manual_seed(0)
        cell = gg.GGCRNNCell(1, F_, K, K, torch.tanh, False, 'edge', 1, True)
        cell.addGSO(S)
        cell = cell.to(DEV)
        gg.graph.get(cell.S, DEV, keep_dense=False).set_option('reorder', mode)
        torch.manual_seed(1)
        X = torch.randn(B, T, 1, N, device=DEV, requires_grad=True)
        h0 = (0.1 * torch.randn(B, F_, N, device=DEV)).requires_grad_(True)
        H = cell(X, h0)
        (H * torch.randn(B, T, F_, N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))).sum().backward()
        res[mode] = [H.detach(), X.grad.clone(), h0.grad.clone()] + [p.grad.clone() for p in cell.parameters() if p.grad is not None]
    for i, (a, b) in enumerate(zip(res[1], res[0])):
        assert relerr(a, b) < (5e-6 if i == 0 else TOL_GRAD), (i, relerr(a, b))


def test_fused_large_knn_properties():
    """cfg5-like sizes (N = 20000 here): finite outputs, |h| <= 1, and batch-sample independence (sample b of a
    batch equals the same sample run alone), which a mis-indexed gather or a cross-sample race would break."""
    N, F_, K, T, B = 20000, 32, 3, 3, 3
    rp, ci, va = gg.graphs.knn_csr(N, 16, seed=1, power_iters=10)
    S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
    torch.manual_seed(0)
    cell = gg.GGCRNNCell(1, F_, K, K, torch.tanh, False, 'edge', 1, True)
    cell.addGSO(S)
    cell = cell.to(DEV)
    X = torch.randn(B, T, 1, N, device=DEV)
    h0 = 0.1 * torch.randn(B, F_, N, device=DEV)
    H = cell(X, h0)
    H.sum().backward()
    g_all = cell.weight_B.grad.clone()
    assert torch.isfinite(H).all() and H.abs().max() <= 1.0
    cell.zero_grad()
    acc = torch.zeros_like(g_all)
    for b in range(B):
        cell.zero_grad()
        Hb = cell(X[b:b + 1], h0[b:b + 1])
        assert torch.allclose(Hb[0], H[b].detach(), rtol=0, atol=2e-6)
        Hb.sum().backward()
        acc += cell.weight_B.grad
    assert relerr(acc, g_all) < 1e-4


@pytest.mark.parametrize('N,B,T', [(2000, 3, 3), (100_000, 2, 2)])
def test_cfg5_generator_vs_sparse_fp64_oracle(N, B, T):
    """cfg5's own graph family (directed 16-NN, Hilbert order) at N = 2000 and at the FULL N = 100000, edge-gated,
    F = 32, K = 3, against the sparse fp64 restatement of the reference (the dense reference cannot hold N = 1e5)."""
    F_, K = 32, 3
    rp, ci, va = gg.graphs.knn_csr(N, 16, seed=3, power_iters=10)
    S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
    torch.manual_seed(0)
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
    try:
        p = orc.init_cell_params(1, F_, K, K, N, False, 'edge', 1, True)
    finally:
        torch.set_default_dtype(prev)
    X, h0, dH = torch.randn(B, T, 1, N).double(), 0.1 * torch.randn(B, F_, N).double(), torch.randn(B, T, F_, N).double()
    Href, gref = orc.cell_forward_backward(p, [S.double().to_sparse_coo().coalesce()], X, h0, dH, False, 'edge', input_grads=True)
    cell = gg.GGCRNNCell(1, F_, K, K, torch.tanh, False, 'edge', 1, True)
    cell.addGSO(S)
    cell.load_state_dict(p)
    cell = cell.to(device=DEV, dtype=torch.float32)
    hg = h0.float().to(DEV).requires_grad_(True)
    H = cell(X.float().to(DEV), hg)
    saved = H.grad_fn.saved_tensors[3]              # the forward's saved state (read before backward frees it)
    (H * dH.float().to(DEV)).sum().backward()
    assert cell._handle(torch.device(DEV)).get_option('last_path') == PATH_NODE32
    errs = {'H': relerr(H, Href)}
    for k, v in cell.named_parameters():
        if gref[k] is not None:
            errs[k] = relerr(v.grad, gref[k])
    # dh0 is compared element-wise.  relu / leaky_relu have kinks: with 6.4e6 outputs and 6.8e6 attention edges at N = 1e5 a handful
    # of pre-activations land within fp32 rounding of 0, where fp32 and fp64 legitimately pick different one-sided derivatives;
    # each such event perturbs the gradient in one 2-hop neighbourhood (a few hundred elements, one tight cluster of node indices
    # - tools/dbg_cfg5_parity.py prints them) and shifts the parameter gradients, which sum over all nodes, by ~1e-3 of their scale.
    # So: all but 1e-3 of the dh0 elements must meet the 1e-4 bound, and at full size the parameter gradients get 2e-3.
    ref = gref['__h0']
    d = (hg.grad.detach().cpu().double() - ref).abs() / ref.abs().max()
    frac_bad = float((d > TOL_GRAD).double().mean())
    errs['dh0_frac_above_tol'] = frac_bad
    tol_g = TOL_GRAD if N <= 2000 else 2e-3
    assert errs['H'] < TOL_OUT, errs
    assert frac_bad < (1e-9 if N <= 2000 else 1e-3), errs
    bad = {k: v for k, v in errs.items() if k not in ('H', 'dh0_frac_above_tol') and v > tol_g}
    assert not bad, errs
    if N > 2000:
        # Proof that the loosened bounds are the kinks and nothing else: the fp64 oracle evaluated with EXACTLY the one-sided ReLU
        # derivatives the GPU picked (gcrnn_debug_edge_relu_masks decodes them from the forward's saved state).  The forward value
        # changes only where |y| is at rounding level; every gradient must then meet the un-loosened bounds.
        masks = torch.empty(2, B, T, F_, N, dtype=torch.uint8, device=DEV)
        _lib.check(_lib.lib().gcrnn_debug_edge_relu_masks(cell._handle(torch.device(DEV)).ptr, C.c_void_p(saved.data_ptr()), saved.numel(),
                                                          B, T, C.c_void_p(masks.data_ptr()), None), 'debug_edge_relu_masks')
        mk = masks.cpu().double()
        calls = []

        def relu_with_gpu_decisions(y):                 # the oracle evaluates input gate then forget gate, step after step
            c = len(calls); calls.append(c)
            m = mk[c % 2, :, c // 2]
            flips = int(((y.detach() > 0).double() != m).sum())
            calls[-1] = flips
            return y * m
        orc.RELU = relu_with_gpu_decisions
        try:
            _, gk = orc.cell_forward_backward(p, [S.double().to_sparse_coo().coalesce()], X, h0, dH, False, 'edge', input_grads=True)
        finally:
            orc.RELU = torch.relu
        assert len(calls) == 2 * T
        tight = {k: relerr(v.grad, gk[k]) for k, v in cell.named_parameters() if gk[k] is not None}
        dk = (hg.grad.detach().cpu().double() - gk['__h0']).abs() / gk['__h0'].abs().max()
        tight['dh0_max'] = float(dk.max())
        tight['relu_decisions_that_differ_from_fp64'] = int(sum(calls))
        import json, os
        os.makedirs('gpurun_out', exist_ok=True)
        with open('gpurun_out/cfg5_kink_proof.json', 'w') as f:
            json.dump(dict(N=N, B=B, T=T, gpu_vs_fp64_oracle=errs, gpu_vs_fp64_oracle_with_gpu_relu_decisions=tight), f, indent=1)
        assert all(v < TOL_GRAD for k, v in tight.items() if k != 'relu_decisions_that_differ_from_fp64'), tight
