"""GPU: on-device builders (csrc/builders.cu) against their host counterparts."""
import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg
from gated_gcrnns_b200 import graph as ggraph

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('N,k', [(500, 4), (5000, 16), (20000, 8)])
def test_knn_graph_matches_kdtree(N, k):
    """Same points: the grid search must return exactly cKDTree's neighbour sets and the same weights; the spectral normalisation
    (same number of power iterations) must agree to 1e-4."""
    from scipy.spatial import cKDTree
    import scipy.sparse as sp
    pts = np.random.RandomState(3).rand(N, 2)
    rp, ci, va, info = gg.graphs.knn_csr_gpu(N, k, points=pts, power_iters=30, reorder=False)
    pts32 = pts.astype(np.float32).astype(np.float64)               # the builder sees fp32 coordinates
    d, idx = cKDTree(pts32).query(pts32, k=k + 1)
    d, idx = d[:, 1:], idx[:, 1:]
    sigma2 = float(np.mean(d[:, -1] ** 2))
    assert abs(info['sigma2'] - sigma2) < 1e-5 * sigma2
    w = np.exp(-d ** 2 / sigma2)
    A = sp.csr_matrix((w.ravel(), (np.repeat(np.arange(N), k), idx.ravel())), shape=(N, N))
    v = np.ones(N) / np.sqrt(N); lam = 1.0
    for _ in range(30):
        v2 = A.T @ (A @ v)
        lam = np.sqrt(np.linalg.norm(v2) / max(np.linalg.norm(v), 1e-30))
        v = v2 / max(np.linalg.norm(v2), 1e-30)
    A = (A / lam).tocsr(); A.sort_indices()
    assert abs(info['lam'] - lam) < 1e-4 * lam
    assert np.array_equal(rp, A.indptr.astype(np.int64))
    same = ci == A.indices.astype(np.int32)
    assert same.mean() > 0.9999, same.mean()                         # fp32 vs fp64 distances may swap an exact near-tie
    assert np.abs(va[same] - A.data.astype(np.float32)[same]).max() < 2e-4 * A.data.max()


def test_knn_reordering_is_a_relabelling_with_locality():
    """reorder=True renumbers the nodes along a Hilbert curve: the graph is the same up to the returned permutation, and neighbour
    indices become close (what the sparse gather kernels' L1 hit rates rest on)."""
    N, k = 20000, 16
    pts = np.random.RandomState(5).rand(N, 2)
    rp0, ci0, va0, _ = gg.graphs.knn_csr_gpu(N, k, points=pts, power_iters=10, reorder=False)
    rp1, ci1, va1, info = gg.graphs.knn_csr_gpu(N, k, points=pts, power_iters=10, reorder=True)
    perm = info['perm']                                              # new -> old
    assert sorted(perm.tolist()) == list(range(N))
    old_rows = {}
    for new in (0, 1, 17, N // 2, N - 1):
        old = perm[new]
        a = sorted(zip(perm[ci1[rp1[new]:rp1[new + 1]]].tolist(), np.round(va1[rp1[new]:rp1[new + 1]], 6).tolist()))
        b = sorted(zip(ci0[rp0[old]:rp0[old + 1]].tolist(), np.round(va0[rp0[old]:rp0[old + 1]], 6).tolist()))
        assert a == b
    med0 = np.median(np.abs(ci0 - np.repeat(np.arange(N), k)))
    med1 = np.median(np.abs(ci1 - np.repeat(np.arange(N), k)))
    assert med1 < 64 and med1 * 20 < med0, (med0, med1)


def test_diffusion_signals_match_dense_recurrence():
    N, R, T = 200, 7, 6
    S = gg.graphs.sbm(N, 4, 0.3, 0.05, seed=2)
    g = ggraph.from_dense(S, DEV)
    gen = torch.Generator().manual_seed(1)
    x0 = torch.rand(R, N, generator=gen).to(DEV)
    noise = 0.1 * torch.randn(T, R, N, generator=gen).to(DEV)
    out = gg.graphs.diffusion_signals(g, x0, T, noise)
    ref = [x0.double()]
    Sd = S[0].double().to(DEV)
    for t in range(T):
        ref.append(ref[-1] @ Sd + noise[t].double())
    ref = torch.stack(ref)
    assert out.shape == (T + 1, R, N)
    assert (out.double() - ref).abs().max() < 1e-5 * ref.abs().max()
