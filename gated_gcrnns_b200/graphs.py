"""Synthetic graph generators for the benchmark configurations.

Host-side setup only (numpy); the reference builds its graphs with Utils/graphTools.py (SBM, kStepPredGRNNs.py:110-116)
or loads Adj.p (epicenterEstimation.py:474).  These are our own generators of the same families.
"""
from __future__ import annotations

import numpy as np
import torch


def dense_random(N=1024, density=0.3, seed=0) -> torch.Tensor:
    """cfg3: W = triu(rand < density, 1); W += W^T; S = W / lambda_max.  Returns [1,N,N] float32."""
    rng = np.random.RandomState(seed)
    W = np.triu(rng.rand(N, N) < density, 1).astype(np.float64)
    W = W + W.T
    lam = np.abs(np.linalg.eigvalsh(W)).max()
    return torch.tensor(W / lam, dtype=torch.float32).reshape(1, N, N)


def sbm(N=80, C=5, p_in=0.8, p_out=0.2, seed=0) -> torch.Tensor:
    """cfg1-like stochastic block model, S = W / lambda_max, [1,N,N] float32."""
    rng = np.random.RandomState(seed)
    lab = np.arange(N) % C
    P = np.where(lab[:, None] == lab[None, :], p_in, p_out)
    W = np.triu(rng.rand(N, N) < P, 1).astype(np.float64)
    W = W + W.T
    lam = np.abs(np.linalg.eigvalsh(W)).max()
    return torch.tensor(W / lam, dtype=torch.float32).reshape(1, N, N)


def _hilbert_index(pts, bits=10):
    """Index of each point of the unit square along a Hilbert curve of 2^bits x 2^bits cells (vectorised xy -> d)."""
    n = 1 << bits
    x = np.minimum((pts[:, 0] * n).astype(np.int64), n - 1)
    y = np.minimum((pts[:, 1] * n).astype(np.int64), n - 1)
    d = np.zeros_like(x)
    s = n >> 1
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        x &= s - 1
        y &= s - 1
        s >>= 1
    return d


def knn_csr(N=100_000, k=16, seed=0, sigma2=None, power_iters=50):
    """cfg5: directed kNN graph on uniform points of the unit square, weights exp(-d^2/sigma^2),
    normalised by |lambda|max estimated with power iterations.  Returns (rowptr int64, colidx int32, vals float32)."""
    from scipy.spatial import cKDTree
    import scipy.sparse as sp
    rng = np.random.RandomState(seed)
    pts = rng.rand(N, 2)
    pts = pts[np.argsort(_hilbert_index(pts, 10), kind='stable')]   # Hilbert-curve node ordering -> neighbour rows share cache lines
    d, idx = cKDTree(pts).query(pts, k=k + 1)
    d, idx = d[:, 1:], idx[:, 1:]
    if sigma2 is None:
        sigma2 = float(np.mean(d[:, -1] ** 2))
    w = np.exp(-d ** 2 / sigma2)
    rows = np.repeat(np.arange(N), k)
    A = sp.csr_matrix((w.ravel(), (rows, idx.ravel())), shape=(N, N))
    v = np.ones(N) / np.sqrt(N)
    lam = 1.0
    for _ in range(power_iters):
        v2 = A.T @ (A @ v)
        lam = np.sqrt(np.linalg.norm(v2) / max(np.linalg.norm(v), 1e-30))
        v = v2 / max(np.linalg.norm(v2), 1e-30)
    A = (A / lam).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float32)


def csr_to_torch_sparse(rowptr, colidx, vals, N) -> torch.Tensor:
    return torch.sparse_csr_tensor(torch.as_tensor(rowptr), torch.as_tensor(colidx).to(torch.int64), torch.as_tensor(vals),
                                   size=(N, N))


# ---------------------------------------------------------------------------------------------------------------------
# on-GPU builders: the same two inputs built by the CUDA library (csrc/builders.cu)
# ---------------------------------------------------------------------------------------------------------------------
def knn_csr_gpu(N=100_000, k=16, seed=0, sigma2=None, power_iters=50, device='cuda:0', reorder=True, points=None):
    """cfg5's graph built on the GPU: uniform-grid kNN search, weights exp(-d^2/sigma^2), normalisation by the spectral norm from
    `power_iters` power iterations, nodes renumbered along a Hilbert curve of the grid cells when ``reorder`` (the locality the
    gather kernels live on is then supplied by the library, not by the caller).  Same points as ``knn_csr`` for the same seed.
    Returns (rowptr int64, colidx int32, vals float32, info) with info = dict(perm (new -> old), sigma2, lam)."""
    import ctypes as C
    from . import _lib
    if points is None:
        points = np.random.RandomState(seed).rand(N, 2)
    dev = torch.device(device)
    xy = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)).to(dev)
    rowptr = np.empty(N + 1, dtype=np.int64); colidx = np.empty(N * k, dtype=np.int32); vals = np.empty(N * k, dtype=np.float32)
    perm = np.empty(N, dtype=np.int32)
    s2, lam = C.c_float(), C.c_float()
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    _lib.check(_lib.lib().gcrnn_build_knn_csr(N, k, C.c_void_p(xy.data_ptr()), float(sigma2) if sigma2 else 0.0, int(power_iters), int(bool(reorder)),
                                              di, rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p),
                                              vals.ctypes.data_as(C.c_void_p), perm.ctypes.data_as(C.c_void_p), C.byref(s2), C.byref(lam)),
               'build_knn_csr')
    return rowptr, colidx, vals, dict(perm=perm, sigma2=s2.value, lam=lam.value)


def diffusion_signals(graph_handle, x0: torch.Tensor, T: int, noise: torch.Tensor = None) -> torch.Tensor:
    """x_{t+1} = x_t S + w_t on the GPU (Utils/dataTools.py:1290-1297).  x0: [R,N] CUDA fp32, noise: [T,R,N] or None -> [T+1,R,N]."""
    import ctypes as C
    from . import _lib
    assert x0.is_cuda and x0.dtype == torch.float32 and x0.dim() == 2
    R, N = x0.shape
    out = torch.empty(T + 1, R, N, dtype=torch.float32, device=x0.device)
    x0c = x0.contiguous()
    nz = noise.contiguous() if noise is not None else None
    st = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
    _lib.check(_lib.lib().gcrnn_data_diffusion(graph_handle.ptr, C.c_void_p(x0c.data_ptr()), C.c_void_p(nz.data_ptr() if nz is not None else 0),
                                               C.c_void_p(out.data_ptr()), R, T, st), 'data_diffusion')
    return out


def rcm_order(rowptr, colidx, N):
    """Reverse Cuthill-McKee permutation (new -> old) of a user's CSR graph and the renumbered CSR: the sparse gather kernels
    read neighbour rows through L1, so a bandwidth-reducing node order roughly doubles their throughput on graphs given in a
    random order (cfg5: 66 -> 130 seq/s).  Apply the permutation to the node axis of X / h0 and invert it on H."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    A = sp.csr_matrix((np.ones(len(colidx), dtype=np.float32), np.asarray(colidx), np.asarray(rowptr)), shape=(N, N))
    perm = reverse_cuthill_mckee((A + A.T).tocsr(), symmetric_mode=True).astype(np.int64)
    return perm


def permute_csr(rowptr, colidx, vals, perm):
    """CSR of P S P^T for perm (new -> old)."""
    import scipy.sparse as sp
    N = len(perm)
    A = sp.csr_matrix((np.asarray(vals), np.asarray(colidx), np.asarray(rowptr)), shape=(N, N))
    B = A[perm][:, perm].tocsr()
    B.sort_indices()
    return B.indptr.astype(np.int64), B.indices.astype(np.int32), B.data.astype(np.float32)
