"""Generate tests/golden/*.npz by running the UNMODIFIED reference (fp64, CPU).

Run once in the build container:  python oracle/make_golden.py
Test infrastructure; needs /root/reference (see oracle/ref_shim.py).  Every
fixture holds the inputs, the reference's parameters (state_dict), its
output and autograd gradients for loss = sum(H * dH), so that the GPU box
needs neither the reference nor this script.
"""
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def rand_gso(N, density, symmetric, seed, E=1):
    g = torch.Generator().manual_seed(seed)
    W = torch.rand(E, N, N, generator=g) * (torch.rand(E, N, N, generator=g) < density)
    for e in range(E):
        W[e].fill_diagonal_(0.0)
    if symmetric:
        W = torch.triu(W, 1)
        W = W + W.transpose(1, 2)
    lam = max(torch.linalg.eigvals(W[e]).abs().max().item() for e in range(E))
    return (W / lam).double()


def adj_p_gso():
    """epicenterEstimation.py:474-479,962-963: Adj.p / |lambda|max as [1,59,59]."""
    with open(os.path.join(ref_shim.reference_root(), 'Adj.p'), 'rb') as f:
        A = np.asarray(pickle.load(f), dtype=np.float64)
    lam = np.abs(np.linalg.eigvals(A)).max()
    return torch.tensor(A / lam).reshape(1, *A.shape)


def sbm_gso(N=80, C=5, p_in=0.8, p_out=0.2, seed=0):
    """SBM graph in the spirit of kStepPredGRNNs.py:110-116 (own generator), S = W / lambda_max."""
    rng = np.random.RandomState(seed)
    lab = np.arange(N) % C
    P = np.where(lab[:, None] == lab[None, :], p_in, p_out)
    W = np.triu(rng.rand(N, N) < P, 1).astype(np.float64)
    W = W + W.T
    lam = np.abs(np.linalg.eigvalsh(W)).max()
    return torch.tensor(W / lam).reshape(1, N, N)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f'{k}::{kk}'] = np.asarray(vv)
        else:
            flat[k] = np.asarray(v)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **flat)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def cell_case(gml, name, S, G, F, Kin, Kst, T, B, tg, sg, bias=True, seed=0, h0_zero=False):
    torch.manual_seed(seed)
    N = S.shape[1]
    E = S.shape[0]
    cell = gml.GGCRNNCell(G, F, Kin, Kst, torch.tanh, tg, sg, E, bias)
    cell.addGSO(S)
    X = torch.randn(B, T, G, N, requires_grad=True)
    h0 = (torch.zeros(B, F, N) if h0_zero else 0.5 * torch.randn(B, F, N)).requires_grad_(True)
    dH = torch.randn(B, T, F, N)
    H = cell(X, h0)
    (H * dH).sum().backward()
    params = {k: v.detach().numpy() for k, v in cell.state_dict().items()}
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(0))
             for k, v in cell.named_parameters()}
    save(name, S=S.numpy(), X=X.detach().numpy(), h0=h0.detach().numpy(), dH=dH.numpy(),
         H=H.detach().numpy(), dX=X.grad.numpy(), dh0=h0.grad.numpy(),
         meta=np.array([G, F, Kin, Kst, T, B, int(tg), {None: 0, 'node': 1, 'edge': 2}[sg], int(bias), seed, E]),
         param=params, grad=grads)


def main():
    torch.set_default_dtype(torch.float64)
    gml = ref_shim.load()

    # ---- LSIGF / GraphFilter (E = 2, padding) -------------------------------------
    torch.manual_seed(1)
    S = rand_gso(10, 0.4, False, 11, E=2)
    gf = gml.GraphFilter(3, 4, 3, 2, True)
    gf.addGSO(S)
    x = torch.randn(3, 3, 10, requires_grad=True)
    dy = torch.randn(3, 4, 10)
    y = gf(x)
    (y * dy).sum().backward()
    xs = torch.randn(3, 3, 7)                      # Nin < N: zero padding path (graphML.py:1181-1194)
    ys = gf(xs)
    save('lsigf_e2', S=S.numpy(), x=x.detach().numpy(), dy=dy.numpy(), y=y.detach().numpy(),
         dx=x.grad.numpy(), weight=gf.weight.detach().numpy(), bias=gf.bias.detach().numpy(),
         dweight=gf.weight.grad.numpy(), dbias=gf.bias.grad.numpy(), x_short=xs.numpy(),
         y_short=ys.detach().numpy())

    # ---- GraphAttentional ---------------------------------------------------------
    torch.manual_seed(2)
    S = rand_gso(12, 0.3, False, 12)
    ga = gml.GraphAttentional(5, 5, 1)
    ga.addGSO(S)
    x = torch.randn(3, 5, 12, requires_grad=True)
    dy = torch.randn(3, 5, 12)
    y = ga(x)
    (y * dy).sum().backward()
    save('gat', S=S.numpy(), x=x.detach().numpy(), dy=dy.numpy(), y=y.detach().numpy(), dx=x.grad.numpy(),
         mixer=ga.mixer.detach().numpy(), weight=ga.weight.detach().numpy(),
         dmixer=ga.mixer.grad.numpy(), dweight=ga.weight.grad.numpy())

    # ---- cell: all six gating modes, non-symmetric S, G=2, Kin != Kst, h0 != 0 ------
    S12 = rand_gso(12, 0.35, False, 13)
    for tg in (False, True):
        for sg in (None, 'node', 'edge'):
            cell_case(gml, f'cell_small_t{int(tg)}_{sg or "none"}', S12, 2, 4, 3, 4, 5, 3, tg, sg, seed=3)
    cell_case(gml, 'cell_small_nobias_t1_node', S12, 2, 4, 3, 4, 4, 2, True, 'node', bias=False, seed=4)
    cell_case(gml, 'cell_small_nobias_t0_edge', S12, 1, 4, 2, 3, 4, 2, False, 'edge', bias=False, seed=5)
    S2 = rand_gso(9, 0.5, False, 14, E=2)
    cell_case(gml, 'cell_small_e2_t1_node', S2, 2, 3, 3, 2, 3, 2, True, 'node', seed=6)

    # ---- cfg1-like: SBM N=80, F=20, K=5, T=5, time-gated ---------------------------
    cell_case(gml, 'cell_cfg1_time', sbm_gso(), 1, 20, 5, 5, 5, 4, True, None, seed=0, h0_zero=True)
    # ---- cfg2: Adj.p, F=20, K=4, T=20, node- and edge-gated -------------------------
    Sq = adj_p_gso()
    cell_case(gml, 'cell_cfg2_node', Sq, 1, 20, 4, 4, 20, 3, False, 'node', seed=0, h0_zero=True)
    cell_case(gml, 'cell_cfg2_edge', Sq, 1, 20, 4, 4, 20, 3, False, 'edge', seed=0, h0_zero=True)

    # ---- init parity: state_dict for seed 0 ---------------------------------------
    for tg, sg in ((True, None), (True, 'node'), (False, 'edge')):
        torch.manual_seed(0)
        cell = gml.GGCRNNCell(2, 3, 3, 2, torch.tanh, tg, sg, 1, True)
        cell.addGSO(S12)
        save(f'init_t{int(tg)}_{sg or "none"}',
             keys=np.array(list(cell.state_dict().keys())),
             param={k: v.numpy() for k, v in cell.state_dict().items()})


if __name__ == '__main__':
    main()
