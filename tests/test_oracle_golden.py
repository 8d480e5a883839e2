"""CPU: the oracle restatement reproduces the reference's own outputs/gradients (golden fixtures)."""
import numpy as np
import pytest
import torch

from oracle import gcrnn_oracle as orc
from tests import _golden as G

TOL = 1e-12


def _close(a, b, tol=TOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_lsigf_golden():
    c = G.load('lsigf_e2')
    S, x, w, b = G.t64(c['S']), G.t64(c['x']).requires_grad_(True), G.t64(c['weight']).requires_grad_(True), \
        G.t64(c['bias']).requires_grad_(True)
    y = orc.lsigf(w, S, x, b)
    (y * G.t64(c['dy'])).sum().backward()
    _close(y.detach(), c['y']); _close(x.grad, c['dx']); _close(w.grad, c['dweight']); _close(b.grad, c['dbias'])
    _close(orc.graph_filter(w.detach(), b.detach(), S, G.t64(c['x_short'])), c['y_short'])


def test_gat_golden():
    c = G.load('gat')
    S, x = G.t64(c['S']), G.t64(c['x']).requires_grad_(True)
    m, w = G.t64(c['mixer']).requires_grad_(True), G.t64(c['weight']).requires_grad_(True)
    y = orc.graph_attention(x, m, w, S)
    (y * G.t64(c['dy'])).sum().backward()
    _close(y.detach(), c['y']); _close(x.grad, c['dx']); _close(m.grad, c['dmixer']); _close(w.grad, c['dweight'])


@pytest.mark.parametrize('name', G.names('cell_'))
def test_cell_golden(name):
    c = G.load(name)
    m = G.cell_meta(c)
    p = {k: G.t64(v) for k, v in c['param'].items()}
    H, g = orc.cell_forward_backward(p, G.t64(c['S']), G.t64(c['X']), G.t64(c['h0']), G.t64(c['dH']),
                                     m['time_gating'], m['spatial_gating'], input_grads=True)
    _close(H, c['H'])
    _close(g['__X'], c['dX']); _close(g['__h0'], c['dh0'])
    for k, ref in c['grad'].items():
        if ref.size == 0:                      # parameter the reference never uses (GFL_out.*, MLP_out.*)
            assert g[k] is None
        else:
            _close(g[k], ref, 1e-11)


@pytest.mark.parametrize('name', G.names('init_'))
def test_init_golden(name):
    c = G.load(name)
    tg = name.split('_')[1] == 't1'
    sg = {'none': None, 'node': 'node', 'edge': 'edge'}[name.split('_')[2]]
    torch.manual_seed(0)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        p = orc.init_cell_params(2, 3, 3, 2, 12, tg, sg, 1, True)
    finally:
        torch.set_default_dtype(prev)
    assert list(p.keys()) == [str(k) for k in c['keys']]
    for k, v in c['param'].items():
        _close(p[k], v, 1e-15)   # nn.Linear draws via kaiming_uniform_: bound differs by <= 1 ulp


def test_sparse_gso_matches_dense():
    c = G.load('cell_small_t0_edge')
    m = G.cell_meta(c)
    p = {k: G.t64(v) for k, v in c['param'].items()}
    S = G.t64(c['S'])
    H = orc.cell_forward(p, [S[0].to_sparse_coo()], G.t64(c['X']), G.t64(c['h0']), m['time_gating'], m['spatial_gating'])
    _close(H.detach(), c['H'])
