"""CPU: the host-side mirror of the reference interface — names, init order, state_dict, error behaviour."""
import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg
from tests import _golden as G


@pytest.mark.parametrize('name', G.names('init_'))
def test_init_matches_reference_bitwise(name):
    c = G.load(name)
    tg = name.split('_')[1] == 't1'
    sg = {'none': None, 'node': 'node', 'edge': 'edge'}[name.split('_')[2]]
    S = G.t64(G.load('cell_small_t0_none')['S'])
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(0)
        cell = gg.GGCRNNCell(2, 3, 3, 2, torch.tanh, tg, sg, 1, True)
        cell.addGSO(S)
    finally:
        torch.set_default_dtype(prev)
    sd = cell.state_dict()
    assert list(sd.keys()) == [str(k) for k in c['keys']]          # same names, same order, S not in it
    for k, v in c['param'].items():
        assert sd[k].shape == v.shape
        assert np.array_equal(sd[k].numpy(), v), k                  # same RNG draw order -> identical values


def test_state_dict_roundtrip_and_unused_params():
    torch.manual_seed(1)
    a = gg.GGCRNNCell(1, 4, 3, 3, torch.tanh, True, 'node')
    a.addGSO(torch.rand(1, 6, 6))
    b = gg.GGCRNNCell(1, 4, 3, 3, torch.tanh, True, 'node')
    b.addGSO(torch.rand(1, 6, 6))
    b.load_state_dict(a.state_dict())
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    used = {n for _, _, n in gg.cell_param_slots(True, 'node', True)}
    allp = {n for n, _ in a.named_parameters()}
    assert allp - used == {'GFL_out.weight_A', 'GFL_out.weight_B', 'GFL_out.bias', 'MLP_out.0.weight', 'MLP_out.0.bias'}


def test_signatures_match_reference_defaults():
    import inspect
    sig = inspect.signature(gg.GGCRNNCell.__init__)
    assert list(sig.parameters)[1:] == ['G', 'F', 'Kin', 'Kst', 'sigma', 'time_gating', 'spatial_gating', 'E', 'bias']
    assert sig.parameters['time_gating'].default is True and sig.parameters['spatial_gating'].default is None
    assert list(inspect.signature(gg.GraphFilter.__init__).parameters)[1:] == ['G', 'F', 'K', 'E', 'bias']
    assert list(inspect.signature(gg.GraphAttentional.__init__).parameters)[1:] == ['G', 'F', 'K', 'E', 'nonlinearity',
                                                                                    'concatenate']
    assert list(inspect.signature(gg.LSIGF).parameters) == ['h', 'S', 'x', 'b']


def test_no_cpu_fallback():
    cell = gg.GGCRNNCell(1, 2, 2, 2, torch.tanh, False, None)
    cell.addGSO(torch.rand(1, 5, 5))
    with pytest.raises(gg.GcrnnError):
        cell(torch.randn(2, 3, 1, 5), torch.zeros(2, 2, 5))
    f = gg.GraphFilter(1, 2, 2)
    f.addGSO(torch.rand(1, 5, 5))
    with pytest.raises(gg.GcrnnError):
        f(torch.randn(2, 1, 5))


def test_shape_asserts_like_reference():
    cell = gg.GGCRNNCell(1, 2, 2, 2, torch.tanh, False, None)
    with pytest.raises(AssertionError):
        cell.addGSO(torch.rand(2, 5, 5))        # E mismatch (graphML.py:2241)
    with pytest.raises(AssertionError):
        cell.addGSO(torch.rand(5, 5))           # not 3-D (graphML.py:2239)
    cell.addGSO(torch.rand(1, 5, 5))
    with pytest.raises(AssertionError):
        cell(torch.randn(2, 3, 1, 5), torch.zeros(3, 2, 5))   # batch mismatch (graphML.py:2339)


def test_install_rebinds_only_hot_path_names():
    import types
    fake = types.ModuleType('fake_graphML')
    for n in ('LSIGF', 'GraphFilter', 'GraphAttentional', 'GGCRNNCell', 'NoPool'):
        setattr(fake, n, object())
    keep = fake.NoPool
    gg.install(fake)
    assert fake.GGCRNNCell is gg.GGCRNNCell and fake.LSIGF is gg.LSIGF and fake.NoPool is keep
    gg.uninstall(fake)
    assert fake.GGCRNNCell is not gg.GGCRNNCell


def test_rcm_order_is_a_bandwidth_reducing_relabelling():
    """Host helper for user graphs given in a random node order (the library renumbers only the graphs it builds itself)."""
    import numpy as np
    import scipy.sparse as sp
    from gated_gcrnns_b200 import graphs
    rng = np.random.RandomState(0)
    N = 400
    pts = rng.rand(N, 2)
    d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    A = sp.csr_matrix((d < 0.01) & (d > 0)).astype(np.float32)
    rp, ci, va = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
    perm = graphs.rcm_order(rp, ci, N)
    assert sorted(perm.tolist()) == list(range(N))
    rp2, ci2, va2 = graphs.permute_csr(rp, ci, va, perm)
    bw = lambda r, c: np.abs(c - np.repeat(np.arange(N), np.diff(r))).mean()
    assert bw(rp2, ci2) < 0.5 * bw(rp, ci)
    B = sp.csr_matrix((va2, ci2, rp2), shape=(N, N))
    assert abs(B.sum() - A.sum()) < 1e-3 and B.nnz == A.nnz
