"""GPU (optional): the reference's own architectures.py + train_rnn.py run UNCHANGED on top of our modules.

Needs a copy of the reference tree (oracle/ref_shim.py looks in $GCRNN_REFERENCE_ROOT, /root/reference and the
git-ignored baseline/_ref); skipped when there is none.  The reference side runs on the CPU (its only mode), ours on
cuda:0 after `gg.install()`; both start from the same seed, so parameters are identical, and the recorded training
losses must agree to fp32 accuracy over the first optimisation steps.
"""
import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg
from oracle import ref_shim

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class _Data:
    """Minimal stand-in for Utils.dataTools.KStepPrediction: what train_rnn.MultipleModels touches."""

    def __init__(self, N, T, nTrain, nValid, device):
        g = torch.Generator().manual_seed(5)
        self.nTrain, self.nValid = nTrain, nValid
        self.x = {'train': torch.randn(nTrain, T, N, generator=g).to(device), 'valid': torch.randn(nValid, T, N, generator=g).to(device)}
        self.y = {k: 0.5 * torch.roll(v, 1, dims=2) for k, v in self.x.items()}

    def getSamples(self, which, idx=None):
        if idx is None:
            return self.x[which], self.y[which]
        return self.x[which][idx], self.y[which][idx]

    def evaluate(self, yHat, y):
        return torch.mean((yHat - y) ** 2)


def _run(device, patched, tmp, spatial, N=20, F=6, dense=False):
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    import Modules.model as model
    import Modules.train_rnn as train
    if patched:
        gg.install(gml)
    try:
        T, K = 4, 3
        S = (gg.graphs.dense_random(N, 0.3, seed=3) if dense else gg.graphs.sbm(N, 4, 0.7, 0.2, seed=3))[0].numpy()
        torch.manual_seed(0)
        np.random.seed(0)
        net = archs.GatedGCRNNforRegression(1, F, K, K, torch.tanh, torch.tanh, [N], S, True, time_gating=True,
                                            spatial_gating=spatial, mlpType='oneMlp')
        net.to(device)
        losses = []
        l1 = torch.nn.L1Loss()

        def loss(yhat, y):
            v = l1(yhat, y)
            losses.append(float(v))
            return v
        opt = torch.optim.SGD(net.parameters(), lr=0.05)
        m = model.Model(net, loss, opt, 'GCRNNdropin', str(tmp), list(range(N)))
        data = _Data(N, T, 12, 4, device)
        train.MultipleModels({'GCRNNdropin': m}, data, 2, 4, T, F, F, validationInterval=100)
        _run.precisions = sorted({k[1] for mod in net.modules() if hasattr(mod, '_handles') for k in mod._handles})
        return losses, {k: v.detach().cpu().double() for k, v in net.state_dict().items()}
    finally:
        if patched:
            gg.uninstall(gml)


@pytest.mark.parametrize('spatial', [None, 'node', 'edge'])
def test_reference_training_loop_runs_unchanged(tmp_path, spatial):
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree on this box')
    ref_losses, ref_sd = _run('cpu', False, tmp_path / 'ref', spatial)
    our_losses, our_sd = _run(DEV, True, tmp_path / 'ours', spatial)
    assert len(ref_losses) == len(our_losses) and len(ref_losses) >= 6
    assert list(ref_sd.keys()) == list(our_sd.keys())
    np.testing.assert_allclose(our_losses, ref_losses, rtol=2e-4, atol=1e-6)
    for k in ref_sd:        # parameters after the SGD steps (and after MultipleModels re-loaded the 'Best' checkpoint)
        assert (our_sd[k] - ref_sd[k]).abs().max() <= 2e-4 * max(1.0, ref_sd[k].abs().max()), k


@pytest.mark.parametrize('spatial', [None, 'node'])
def test_reference_training_loop_on_the_tensor_core_path(tmp_path, spatial):
    """The same unchanged reference loop on a DENSE N = 256 graph with precision 'auto': the cell (time gates, and time + node gates)
    then runs on the split-bf16 tcgen05 path.  Losses of the first optimisation steps against the reference's own fp32 CPU run, at
    the tensor-core mode's bound instead of the fp32 one."""
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree on this box')
    ref_losses, ref_sd = _run('cpu', False, tmp_path / 'ref', spatial, N=256, F=16, dense=True)
    l0 = gg._lib.lib().gcrnn_debug_launch_count()
    try:
        gg.set_precision('auto')
        our_losses, our_sd = _run(DEV, True, tmp_path / 'ours', spatial, N=256, F=16, dense=True)
    finally:
        gg.set_precision('fp32')
    assert gg._lib.lib().gcrnn_debug_launch_count() > l0
    assert _run.precisions == [gg._lib.PREC_BF16X2_TC], f'the cell did not take the split-bf16 tensor-core path: {_run.precisions}'
    assert len(ref_losses) == len(our_losses) and len(ref_losses) >= 6
    np.testing.assert_allclose(our_losses, ref_losses, rtol=2e-3, atol=1e-5)
    for k in ref_sd:
        assert (our_sd[k] - ref_sd[k]).abs().max() <= 5e-3 * max(1.0, ref_sd[k].abs().max()), k


# ---------------------------------------------------------------------------------------------------------------
# SURVEY 8f rank 1 + 2: readouts.  The reference architectures run UNCHANGED on our modules (after
# install(gml, architectures)) and must reproduce the reference's own CPU run: outputs, every parameter gradient and a few
# optimisation steps of the reference's training loops.
# ---------------------------------------------------------------------------------------------------------------
def _adj_p():
    """The seismograph graph of epicenterEstimation.py (Adj.p / |lambda|max), from the committed golden fixture."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cell_cfg2_node.npz'), allow_pickle=False)
    return np.asarray(z['S'], dtype=np.float32).reshape(59, 59)      # fp32: both sides run in the default dtype


def _fwd_bwd(net, x, h0, target, loss_fn):
    net.zero_grad()
    y = net(x, h0)
    loss = loss_fn(y, target)
    loss.backward()
    return y.detach().double().cpu(), float(loss), {k: (None if p.grad is None else p.grad.detach().double().cpu()) for k, p in net.named_parameters()}


def _compare(ref, ours, tol_y=1e-5, tol_g=2e-4):
    (yr, lr, gr), (yo, lo, go) = ref, ours
    assert yr.shape == yo.shape
    assert (yr - yo).abs().max() <= tol_y * max(1.0, yr.abs().max())
    assert abs(lr - lo) <= tol_y * max(1.0, abs(lr))
    assert list(gr.keys()) == list(go.keys())
    for k in gr:
        if gr[k] is None:
            assert go[k] is None, k
            continue
        assert (gr[k] - go[k]).abs().max() <= tol_g * max(gr[k].abs().max(), 1e-30), (k, float((gr[k] - go[k]).abs().max()), float(gr[k].abs().max()))


@pytest.mark.parametrize('spatial', ['node', 'edge'])
def test_classifier_last_state_only_matches_reference(spatial):
    """GatedGCRNNforClassification on Adj.p (cfg2: N=59, F=20, K=4, T=20, 11 classes): install() switches its cell to
    last-state-only mode (no [B,T,F,N] output / output-gradient tensor); logits, loss and all gradients equal the reference's."""
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree on this box')
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    S = _adj_p()
    N, F, K, T, B, C_ = 59, 20, 4, 20, 16, 11
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, 1, N, generator=g)
    target = torch.randint(0, C_, (B,), generator=g)
    ce = torch.nn.CrossEntropyLoss()

    def build():
        torch.manual_seed(0)
        return archs.GatedGCRNNforClassification(1, F, K, K, torch.tanh, torch.nn.ReLU, [C_], S, True, time_gating=False,
                                                 spatial_gating=spatial)
    ref = _fwd_bwd(build(), x, torch.zeros(B, F, N), target, ce)
    gg.install(gml, archs)
    try:
        net = build()
        assert net.stateGCRNN.last_state_only
        net.to(DEV)
        H = net.stateGCRNN(x.to(DEV), torch.zeros(B, F, N, device=DEV))
        assert H.shape == (B, T, F, N) and H.stride(1) == 0           # a stride-0 expansion of the last state: nothing materialised
        ours = _fwd_bwd(net, x.to(DEV), torch.zeros(B, F, N), target.to(DEV), ce)
        # the same model WITHOUT the last-state shortcut must agree too
        net.stateGCRNN.last_state_only = False
        full = _fwd_bwd(net, x.to(DEV), torch.zeros(B, F, N), target.to(DEV), ce)
    finally:
        gg.uninstall(gml, archs)
    _compare(ref, ours)
    _compare(ref, full)


@pytest.mark.parametrize('mlp', ['multipMlp', 'oneMlp'])
def test_regression_batched_node_readout_matches_reference(mlp):
    """GatedGCRNNforRegression: the reference applies the per-node MLP in a Python loop over N (architectures.py:1613-1636);
    readout.regression_forward is one batched contraction.  Reference on the CPU vs ours on the GPU."""
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree on this box')
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    N, F, K, T, B = 20, 6, 3, 4, 5
    S = gg.graphs.sbm(N, 4, 0.7, 0.2, seed=3)[0].numpy()
    g = torch.Generator().manual_seed(4)
    x, y = torch.randn(B, T, 1, N, generator=g), torch.randn(B, T, 1, N, generator=g)
    l1 = torch.nn.L1Loss()

    def build():
        torch.manual_seed(0)
        dims = [1] if mlp == 'multipMlp' else [N]
        return archs.GatedGCRNNforRegression(1, F, K, K, torch.tanh, torch.tanh, dims, S, True, time_gating=True,
                                             spatial_gating=None, mlpType=mlp)
    ref = _fwd_bwd(build(), x, torch.zeros(B, F, N), y, l1)
    gg.install(gml, archs)
    try:
        net = build()
        net.to(DEV)
        ours = _fwd_bwd(net, x.to(DEV), torch.zeros(B, F, N), y.to(DEV), l1)
    finally:
        gg.uninstall(gml, archs)
    _compare(ref, ours)


def test_gcrnn_gnn_and_selection_gnn_on_our_graph_filter():
    """the GCRNN with a Selection-GNN readout (kStepPredGRNNs.py:308-335: `GCRNN_GNN`) and the stand-alone
    SelectionGNN baseline (architectures.py:10-177) run on OUR GraphFilter after install(); outputs and gradients match the
    reference's CPU run."""
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree on this box')
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    N, F, K, T, B = 20, 6, 3, 4, 5
    S = gg.graphs.sbm(N, 4, 0.7, 0.2, seed=3)[0].numpy()
    g = torch.Generator().manual_seed(4)
    x, y = torch.randn(B, T, 1, N, generator=g), torch.randn(B, T, 1, N, generator=g)
    l1 = torch.nn.L1Loss()

    def build_gcrnn_gnn():
        torch.manual_seed(0)
        # dimNodeSignals [F_h, 4, 1], taps [3, 3], no pooling (NoPool, all nodes kept), no final MLP
        return archs.GatedGCRNNforRegression(1, F, K, K, torch.tanh, torch.nn.Tanh, [], S, True, time_gating=True, spatial_gating=None,
                                             dimNodeSignals=[F, 4, 1], nFilterTaps=[3, 3], nSelectedNodes=[N, N],
                                             poolingFunction=gml.NoPool, poolingSize=[1, 1])

    def build_sel():
        torch.manual_seed(0)
        return archs.SelectionGNN([1, 5, 3], [3, 2], True, torch.nn.Tanh, [N, N], gml.NoPool, [1, 1], [2], S)

    ref_a = _fwd_bwd(build_gcrnn_gnn(), x, torch.zeros(B, F, N), y, l1)
    sel_ref = build_sel()
    xs = torch.randn(7, 1, N, generator=g)
    ys_ref = sel_ref(xs); ys_ref.square().sum().backward()
    gref = {k: p.grad.double() for k, p in sel_ref.named_parameters()}
    gg.install(gml, archs)
    try:
        net = build_gcrnn_gnn()
        assert type(net.outputNN[0].GFL[0]).__module__.startswith('gated_gcrnns_b200')      # the readout filters are ours
        net.to(DEV)
        ours_a = _fwd_bwd(net, x.to(DEV), torch.zeros(B, F, N), y.to(DEV), l1)
        sel = build_sel()
        sel.to(DEV)
        ys = sel(xs.to(DEV)); ys.square().sum().backward()
        gour = {k: p.grad.double().cpu() for k, p in sel.named_parameters()}
    finally:
        gg.uninstall(gml, archs)
    _compare(ref_a, ours_a)
    assert (ys.detach().cpu() - ys_ref.detach()).abs().max() <= 1e-5 * max(1.0, ys_ref.abs().max())
    for k in gref:
        assert (gref[k] - gour[k]).abs().max() <= 2e-4 * gref[k].abs().max(), k


@pytest.mark.parametrize('K,concat,nl', [(3, True, 'tanh'), (3, False, 'relu'), (2, False, 'sigmoid'), (1, True, 'relu')])
def test_graph_attentional_general_heads_and_nonlinearity(K, concat, nl):
    """GraphAttentional beyond what the cell uses (graphML.py:1999-2128): several heads, averaging instead of concatenation, any
    nonlinearity — against the reference module on the CPU."""
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree on this box')
    gml = ref_shim.load()
    fn = {'tanh': torch.tanh, 'relu': torch.nn.functional.relu, 'sigmoid': torch.sigmoid}[nl]
    N, G_, F_, B = 30, 4, 5, 6
    S = torch.rand(1, N, N) * (torch.rand(1, N, N) < 0.2)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, G_, N, generator=g)
    dy = torch.randn(B, K * F_ if concat else F_, N, generator=g)
    torch.manual_seed(0)
    ref = gml.GraphAttentional(G_, F_, K, 1, fn, concat)
    ref.addGSO(S)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    (yr * dy).sum().backward()
    torch.manual_seed(0)
    ours = gg.GraphAttentional(G_, F_, K, 1, fn, concat)
    ours.addGSO(S)
    ours = ours.to(DEV)
    xo = x.clone().to(DEV).requires_grad_(True)
    yo = ours(xo)
    (yo * dy.to(DEV)).sum().backward()
    assert yo.shape == yr.shape
    assert (yo.detach().cpu() - yr.detach()).abs().max() <= 2e-5 * max(1.0, yr.abs().max())
    assert (xo.grad.cpu() - xr.grad).abs().max() <= 2e-4 * max(xr.grad.abs().max(), 1e-30)
    for (k, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
        assert (po.grad.cpu() - pr.grad).abs().max() <= 2e-4 * max(pr.grad.abs().max(), 1e-30), k
