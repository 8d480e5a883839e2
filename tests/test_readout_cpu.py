"""CPU: the batched per-node readout (gated_gcrnns_b200/readout.py) is the SAME function as the reference's Python loop over the
N nodes (Modules/architectures.py:1613-1636).  Runs the reference's own modules on the CPU (skipped without a copy of the
reference tree); the GPU side of the same claim is tests/test_gpu_dropin.py."""
import pytest
import torch

import gated_gcrnns_b200 as gg
from gated_gcrnns_b200 import readout
from oracle import ref_shim


@pytest.mark.parametrize('mlp', ['multipMlp', 'oneMlp', 'gnn'])
def test_regression_forward_is_bit_identical_to_the_reference_loop(mlp):
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree')
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    N, F, K, T, B = 20, 6, 3, 4, 5
    S = gg.graphs.sbm(N, 4, 0.7, 0.2, seed=3)[0].numpy()
    x = torch.randn(B, T, 1, N, generator=torch.Generator().manual_seed(4))
    torch.manual_seed(0)
    if mlp == 'gnn':
        net = archs.GatedGCRNNforRegression(1, F, K, K, torch.tanh, torch.nn.Tanh, [], S, True, time_gating=True, spatial_gating=None,
                                            dimNodeSignals=[F, 4, 1], nFilterTaps=[3, 3], nSelectedNodes=[N, N],
                                            poolingFunction=gml.NoPool, poolingSize=[1, 1])
    else:
        net = archs.GatedGCRNNforRegression(1, F, K, K, torch.tanh, torch.tanh, [1] if mlp == 'multipMlp' else [N], S, True,
                                            time_gating=True, spatial_gating=None, mlpType=mlp)
    h0 = torch.zeros(B, F, N)
    y_ref = net(x, h0)
    y_new = readout.regression_forward(net, x, h0)
    assert y_ref.shape == y_new.shape == (B, T, 1, N)
    assert torch.equal(y_ref, y_new)


def test_install_patches_and_restores_the_architectures():
    if not ref_shim.available():
        pytest.skip('no copy of the reference tree')
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    orig_fwd, orig_init = archs.GatedGCRNNforRegression.forward, archs.GatedGCRNNforClassification.__init__
    gg.install(gml, archs)
    try:
        assert archs.GatedGCRNNforRegression.forward is readout.regression_forward
        S = gg.graphs.sbm(12, 3, 0.7, 0.2, seed=1)[0].numpy()
        net = archs.GatedGCRNNforClassification(1, 4, 2, 2, torch.tanh, torch.nn.ReLU, [3], S, True, time_gating=False, spatial_gating='node')
        assert isinstance(net.stateGCRNN, gg.GGCRNNCell) and net.stateGCRNN.last_state_only
    finally:
        gg.uninstall(gml, archs)
    assert archs.GatedGCRNNforRegression.forward is orig_fwd and archs.GatedGCRNNforClassification.__init__ is orig_init
    assert gml.GGCRNNCell is not gg.GGCRNNCell
