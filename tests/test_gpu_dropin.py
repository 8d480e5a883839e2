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


def _run(device, patched, tmp, spatial):
    gml = ref_shim.load()
    archs = ref_shim.load_architectures()
    import Modules.model as model
    import Modules.train_rnn as train
    if patched:
        gg.install(gml)
    try:
        N, T, F, K = 20, 4, 6, 3
        S = gg.graphs.sbm(N, 4, 0.7, 0.2, seed=3)[0].numpy()
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
