"""GPU: whole-step CUDA-graph capture (gated_gcrnns_b200/train.py).

One captured graph = node-reordering gather + recurrence + readout + loss + backward + optimiser update.  Its loss trajectory and
final parameters must equal the eager loop's (same kernels, same order: only float-atomic summation order may differ)."""
import numpy as np
import pytest
import torch

import gated_gcrnns_b200 as gg

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class _Net(torch.nn.Module):
    """k-step prediction model of the reference's shape: gated GCRNN + the same per-node readout applied at every node."""

    def __init__(self, S, F, K, tg, sg):
        super().__init__()
        self.cell = gg.GGCRNNCell(1, F, K, K, torch.tanh, tg, sg, 1, True)
        self.cell.addGSO(S)
        self.readout = torch.nn.Linear(F, 1)

    def forward(self, x, h0):
        H = self.cell(x, h0)                                         # [B,T,F,N]
        return self.readout(H.transpose(2, 3)).squeeze(-1).unsqueeze(2)   # [B,T,1,N]


@pytest.mark.parametrize('tg,sg', [(True, None), (False, 'node'), (False, 'edge')])
def test_graphed_step_matches_eager_loop(tg, sg):
    gg.set_precision('fp32')
    N, F, K, T, B = 40, 8, 3, 4, 12
    S = gg.graphs.sbm(N, 4, 0.7, 0.2, seed=3)
    g = torch.Generator().manual_seed(9)
    xs = [torch.randn(B, T, 1, N, generator=g).to(DEV) for _ in range(6)]
    ys = [torch.randn(B, T, 1, N, generator=g).to(DEV) for _ in range(6)]
    h0 = torch.zeros(B, F, N, device=DEV)
    order = list(np.random.RandomState(0).permutation(N))
    l1 = torch.nn.L1Loss()

    def build():
        torch.manual_seed(0)
        net = _Net(S, F, K, tg, sg).to(DEV)
        return net, torch.optim.Adam(net.parameters(), lr=1e-2, capturable=True)

    net_e, opt_e = build()
    eager = []
    oidx = torch.as_tensor(order, device=DEV)
    for x, y in zip(xs, ys):
        opt_e.zero_grad(set_to_none=True)
        loss = l1(net_e(x.index_select(-1, oidx), h0), y)
        loss.backward()
        opt_e.step()
        eager.append(float(loss))

    net_g, opt_g = build()
    sd0 = {k: v.clone() for k, v in net_g.state_dict().items()}
    step = gg.train.GraphedStep(net_g, l1, opt_g, xs[0], h0, target=ys[0], order=order, warmup=2)
    # the warm-up and capture passes trained on the example batch: restore the initial state (parameters and Adam moments)
    net_g.load_state_dict(sd0)
    for st in opt_g.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()
    L = gg._lib.lib()
    l0 = L.gcrnn_debug_launch_count()
    graphed = [float(step(x, h0, target=y)) for x, y in zip(xs, ys)]
    assert L.gcrnn_debug_launch_count() == l0            # replays launch nothing through the library's host path
    np.testing.assert_allclose(graphed, eager, rtol=2e-4, atol=1e-6)
    for (k, a), (_, b) in zip(net_g.state_dict().items(), net_e.state_dict().items()):
        assert (a - b).abs().max() <= 2e-4 * max(1.0, b.abs().max()), k
