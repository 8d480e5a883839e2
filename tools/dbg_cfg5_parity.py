import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gated_gcrnns_b200 as gg
from gated_gcrnns_b200 import _lib
from oracle import gcrnn_oracle as orc
DEV = 'cuda:0'
def relerr(a, b):
    a = a.detach().cpu().double().numpy(); b = b.detach().cpu().double().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
N, B, T, F_, K = int(sys.argv[1]), 2, int(sys.argv[2]), 32, 3
rp, ci, va = gg.graphs.knn_csr(N, 16, seed=3, power_iters=10)
S = gg.graphs.csr_to_torch_sparse(rp, ci, va, N)
torch.manual_seed(0)
torch.set_default_dtype(torch.float64)
p = orc.init_cell_params(1, F_, K, K, N, False, 'edge', 1, True)
torch.set_default_dtype(torch.float32)
X, h0, dH = torch.randn(B, T, 1, N).double(), 0.1 * torch.randn(B, F_, N).double(), torch.randn(B, T, F_, N).double()
Href, gref = orc.cell_forward_backward(p, [S.double().to_sparse_coo().coalesce()], X, h0, dH, False, 'edge', input_grads=True)
L = _lib.lib()
res = {}
for fused in (1, 0):
    gg.options.set('sparse_fused', fused)
    cell = gg.GGCRNNCell(1, F_, K, K, torch.tanh, False, 'edge', 1, True)
    cell.addGSO(S); cell.load_state_dict(p); cell = cell.to(device=DEV, dtype=torch.float32)
    hg = h0.float().to(DEV).requires_grad_(True)
    H = cell(X.float().to(DEV), hg)
    (H * dH.float().to(DEV)).sum().backward()
    errs = {'H': relerr(H, Href), 'dh0': relerr(hg.grad, gref['__h0'])}
    for k, v in cell.named_parameters():
        if gref[k] is not None: errs[k] = relerr(v.grad, gref[k])
    res[fused] = (hg.grad.detach().clone(), {k: v.grad.clone() for k, v in cell.named_parameters() if v.grad is not None})
    print('fused' if fused else 'generic', {k: '%.1e' % v for k, v in errs.items()})
d = (res[1][0] - res[0][0]).abs()
print('fused-vs-generic dh0 maxdiff', float(d.max()), 'at', np.unravel_index(int(d.argmax()), d.shape), 'max|dh0|', float(res[0][0].abs().max()))
ref = gref['__h0']
dd = (res[1][0].cpu().double() - ref).abs()
idx = np.unravel_index(int(dd.argmax()), dd.shape)
print('fused-vs-oracle worst at', idx, float(dd.max()), 'in-degree of that node', int(np.bincount(ci, minlength=N)[idx[2]]), 'count > 1e-4*max:', int((dd > 1e-4 * ref.abs().max()).sum()))
gg.options.set('sparse_fused', 1)
for name, fused in (('fused', 1), ('generic', 0)):
    dd = (res[fused][0].cpu().double() - ref).abs()
    bad = (dd > 1e-4 * ref.abs().max()).nonzero()
    nodes = np.unique(bad[:, 2].numpy())
    gaps = np.diff(nodes)
    clusters = 1 + int((gaps > 3000).sum()) if len(nodes) else 0
    print(name, 'bad elements', len(bad), 'bad nodes', len(nodes), 'clusters (gap > 3000 in Hilbert index)', clusters,
          'samples', np.unique(bad[:, 0].numpy()).tolist(), 'node range', (int(nodes.min()), int(nodes.max())) if len(nodes) else None)
