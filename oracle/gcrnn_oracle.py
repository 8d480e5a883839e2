"""CPU oracle for the gated GCRNN recurrence (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product path
(``gated_gcrnns_b200``) never does; it fails loudly when its CUDA library is
missing.

What it is: an independent fp64 restatement, in plain torch-CPU tensor algebra
(einsum + explicit per-edge loops, autograd only to differentiate the restated
forward), of the algorithm in the reference's ``Utils/graphML.py``:

  * ``lsigf``            <- ``LSIGF``            graphML.py:47-140
  * ``graph_attention``  <- ``graphAttention``   graphML.py:521-627
                            + ``GraphAttentional.forward`` graphML.py:2084-2116
  * ``graph_filter``     <- ``GraphFilter.forward`` graphML.py:1175-1194
  * ``cell_forward``     <- ``GGCRNNCell.forward`` graphML.py:2336-2428
  * ``init_cell_params`` <- ``GGCRNNCell.__init__/reset_parameters/addGSO``
                            graphML.py:2196-2334 (parameter names, shapes and
                            RNG draw order)

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against the reference ITSELF: ``oracle/make_golden.py``
imports ``/root/reference`` (possible in the build container only), runs the
unmodified ``GGCRNNCell`` / ``GraphFilter`` / ``GraphAttentional`` in fp64 and
commits inputs, parameters, outputs and gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
to <= 1e-12.

All third-party arithmetic in the reference is PyTorch ATen (no pinned
version); the formulas below are those call sites restated.

The GSO may be given dense (``[E,N,N]`` tensor) or, for graphs too large for
the dense reference, as a list of E ``torch.sparse`` matrices; the attention
is evaluated edge-wise over the pattern of ``S + I`` in both cases.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Union

import torch

ZERO_TOL = 1e-9          # graphML.py:42  zeroTolerance
NEG_SLOPE = 0.2          # graphML.py:521 negative_slope default
RELU = torch.relu        # the attention layer's nonlinearity (:2101).  A test may swap it for a relu with GIVEN one-sided
                         # derivatives to show that a deviation comes from the kink and nothing else; never changed otherwise.

GSO = Union[torch.Tensor, Sequence[torch.Tensor]]


# ---------------------------------------------------------------------------
# GSO helpers
# ---------------------------------------------------------------------------
def _gso_list(S: GSO) -> List[torch.Tensor]:
    """Split ``S`` into E per-edge-feature operators (dense 2-D or sparse 2-D)."""
    if isinstance(S, torch.Tensor) and not S.is_sparse and S.layout == torch.strided:
        assert S.dim() == 3 and S.shape[1] == S.shape[2]
        return [S[e] for e in range(S.shape[0])]
    if isinstance(S, torch.Tensor):
        return [S]
    return list(S)


def _shift(z: torch.Tensor, Se: torch.Tensor) -> torch.Tensor:
    """Row-vector shift ``z @ S_e`` on the last axis (graphML.py:123)."""
    if Se.layout == torch.strided:
        return z @ Se
    lead = z.shape[:-1]
    zt = z.reshape(-1, z.shape[-1]).t()                 # [N, R]
    return torch.sparse.mm(Se.t().coalesce(), zt).t().reshape(*lead, -1)


# ---------------------------------------------------------------------------
# LSIGF  (graphML.py:47-140)
# ---------------------------------------------------------------------------
def lsigf(h: torch.Tensor, S: GSO, x: torch.Tensor,
          b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[b,f,n] = sum_{e,k,g} h[f,e,k,g] (x S_e^k)[b,g,n] + b[f]."""
    F, E, K, G = h.shape
    Ss = _gso_list(S)
    assert len(Ss) == E
    assert x.shape[1] == G
    y = torch.zeros(x.shape[0], F, x.shape[2], dtype=x.dtype)
    for e in range(E):
        z = x                                            # k = 0: S^0 = I for every e (graphML.py:117)
        for k in range(K):
            if k > 0:
                z = _shift(z, Ss[e])                     # graphML.py:123
            y = y + torch.einsum('fg,bgn->bfn', h[:, e, k, :], z)   # graphML.py:134-135
    if b is not None:
        y = y + b.reshape(1, F, -1)                      # graphML.py:138-139
    return y


def graph_filter(weight, bias, S: GSO, x, N: Optional[int] = None):
    """GraphFilter.forward incl. node zero-padding (graphML.py:1175-1194)."""
    Ss = _gso_list(S)
    N = Ss[0].shape[0] if N is None else N
    Nin = x.shape[2]
    if Nin < N:
        x = torch.cat([x, torch.zeros(x.shape[0], x.shape[1], N - Nin, dtype=x.dtype)], 2)
    u = lsigf(weight, S, x, bias)
    return u[:, :, :Nin]


# ---------------------------------------------------------------------------
# Graph attention (edge gate)  (graphML.py:521-627, 2084-2116)
# ---------------------------------------------------------------------------
def _edges_of_S_plus_I(Se: torch.Tensor):
    """COO (i, j, value) of S' = S + I restricted to |S'| > ZERO_TOL (graphML.py:577, 611-613)."""
    N = Se.shape[0]
    if Se.layout == torch.strided:
        Sp = Se.detach() + torch.eye(N, dtype=Se.dtype)
        idx = (Sp.abs() > ZERO_TOL).nonzero(as_tuple=False)
        i, j = idx[:, 0], idx[:, 1]
        return i, j, Sp[i, j]
    eye = torch.sparse_coo_tensor(torch.arange(N).repeat(2, 1), torch.ones(N, dtype=Se.dtype), (N, N))
    Sp = (Se.detach().to_sparse_coo() + eye).coalesce()
    i, j = Sp.indices()
    v = Sp.values()
    keep = v.abs() > ZERO_TOL
    return i[keep], j[keep], v[keep]


def graph_attention(x, mixer, weight, S: GSO):
    """One-head, one-edge-feature GAT exactly as the reference evaluates it.

    e_ij  = leaky_relu(a2.Wx_i + a1.Wx_j)                  graphML.py:591-603
    al_ij = softmax over j in {j: |S'_ij| > tol}           graphML.py:611-622
    y[:,j]= sum_i Wx[:,i] S'_ij al_ij                      graphML.py:625
    Returns [B, K*F, N] after ReLU + head concat (K heads)  graphML.py:2099-2107.
    """
    K, E, twoF = mixer.shape
    Fo = twoF // 2
    Ss = _gso_list(S)
    assert len(Ss) == E
    B, G, N = x.shape
    heads = []
    for k in range(K):
        yk = torch.zeros(B, Fo, N, dtype=x.dtype)
        for e in range(E):
            ei, ej, ev = _edges_of_S_plus_I(Ss[e])
            # NOTE: the reference builds ONE mask from sum_e |S'_e| (graphML.py:611);
            # with E = 1 (the only value the cell uses, graphML.py:2327) it is this one.
            assert E == 1, "oracle restates the E=1 attention used by GGCRNNCell"
            Wx = torch.einsum('fg,bgn->bfn', weight[k, e], x)          # :586-588
            r = torch.einsum('f,bfn->bn', mixer[k, e, :Fo], Wx)        # a1 . Wx_j
            c = torch.einsum('f,bfn->bn', mixer[k, e, Fo:], Wx)        # a2 . Wx_i
            s = c[:, ei] + r[:, ej]                                    # [B, nnz]
            s = torch.where(s > 0, s, NEG_SLOPE * s)
            # row-wise (over j) softmax restricted to edges
            m = torch.full((B, N), -float('inf'), dtype=x.dtype)
            m = m.scatter_reduce(1, ei.expand(B, -1), s.detach(), 'amax', include_self=True)
            p = torch.exp(s - m[:, ei])
            den = torch.zeros(B, N, dtype=x.dtype).index_add(1, ei, p)
            al = p / den[:, ei]
            contrib = Wx[:, :, ei] * (ev * al).unsqueeze(1)            # [B,F,nnz]
            yk = yk + torch.zeros(B, Fo, N, dtype=x.dtype).index_add(2, ej, contrib)
        heads.append(RELU(yk))                                         # :2101
    return torch.cat(heads, 1)                                         # (k, f) order :2105-2107


# ---------------------------------------------------------------------------
# Parameter construction with the reference's names and RNG order
# ---------------------------------------------------------------------------
def _uniform(shape, s):
    return torch.empty(*shape).uniform_(-s, s)


def _plain_cell_params(G, F, Kin, Kst, E, bias, prefix, out):
    s = 1.0 / math.sqrt(G * Kin)                        # graphML.py:2231 (also used for weight_B!)
    out[prefix + 'weight_A'] = _uniform((F, E, Kin, G), s)
    out[prefix + 'weight_B'] = _uniform((F, E, Kst, F), s)
    if bias:
        out[prefix + 'bias'] = _uniform((F, 1), s)


def _linear_params(fan_in, bias, prefix, out):
    # torch.nn.Linear.reset_parameters: kaiming_uniform_(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for weight and bias
    bound = 1.0 / math.sqrt(fan_in)
    out[prefix + 'weight'] = _uniform((1, fan_in), bound)
    if bias:
        out[prefix + 'bias'] = _uniform((1,), bound)


def init_cell_params(G, F, Kin, Kst, N, time_gating=True, spatial_gating=None, E=1, bias=True
                     ) -> Dict[str, torch.Tensor]:
    """state_dict of ``GGCRNNCell(...)`` followed by ``addGSO`` drawn from the current torch RNG."""
    p: Dict[str, torch.Tensor] = {}
    _plain_cell_params(G, F, Kin, Kst, E, bias, '', p)                 # graphML.py:2215-2227
    if time_gating:                                                    # graphML.py:2249-2291
        for g in ('in', 'forget', 'out'):
            _plain_cell_params(G, F, Kin, Kst, E, bias, f'GFL_{g}.', p)
            _linear_params(N * F, bias, f'MLP_{g}.0.', p)
    if spatial_gating == 'node':                                       # graphML.py:2294-2323
        for g in ('in', 'forget'):
            _plain_cell_params(G, F, Kin, Kst, E, bias, f'GRNN_node_{g}.', p)
            s = 1.0 / math.sqrt(F * Kst)                               # GraphFilter(F,1,Kst) graphML.py:1160
            p[f'GFL_node_{g}.0.weight'] = _uniform((1, E, Kst, F), s)
            if bias:
                p[f'GFL_node_{g}.0.bias'] = _uniform((1, 1), s)
    elif spatial_gating == 'edge':                                     # graphML.py:2325-2334
        for g in ('input', 'forget'):
            s = 1.0 / math.sqrt(F * 1)                                 # graphML.py:2066 (G=F, K=1)
            # draw order in GraphAttentional.reset_parameters: weight, then mixer;
            # registration order (state_dict order): mixer, then weight
            w = _uniform((1, 1, F, F), s)
            m = _uniform((1, 1, 2 * F), s)
            p[f'{g}_attention.mixer'] = m
            p[f'{g}_attention.weight'] = w
    return p


# ---------------------------------------------------------------------------
# GGCRNNCell.forward  (graphML.py:2336-2428)
# ---------------------------------------------------------------------------
def _sub(p, prefix):
    return p[prefix + 'weight_A'], p[prefix + 'weight_B'], p.get(prefix + 'bias')


def _subcell_state(p, prefix, S, x_t, h0):
    """One step of an ungated sub-cell from the INITIAL state h0 (graphML.py:2362,2417-2423)."""
    A, Bw, b = _sub(p, prefix)
    return torch.tanh(lsigf(A, S, x_t, b) + lsigf(Bw, S, h0, b))


def cell_forward(p: Dict[str, torch.Tensor], S: GSO, X: torch.Tensor, h0: torch.Tensor,
                 time_gating=True, spatial_gating=None) -> torch.Tensor:
    """H[B,T,F,N] of the gated GCRNN.  ``p`` uses the reference's state_dict keys."""
    B, T, G, N = X.shape
    A, Bw, b = _sub(p, '')
    F = A.shape[0]
    h = h0
    out = []
    for t in range(T):
        x = X[:, t]
        gi = gf = torch.ones(B, 1, 1, dtype=X.dtype)
        if time_gating:                                                # :2357-2374
            ui = _subcell_state(p, 'GFL_in.', S, x, h0).reshape(B, F * N)
            gi = torch.sigmoid(ui @ p['MLP_in.0.weight'].t()
                               + (p['MLP_in.0.bias'] if 'MLP_in.0.bias' in p else 0)).reshape(B, 1, 1)
            uf = _subcell_state(p, 'GFL_forget.', S, x, h0).reshape(B, F * N)
            gf = torch.sigmoid(uf @ p['MLP_forget.0.weight'].t()
                               + (p['MLP_forget.0.bias'] if 'MLP_forget.0.bias' in p else 0)).reshape(B, 1, 1)
        a = lsigf(A, S, x, b)                                          # input filter
        r = lsigf(Bw, S, h, b)                                         # state filter (same bias, :2405-2407)
        if spatial_gating == 'node':                                   # :2379-2407
            si = _subcell_state(p, 'GRNN_node_in.', S, x, h0)
            qi = torch.sigmoid(graph_filter(p['GFL_node_in.0.weight'], p.get('GFL_node_in.0.bias'), S, si))
            sf = _subcell_state(p, 'GRNN_node_forget.', S, x, h0)
            qf = torch.sigmoid(graph_filter(p['GFL_node_forget.0.weight'], p.get('GFL_node_forget.0.bias'), S, sf))
            a = qi * a
            r = qf * r
        elif spatial_gating == 'edge':                                 # :2409-2416
            a = graph_attention(a, p['input_attention.mixer'], p['input_attention.weight'], S)
            r = graph_attention(r, p['forget_attention.mixer'], p['forget_attention.weight'], S)
        else:
            assert spatial_gating is None
        h = torch.tanh(gi * a + gf * r)                                # :2417-2423
        out.append(h)
    return torch.stack(out, 1)                                         # :2425-2427


def cell_forward_backward(p, S, X, h0, dH, time_gating=True, spatial_gating=None,
                          input_grads=False):
    """(H, {name: grad}) with grads of sum(H*dH); unused params map to None like the reference."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    Xr, hr = X, h0
    if input_grads:
        Xr = X.detach().clone().requires_grad_(True)
        hr = h0.detach().clone().requires_grad_(True)
    H = cell_forward(q, S, Xr, hr, time_gating, spatial_gating)
    (H * dH).sum().backward()
    g = {k: v.grad for k, v in q.items()}
    if input_grads:
        g['__X'] = Xr.grad
        g['__h0'] = hr.grad
    return H.detach(), g
