"""nn.Modules with the reference's names, constructor signatures, parameter names and init order.

Host-side mirror of ``GraphFilter`` (Utils/graphML.py:1086-1205), ``GraphAttentional`` (:1999-2128) and
``GGCRNNCell`` (:2130-2428).  ``Modules/architectures.py`` builds these through ``gml.<Name>(...)`` +
``addGSO(S)`` + ``forward`` (architectures.py:1521-1524, :1611), so after ``install()`` the reference's
architectures and training loops run unchanged on top of the CUDA library.

Differences that are deliberate:
  * ``forward`` needs CUDA tensors (there is no CPU path here; the reference's CPU path is the baseline);
    a CPU ``h0`` next to a CUDA ``X`` is moved, because ``train_rnn.py:256`` creates ``h0`` on the CPU;
  * ``GGCRNNCell.forward`` runs ONE fused recurrence call instead of the reference's Python time loop;
    the gate sub-modules below are parameter containers with the reference's attribute names (they stay
    usable on their own);
  * ``sigma`` must be tanh (the only nonlinearity the reference scripts ever pass: kStepPredGRNNs.py:268,
    epicenterEstimation.py:217).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as Fn
from . import graph as _graph
from . import _lib


def _is_tanh(sigma) -> bool:
    if sigma in (torch.tanh, nn.functional.tanh, nn.Tanh, torch.Tensor.tanh):
        return True
    return isinstance(sigma, nn.Tanh)


def _cuda_only(x, what):
    if not x.is_cuda:
        raise _lib.GcrnnError(f'{what}: input is on {x.device}; gated_gcrnns_b200 has no CPU path '
                              '(the reference implementation is the CPU path)')


class GraphFilter(nn.Module):
    """GraphFilter(G, F, K, E=1, bias=True): y = LSIGF(weight, S, x, bias)   (graphML.py:1086)."""

    def __init__(self, G, F, K, E=1, bias=True):
        super().__init__()
        self.G, self.F, self.K, self.E = G, F, K, E
        self.S = None
        self.weight = nn.parameter.Parameter(torch.Tensor(F, E, K, G))
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.K)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        E, N = _graph.gso_shape(S)
        assert E == self.E
        self.N = N
        self.S = S

    def forward(self, x):
        _cuda_only(x, 'GraphFilter')
        B, _, Nin = x.shape
        if Nin < self.N:                                     # zero padding, graphML.py:1181-1185
            x = torch.cat((x, x.new_zeros(B, x.shape[1], self.N - Nin)), dim=2)
        u = Fn.LSIGF(self.weight, self.S, x, self.bias)
        if Nin < self.N:
            u = u[:, :, :Nin]
        return u

    def extra_repr(self):
        s = 'in_features=%d, out_features=%d, filter_taps=%d, edge_features=%d, bias=%s, ' % (
            self.G, self.F, self.K, self.E, self.bias is not None)
        return s + ('GSO stored' if self.S is not None else 'no GSO stored')


class GraphAttentional(nn.Module):
    """GraphAttentional(G, F, K, E=1, nonlinearity=relu, concatenate=True)   (graphML.py:1999)."""

    def __init__(self, G, F, K, E=1, nonlinearity=nn.functional.relu, concatenate=True):
        super().__init__()
        self.G, self.F, self.K, self.E = G, F, K, E
        self.S = None
        self.nonlinearity = nonlinearity
        self.concatenate = concatenate
        self.mixer = nn.parameter.Parameter(torch.Tensor(K, E, 2 * F))
        self.weight = nn.parameter.Parameter(torch.Tensor(K, E, F, G))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.K)
        self.weight.data.uniform_(-stdv, stdv)
        self.mixer.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        E, N = _graph.gso_shape(S)
        assert E == self.E
        self.N = N
        self.S = S

    def _preactivation(self, x, k):
        """Attention output of head k BEFORE the nonlinearity.  The CUDA kernel fuses the ReLU the cell uses (graphML.py:2327);
        negating both the mixer and the weight leaves every attention coefficient unchanged and flips the sign of the output,
        so  y = relu(GAT(x; a, W)) - relu(GAT(x; -a, -W))  recovers it from two calls of the same kernel."""
        a, w = self.mixer[k, 0], self.weight[k, 0]
        return Fn.graph_attention_relu(x, a, w, self.S) - Fn.graph_attention_relu(x, -a, -w, self.S)

    def forward(self, x):
        _cuda_only(x, 'GraphAttentional')
        if self.E != 1:
            raise NotImplementedError('the CUDA attention kernel implements one edge feature (E = 1), which is all GGCRNNCell uses '
                                      '(graphML.py:2327); there is no fallback path')
        B, _, Nin = x.shape
        if Nin < self.N:
            x = torch.cat((x, x.new_zeros(B, x.shape[1], self.N - Nin)), dim=2)
        relu = self.nonlinearity in (nn.functional.relu, torch.relu)
        if self.concatenate:
            # nonlinearity per head, then heads stacked (k, f)-major (graphML.py:2099-2107)
            heads = [Fn.graph_attention_relu(x, self.mixer[k, 0], self.weight[k, 0], self.S) if relu
                     else self.nonlinearity(self._preactivation(x, k)) for k in range(self.K)]
            y = heads[0] if self.K == 1 else torch.cat(heads, dim=1)
        elif self.K == 1 and relu:
            y = Fn.graph_attention_relu(x, self.mixer[0, 0], self.weight[0, 0], self.S)
        else:
            # average over the heads first, then the nonlinearity (graphML.py:2108-2112)
            y = self.nonlinearity(torch.stack([self._preactivation(x, k) for k in range(self.K)], dim=0).mean(dim=0))
        if Nin < self.N:
            y = y[:, :, :Nin]
        return y

    def extra_repr(self):
        s = 'in_features=%d, out_features=%d, attention_heads=%d, edge_features=%d, ' % (self.G, self.F, self.K, self.E)
        return s + ('GSO stored: number_nodes=%d' % self.N if self.S is not None else 'no GSO stored')


class GGCRNNCell(nn.Module):
    """GGCRNNCell(G, F, Kin, Kst, sigma=nn.Tanh, time_gating=True, spatial_gating=None, E=1, bias=True)

    forward(X[B,T,G,N], h0[B,F,N]) -> H[B,T,F,N]   (graphML.py:2130-2428)."""

    def __init__(self, G, F, Kin, Kst, sigma=nn.Tanh, time_gating=True, spatial_gating=None, E=1, bias=True):
        super().__init__()
        self.G, self.F, self.Kin, self.Kst, self.E = G, F, Kin, Kst, E
        self.S = None
        self.weight_A = nn.parameter.Parameter(torch.Tensor(F, E, Kin, G))
        self.weight_B = nn.parameter.Parameter(torch.Tensor(F, E, Kst, F))
        self.sigma = sigma
        self.time_gating = time_gating
        self.spatial_gating = spatial_gating
        self.bias_flag = bias
        if bias:
            self.bias = nn.parameter.Parameter(torch.Tensor(F, 1))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()
        self._handles = {}
        # True: forward computes every state but exposes only the last one, as a stride-0 expansion [B,T,F,N] of H[:, T-1]
        # (so `H.select(1, -1)` of the reference's classifier, architectures.py:1844, works unchanged); no [B,T,F,N] output or
        # output-gradient tensor exists.  Set by install() on the classification architecture's cell, or by hand.
        self.last_state_only = False

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.G * self.Kin)       # also for weight_B, as in the reference (graphML.py:2231)
        self.weight_A.data.uniform_(-stdv, stdv)
        self.weight_B.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def addGSO(self, S):
        E, N = _graph.gso_shape(S)
        assert E == self.E
        self.N = N
        self.S = S
        self._handles = {}
        mk = lambda: GGCRNNCell(self.G, self.F, self.Kin, self.Kst, self.sigma, time_gating=False, E=self.E,
                                bias=self.bias_flag)
        # creation order = RNG draw order = state_dict order of the reference (graphML.py:2249-2334)
        if self.time_gating == True:  # noqa: E712  (the reference compares with ==)
            for g in ('in', 'forget', 'out'):
                sub = mk()
                sub.addGSO(self.S)
                setattr(self, f'GFL_{g}', sub)
                setattr(self, f'MLP_{g}', nn.Sequential(nn.Linear(self.N * self.F, 1, bias=self.bias_flag), nn.Sigmoid()))
        if self.spatial_gating is not None:
            if self.spatial_gating == 'node':
                for g in ('in', 'forget'):
                    sub = mk()
                    sub.addGSO(self.S)
                    setattr(self, f'GRNN_node_{g}', sub)
                    head = GraphFilter(self.F, 1, self.Kst, self.E, self.bias_flag)
                    head.addGSO(self.S)
                    setattr(self, f'GFL_node_{g}', nn.Sequential(head, nn.Sigmoid()))
            elif self.spatial_gating == 'edge':
                for g in ('input', 'forget'):
                    att = GraphAttentional(self.F, self.F, 1)
                    att.addGSO(self.S)
                    setattr(self, f'{g}_attention', att)

    # -- the fused path ------------------------------------------------------------------------------------
    def _tc_unsupported(self, split: bool, need_dx: bool):
        """Why the dense tensor-core kernels cannot take this cell (None if they can).  Mirrors EVERY check of the
        library's tensor-core path (csrc/gcrnn_tc.cu: tc_dims, launch_tap, cell_backward_tc) so that 'auto' never
        picks a path that raises later."""
        if self.S.layout != torch.strided:
            return 'the GSO is sparse (dense [E,N,N] tensor needed)'
        if self.E != 1:
            return f'E={self.E} (E == 1 needed)'
        if self.spatial_gating == 'edge':
            return "spatial_gating='edge' (time and node gating only)"
        if self.spatial_gating == 'node' and self.Kin * self.G > 8:
            return f'node gating with Kin*G={self.Kin * self.G} (<= 8)'
        if self.N % (256 if split else 128) != 0:
            return f'N={self.N} (N % {256 if split else 128} == 0 needed)'
        if self.F not in (16, 32, 64):
            return f'F={self.F} (F in 16, 32, 64)'
        if self.Kin * self.G > 32:
            return f'Kin*G={self.Kin * self.G} (<= 32)'
        if not 1 <= self.Kst <= 5 or self.Kst * self.F > 384:
            return f'Kst={self.Kst} (<= 5)'
        if need_dx and self.Kin * self.G > 8:
            return f'X requires grad with Kin*G={self.Kin * self.G} (input gradients on the tensor-core path need <= 8)'
        return None

    def _precision_for(self, device, need_dx=False):
        mode = Fn.get_precision()
        if mode in ('bf16', 'bf16x2'):
            why = self._tc_unsupported(mode == 'bf16x2', need_dx)
            if why is not None:
                raise _lib.GcrnnError(f"precision {mode!r}: the tensor-core path does not take this cell: {why}")
            return _lib.PREC_BF16_TC if mode == 'bf16' else _lib.PREC_BF16X2_TC
        if mode == 'auto' and self.N >= 256 and self._tc_unsupported(True, need_dx) is None:
            return _lib.PREC_BF16X2_TC          # never plain bf16: 'auto' must not silently loosen the numerics
        return _lib.PREC_FP32

    def _handle(self, device, need_dx=False):
        prec = self._precision_for(device, need_dx)
        key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), prec)
        h = self._handles.get(key)
        if h is None:
            g = _graph.get(self.S, device, keep_dense=(prec != _lib.PREC_FP32))
            sg = self.spatial_gating if self.spatial_gating in ('node', 'edge') else None
            h = Fn.CellHandle(g, self.G, self.F, self.Kin, self.Kst, self.E, self.time_gating == True, sg,  # noqa: E712
                              self.bias_flag, prec)
            self._handles[key] = h
        return h

    # handles wrap raw C pointers owned by a finalizer: never copy or pickle them (copy.deepcopy(model) / torch.save(model)
    # are how the reference snapshots its best model); they are rebuilt lazily on the next forward
    def __getstate__(self):
        st = self.__dict__.copy()
        st['_handles'] = {}
        return st

    def _used_parameters(self, slots):
        sd = dict(self.named_parameters())
        return [sd[name] for _, _, name in slots]

    def forward(self, X, h0):
        assert h0.shape[0] == X.shape[0]
        assert self.S is not None, 'call addGSO(S) first'
        if not _is_tanh(self.sigma):
            raise NotImplementedError('GGCRNNCell on B200 implements sigma = tanh only (no fallback path)')
        if self.spatial_gating not in (None, 'node', 'edge'):
            # the reference silently skips the state update for unknown strings (graphML.py:2379-2416)
            raise NotImplementedError(f'spatial_gating={self.spatial_gating!r}')
        _cuda_only(X, 'GGCRNNCell')
        assert X.shape[3] == self.N and X.shape[2] == self.G
        assert h0.shape[1] == self.F and h0.shape[2] == self.N
        if h0.device != X.device:                      # train_rnn.py:256 builds h0 on the CPU
            h0 = h0.to(X.device)
        cell = self._handle(X.device, need_dx=bool(X.requires_grad and torch.is_grad_enabled()))
        if getattr(self, 'last_state_only', False):
            hl = Fn.gated_gcrnn(cell, X, h0, self._used_parameters(cell.slots), last_state_only=True)
            return hl.unsqueeze(1).expand(X.shape[0], X.shape[1], self.F, self.N)
        return Fn.gated_gcrnn(cell, X, h0, self._used_parameters(cell.slots))

    def extra_repr(self):
        return 'in_features=%d, state_features=%d, input_taps=%d, state_taps=%d, edge_features=%d, time_gating=%s, ' \
               'spatial_gating=%s, bias=%s' % (self.G, self.F, self.Kin, self.Kst, self.E, self.time_gating,
                                               self.spatial_gating, self.bias_flag)
