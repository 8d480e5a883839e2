"""autograd.Functions over the C ABI: LSIGF, graph attention and the fused gated-GCRNN recurrence.

Host-side mirror of the reference functionals in Utils/graphML.py (LSIGF :47-140,
graphAttention :521-627, GGCRNNCell.forward :2336-2428).  All arithmetic happens in
libgcrnn_b200.so; torch is used for device memory, streams and autograd plumbing only.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import graph as _graph
from . import dist as _dist
from . import options as _options

_precision = 'fp32'     # 'fp32' | 'bf16x2' | 'bf16' | 'auto'
PRECISIONS = ('fp32', 'bf16x2', 'bf16', 'auto')


def set_precision(mode: str):
    """'fp32'  : exact CUDA-core path everywhere (sparse shift).
    'bf16x2': dense tcgen05 path with every operand split into bf16 hi + lo planes (16-bit mantissa, fp32 accumulate):
              the tensor-core mode with a tight bound; raises for cells it does not support.
    'bf16'  : dense tcgen05 path with plain bf16 operands (8-bit mantissa): fastest, short horizons / contractive
              recurrences only; raises for cells it does not support.
    'auto'  : 'bf16x2' for every cell whose shape the tensor-core kernels take, 'fp32' otherwise.  Never plain bf16."""
    global _precision
    assert mode in PRECISIONS
    _precision = mode


def get_precision() -> str:
    return _precision


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _bytes(n, device):
    return torch.empty(max(int(n), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------
# LSIGF
# ---------------------------------------------------------------------------
class _LsigfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, x, b, g: _graph.Graph):
        dev = x.device
        F, E, K, G = h.shape
        B = x.shape[0]
        h32, x32 = _f32(h, dev), _f32(x, dev)
        b32 = _f32(b, dev).reshape(-1) if b is not None else None
        y = torch.empty(B, F, g.N, dtype=torch.float32, device=dev)
        L = _lib.lib()
        wsb = L.gcrnn_lsigf_workspace_bytes(g.ptr, F, K, G, B)
        ws = _bytes(wsb, dev)
        _lib.check(L.gcrnn_lsigf_forward(g.ptr, _ptr(h32), _ptr(b32), _ptr(x32), _ptr(y), F, K, G, B, _ptr(ws),
                                         ws.numel(), _stream(dev)), 'lsigf_forward')
        ctx.g, ctx.dims = g, (F, E, K, G, B)
        ctx.types = (h.dtype, x.dtype, b.dtype if b is not None else None, b.shape if b is not None else None)
        ctx.save_for_backward(h32, x32)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        h32, x32 = ctx.saved_tensors
        g = ctx.g
        F, E, K, G, B = ctx.dims
        hd, xd, bd, bshape = ctx.types
        dev = x32.device
        dy32 = _f32(dy, dev)
        need_h, need_x, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and bd is not None
        dx = torch.empty_like(x32) if need_x else None
        dh = torch.zeros_like(h32) if need_h else None
        db = torch.zeros(F, dtype=torch.float32, device=dev) if need_b else None
        L = _lib.lib()
        ws = _bytes(L.gcrnn_lsigf_workspace_bytes(g.ptr, F, K, G, B), dev)
        _lib.check(L.gcrnn_lsigf_backward(g.ptr, _ptr(h32), _ptr(x32), _ptr(dy32), _ptr(dx), _ptr(dh), _ptr(db), F, K, G,
                                          B, _ptr(ws), ws.numel(), _stream(dev)), 'lsigf_backward')
        return (dh.to(hd) if need_h else None, dx.to(xd) if need_x else None,
                db.reshape(bshape).to(bd) if need_b else None, None)


def _as_graph(S, device) -> _graph.Graph:
    return S if isinstance(S, _graph.Graph) else _graph.get(S, device)


def LSIGF(h, S, x, b=None):
    """Drop-in for ``Utils.graphML.LSIGF(h, S, x, b=None)`` (graphML.py:47): same argument order and shapes.

    h:[F,E,K,G], S:[E,N,N] (dense or torch sparse; any device) , x:[B,G,N] on a CUDA device, b:[F,1] -> [B,F,N].
    """
    F, E, K, G = h.shape
    Es, N = (S.E, S.N) if isinstance(S, _graph.Graph) else _graph.gso_shape(S)
    assert Es == E
    assert x.shape[1] == G
    assert x.shape[2] == N
    if b is not None:
        assert b.numel() == F, 'only one scalar bias per output feature is supported (graphML.py:45)'
    return _LsigfFn.apply(h, x, b, _as_graph(S, x.device))


# ---------------------------------------------------------------------------
# graph attention (one head) + ReLU
# ---------------------------------------------------------------------------
class _GatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mixer, weight, g: _graph.Graph):
        dev = x.device
        F, G = weight.shape
        B = x.shape[0]
        x32, m32, w32 = _f32(x, dev), _f32(mixer, dev), _f32(weight, dev)
        y = torch.empty(B, F, g.N, dtype=torch.float32, device=dev)
        L = _lib.lib()
        ws = _bytes(L.gcrnn_gat_workspace_bytes(g.ptr, F, G, B), dev)
        _lib.check(L.gcrnn_gat_forward(g.ptr, _ptr(m32), _ptr(w32), _ptr(x32), _ptr(y), F, G, B, _ptr(ws), ws.numel(),
                                       _stream(dev)), 'gat_forward')
        ctx.g, ctx.dims, ctx.types = g, (F, G, B), (x.dtype, mixer.dtype, weight.dtype)
        ctx.save_for_backward(x32, m32, w32)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        x32, m32, w32 = ctx.saved_tensors
        g = ctx.g
        F, G, B = ctx.dims
        dev = x32.device
        dy32 = _f32(dy, dev)
        dx = torch.empty_like(x32) if ctx.needs_input_grad[0] else None
        dm, dw = torch.zeros_like(m32), torch.zeros_like(w32)
        L = _lib.lib()
        ws = _bytes(L.gcrnn_gat_workspace_bytes(g.ptr, F, G, B), dev)
        _lib.check(L.gcrnn_gat_backward(g.ptr, _ptr(m32), _ptr(w32), _ptr(x32), _ptr(dy32), _ptr(dx), _ptr(dm), _ptr(dw),
                                        F, G, B, _ptr(ws), ws.numel(), _stream(dev)), 'gat_backward')
        xd, md, wd = ctx.types
        return (dx.to(xd) if dx is not None else None, dm.to(md), dw.to(wd), None)


def graph_attention_relu(x, mixer, weight, S):
    """relu(graphAttention(x, a, W, S)) for one head and one edge feature (graphML.py:521-627, :2101).

    x:[B,G,N], mixer:[2F], weight:[F,G] -> [B,F,N]."""
    return _GatFn.apply(x, mixer.reshape(-1), weight, _as_graph(S, x.device))


# ---------------------------------------------------------------------------
# the fused gated GCRNN recurrence
# ---------------------------------------------------------------------------
def cell_param_slots(time_gating: bool, spatial_gating: Optional[str], bias: bool) -> List[Tuple[str, Optional[int], str]]:
    """(C struct field, gate index, reference state_dict key) for every parameter the recurrence USES.

    ``GFL_out.*`` / ``MLP_out.0.*`` exist in the reference's state_dict (graphML.py:2281-2291) but never enter
    ``forward``; they are not listed, so their gradient stays ``None`` exactly as in the reference."""
    s: List[Tuple[str, Optional[int], str]] = [('weight_A', None, 'weight_A'), ('weight_B', None, 'weight_B')]
    if bias:
        s.append(('bias', None, 'bias'))
    if time_gating:
        for i, g in enumerate(('in', 'forget')):
            s += [('t_weight_A', i, f'GFL_{g}.weight_A'), ('t_weight_B', i, f'GFL_{g}.weight_B')]
            if bias:
                s.append(('t_bias', i, f'GFL_{g}.bias'))
            s.append(('t_mlp_w', i, f'MLP_{g}.0.weight'))
            if bias:
                s.append(('t_mlp_b', i, f'MLP_{g}.0.bias'))
    if spatial_gating == 'node':
        for i, g in enumerate(('in', 'forget')):
            s += [('n_weight_A', i, f'GRNN_node_{g}.weight_A'), ('n_weight_B', i, f'GRNN_node_{g}.weight_B')]
            if bias:
                s.append(('n_bias', i, f'GRNN_node_{g}.bias'))
            s.append(('n_head_w', i, f'GFL_node_{g}.0.weight'))
            if bias:
                s.append(('n_head_b', i, f'GFL_node_{g}.0.bias'))
    elif spatial_gating == 'edge':
        for i, g in enumerate(('input', 'forget')):
            s += [('e_mixer', i, f'{g}_attention.mixer'), ('e_weight', i, f'{g}_attention.weight')]
    return s


def _fill_struct(slots, tensors) -> _lib.CellParams:
    st = _lib.CellParams()
    for (field, idx, _), t in zip(slots, tensors):
        if idx is None:
            setattr(st, field, t.data_ptr())
        else:
            getattr(st, field)[idx] = t.data_ptr()
    return st


class CellHandle:
    """One ``gcrnn_cell*`` (a cell configuration bound to a graph handle)."""

    def __init__(self, g: _graph.Graph, G, F, Kin, Kst, E, time_gating, spatial_gating, bias, precision):
        import weakref
        self.graph = g
        self.desc = _lib.CellDesc(G, F, Kin, Kst, E, int(bool(time_gating)), _lib.SPATIAL[spatial_gating], int(bool(bias)),
                                  precision)
        out = C.c_void_p()
        _lib.check(_lib.lib().gcrnn_cell_create(C.byref(out), C.byref(self.desc), g.ptr), 'cell_create')
        self.handle = out.value
        self.slots = cell_param_slots(time_gating, spatial_gating, bias)
        self.F, self.G, self.N = F, G, g.N
        self._opt_version = -1
        self._fin = weakref.finalize(self, _destroy_cell, self.handle)
        self.sync_options()

    def sync_options(self):
        """Apply the host-side switch table (gated_gcrnns_b200.options) to this handle if it changed."""
        v = _options.version()
        if v != self._opt_version:
            for name, value in _options.items():
                self.set_option(name, value)
            self._opt_version = v

    @property
    def ptr(self):
        return C.c_void_p(self.handle)

    def set_option(self, name: str, value: int):
        _lib.check(_lib.lib().gcrnn_cell_set_option(self.ptr, name.encode(), int(value)), 'cell_set_option')

    def get_option(self, name: str) -> int:
        v = C.c_int32()
        _lib.check(_lib.lib().gcrnn_cell_get_option(self.ptr, name.encode(), C.byref(v)), 'cell_get_option')
        return v.value

    def workspace(self, B, T, input_grads):
        s, f, b = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.check(_lib.lib().gcrnn_cell_workspace_bytes(self.ptr, B, T, int(input_grads), C.byref(s), C.byref(f),
                                                         C.byref(b)), 'cell_workspace_bytes')
        return s.value, f.value, b.value


def _destroy_cell(h):
    try:
        _lib.lib().gcrnn_cell_destroy(C.c_void_p(h))
    except Exception:
        pass


class _CellFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cell: CellHandle, last_only: bool, X, h0, *params):
        dev = X.device
        B, T = int(X.shape[0]), int(X.shape[1])
        X32, h32 = _f32(X, dev), _f32(h0, dev)
        p32 = [_f32(p, dev) for p in params]
        H = torch.empty(B, T, cell.F, cell.N, dtype=torch.float32, device=dev)
        # execution-path selection (fp32 sparse precision): tell the library whether backward will want dX, let it choose,
        # and remember the choice so that backward runs on the path whose saved state this forward wrote
        cell.sync_options()
        cell.set_option('path', -1)
        cell.set_option('need_dx', int(ctx.needs_input_grad[2]))
        sb, fb, _ = cell.workspace(B, T, False)
        saved = _bytes(sb, dev)
        ws = _bytes(fb, dev)
        st = _fill_struct(cell.slots, p32)
        _lib.check(_lib.lib().gcrnn_cell_forward(cell.ptr, C.byref(st), _ptr(X32), _ptr(h32), _ptr(H), _ptr(saved),
                                                 saved.numel(), _ptr(ws), ws.numel(), B, T, _stream(dev)), 'cell_forward')
        ctx.cell = cell
        ctx.last_only = bool(last_only)
        ctx.path = cell.get_option('last_path')
        ctx.types = (X.dtype, h0.dtype, [p.dtype for p in params], [p.shape for p in params])
        ctx.save_for_backward(X32, h32, H, saved, *p32)
        if last_only:                  # only H[:, T-1] leaves; every state stays in `H` for the reverse sweep
            out = H[:, T - 1]
            return out.to(X.dtype) if X.dtype != torch.float32 else out.clone()
        return H.to(X.dtype) if X.dtype != torch.float32 else H

    @staticmethod
    def backward(ctx, dH):
        X32, h32, H, saved, *p32 = ctx.saved_tensors
        cell = ctx.cell
        dev = X32.device
        B, T = int(X32.shape[0]), int(X32.shape[1])
        xd, hd, pds, pshapes = ctx.types
        need_x, need_h = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        dH32 = _f32(dH, dev)              # [B,T,F,N], or [B,F,N] in last-state-only mode
        # one flat fp32 bucket for every parameter gradient: the kernels accumulate straight into it and the
        # data-parallel all-reduce (if a process group is active) runs on it in one call
        sizes = [p.numel() for p in p32]
        bucket = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        views = list(torch.split(bucket, sizes))
        gst = _fill_struct(cell.slots, views)
        pst = _fill_struct(cell.slots, p32)
        dX = torch.empty_like(X32) if need_x else None
        dh0 = torch.empty_like(h32) if need_h else None
        cell.sync_options()
        cell.set_option('path', ctx.path)
        cell.set_option('dh_last_only', int(ctx.last_only))
        _, _, bb = cell.workspace(B, T, need_x or need_h)
        ws = _bytes(bb, dev)
        _lib.check(_lib.lib().gcrnn_cell_backward(cell.ptr, C.byref(pst), _ptr(X32), _ptr(h32), _ptr(H), _ptr(dH32),
                                                  _ptr(saved), saved.numel(), C.byref(gst), _ptr(dX), _ptr(dh0), _ptr(ws),
                                                  ws.numel(), B, T, _stream(dev)), 'cell_backward')
        cell.set_option('path', -1)
        cell.set_option('dh_last_only', 0)
        _dist.reduce_cell_bucket(bucket)          # only in dist.enable(reduce_in_backward=True) mode: cell parameters only, SUM
        grads = [v.reshape(s).to(d) for v, s, d in zip(views, pshapes, pds)]
        return (None, None, dX.to(xd) if need_x else None, dh0.to(hd) if need_h else None, *grads)


def gated_gcrnn(cell: CellHandle, X, h0, params: Sequence[torch.Tensor], last_state_only: bool = False):
    """H = recurrence(X, h0) for the configuration in ``cell``; ``params`` ordered as ``cell.slots``.

    ``last_state_only=True`` returns H[:, T-1] ([B,F,N]) only: no [B,T,F,N] output or output-gradient tensor is handed to
    autograd (the classification readout uses nothing else: Modules/architectures.py:1841-1850)."""
    return _CellFn.apply(cell, bool(last_state_only), X, h0, *params)
