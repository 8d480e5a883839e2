"""Device-side graph shift operator handles (what ``addGSO(S)`` turns into).

The reference keeps the dense ``S`` tensor as a module attribute
(Utils/graphML.py:1166-1173, :2237-2244) and multiplies by it.  Here ``S`` is
converted once per (tensor, device) into the library's sparse gather forms
(+ bf16 dense tiles for the tensor-core path) and cached.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib


class Graph:
    """Owns one ``gcrnn_graph*``."""

    def __init__(self, handle, N, E, device_index, dense):
        self.handle = handle
        self.N, self.E, self.device_index, self.dense = N, E, device_index, dense
        self._fin = weakref.finalize(self, _destroy, handle)

    @property
    def ptr(self):
        return C.c_void_p(self.handle)

    def set_option(self, name: str, value: int):
        """Tuning switch for calls made on the graph handle itself (the debug shift GEMM), and ``'reorder'`` (0 never, 1 when it
        pays, 2 always): the library-owned node renumbering of the fused sparse path (include/gcrnn_b200.h)."""
        _lib.check(_lib.lib().gcrnn_graph_set_option(self.ptr, name.encode(), int(value)), 'graph_set_option')

    def get_option(self, name: str) -> int:
        """Also reads ``'reordered'`` and ``'tile_rows_before_x100'`` / ``'tile_rows_after_x100'`` (locality diagnostics)."""
        v = C.c_int32()
        _lib.check(_lib.lib().gcrnn_graph_get_option(self.ptr, name.encode(), C.byref(v)), 'graph_get_option')
        return v.value

    def info(self):
        n, e, nnz, na = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().gcrnn_graph_info(self.ptr, C.byref(n), C.byref(e), C.byref(nnz), C.byref(na)), 'graph_info')
        return dict(N=n.value, E=e.value, nnz=nnz.value, nnz_att=na.value)


def _destroy(handle):
    try:
        _lib.lib().gcrnn_graph_destroy(C.c_void_p(handle))
    except Exception:
        pass


def _device_index(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise _lib.GcrnnError(f'gated_gcrnns_b200 runs on CUDA devices only (got {device}); there is no CPU fallback')
    return torch.cuda.current_device() if device.index is None else device.index


def from_dense(S: torch.Tensor, device, keep_dense=False) -> Graph:
    """S: [E,N,N] dense tensor on any device / float dtype."""
    assert S.dim() == 3 and S.shape[1] == S.shape[2]
    E, N = int(S.shape[0]), int(S.shape[1])
    host = np.ascontiguousarray(S.detach().to('cpu', torch.float32).numpy())
    di = _device_index(device)
    out = C.c_void_p()
    _lib.check(_lib.lib().gcrnn_graph_create_dense(C.byref(out), N, E, host.ctypes.data_as(C.c_void_p),
                                                  1 if keep_dense else 0, di), 'graph_create_dense')
    return Graph(out.value, N, E, di, bool(keep_dense))


def from_csr(ops, N, device) -> Graph:
    """ops: list (one per edge feature) of (rowptr int64[N+1], colidx int32[nnz], vals float32[nnz]) numpy arrays."""
    E = len(ops)
    keep = []
    rp, ci, vv = (C.c_void_p * E)(), (C.c_void_p * E)(), (C.c_void_p * E)()
    for e, (r, c, v) in enumerate(ops):
        r = np.ascontiguousarray(r, dtype=np.int64)
        c = np.ascontiguousarray(c, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        assert r.shape == (N + 1,) and c.shape == v.shape == (int(r[-1]),)
        keep += [r, c, v]
        rp[e], ci[e], vv[e] = r.ctypes.data, c.ctypes.data, v.ctypes.data
    di = _device_index(device)
    out = C.c_void_p()
    _lib.check(_lib.lib().gcrnn_graph_create_csr(C.byref(out), N, E, rp, ci, vv, di), 'graph_create_csr')
    return Graph(out.value, N, E, di, False)


def from_sparse_tensor(S: torch.Tensor, device) -> Graph:
    """S: torch sparse tensor [N,N] or [1,N,N] (graphs too large for the reference's dense S)."""
    if S.dim() == 3:
        assert S.shape[0] == 1, 'sparse GSOs are supported for E == 1'
        S = S[0] if S.layout == torch.sparse_coo else S
    csr = S.detach().cpu().to_sparse_csr() if S.layout != torch.sparse_csr else S.detach().cpu()
    N = int(csr.shape[-1])
    return from_csr([(csr.crow_indices().numpy(), csr.col_indices().numpy(), csr.values().to(torch.float32).numpy())],
                    N, device)


_cache = {}


def gso_shape(S):
    """(E, N) of whatever ``addGSO`` was given."""
    if S.layout == torch.strided:
        assert S.dim() == 3, 'the GSO must be edge_features x nodes x nodes'
        return int(S.shape[0]), int(S.shape[1])
    return (int(S.shape[0]), int(S.shape[1])) if S.dim() == 3 else (1, int(S.shape[0]))


def get(S: torch.Tensor, device, keep_dense=False) -> Graph:
    """Cached handle for the GSO tensor ``S`` on ``device`` (sub-modules sharing one S share one handle)."""
    di = _device_index(device)
    key = (id(S), getattr(S, '_version', 0), tuple(S.shape), di, bool(keep_dense))
    hit = _cache.get(key)
    if hit is not None and hit[0]() is S:
        return hit[1]
    g = from_dense(S, device, keep_dense) if S.layout == torch.strided else from_sparse_tensor(S, device)
    # the entry (and with it the device copies of the operator) goes away as soon as the GSO tensor dies
    _cache[key] = (weakref.ref(S, lambda _ref, key=key: _cache.pop(key, None)), g)
    return g
