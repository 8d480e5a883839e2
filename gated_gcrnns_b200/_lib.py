"""ctypes binding of libgcrnn_b200.so (the C ABI in include/gcrnn_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this
module raises.  Build it with ``python -m gated_gcrnns_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgcrnn_b200.so')

SPATIAL = {None: 0, 'node': 1, 'edge': 2}
PREC_FP32, PREC_BF16_TC, PREC_BF16X2_TC = 0, 1, 2

# every symbol include/gcrnn_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    'gcrnn_abi_version', 'gcrnn_last_error', 'gcrnn_debug_launch_count', 'gcrnn_debug_shift_gemm', 'gcrnn_debug_edge_relu_masks',
    'gcrnn_graph_set_option', 'gcrnn_graph_get_option',
    'gcrnn_graph_create_csr', 'gcrnn_graph_create_dense', 'gcrnn_graph_destroy', 'gcrnn_graph_info',
    'gcrnn_lsigf_workspace_bytes', 'gcrnn_lsigf_forward', 'gcrnn_lsigf_backward',
    'gcrnn_gat_workspace_bytes', 'gcrnn_gat_forward', 'gcrnn_gat_backward',
    'gcrnn_cell_create', 'gcrnn_cell_destroy', 'gcrnn_cell_set_option', 'gcrnn_cell_get_option', 'gcrnn_cell_workspace_bytes',
    'gcrnn_cell_forward', 'gcrnn_cell_backward',
    'gcrnn_comm_unique_id', 'gcrnn_comm_create', 'gcrnn_comm_destroy', 'gcrnn_allreduce_sum',
    'gcrnn_build_knn_csr', 'gcrnn_data_diffusion',
]


class CellDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ('G', 'F', 'Kin', 'Kst', 'E', 'time_gating', 'spatial_gating', 'bias', 'precision')]


_P = C.c_void_p


class CellParams(C.Structure):
    _fields_ = [
        ('weight_A', _P), ('weight_B', _P), ('bias', _P),
        ('t_weight_A', _P * 2), ('t_weight_B', _P * 2), ('t_bias', _P * 2),
        ('t_mlp_w', _P * 2), ('t_mlp_b', _P * 2),
        ('n_weight_A', _P * 2), ('n_weight_B', _P * 2), ('n_bias', _P * 2),
        ('n_head_w', _P * 2), ('n_head_b', _P * 2),
        ('e_mixer', _P * 2), ('e_weight', _P * 2),
    ]


_lib = None


class GcrnnError(RuntimeError):
    pass


def lib():
    """Load the library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise GcrnnError(
            f'{LIB_PATH} not found: the CUDA library is not built. Run `python -m gated_gcrnns_b200.build`. '
            'There is no CPU or PyTorch fallback for this path.')
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.gcrnn_abi_version.restype = C.c_int
    L.gcrnn_last_error.restype = C.c_char_p
    L.gcrnn_debug_launch_count.restype = C.c_uint64
    L.gcrnn_debug_edge_relu_masks.argtypes = [_P, _P, C.c_size_t, C.c_int64, C.c_int64, _P, _P]
    L.gcrnn_debug_shift_gemm.argtypes = [_P, C.c_int32, _P, C.c_int64, C.c_int32, _P, C.c_int32, _P, _P]
    L.gcrnn_graph_set_option.argtypes = [_P, C.c_char_p, C.c_int32]
    L.gcrnn_graph_get_option.argtypes = [_P, C.c_char_p, C.POINTER(C.c_int32)]
    L.gcrnn_lsigf_workspace_bytes.restype = C.c_size_t
    L.gcrnn_lsigf_workspace_bytes.argtypes = [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int64]
    L.gcrnn_gat_workspace_bytes.restype = C.c_size_t
    L.gcrnn_gat_workspace_bytes.argtypes = [_P, C.c_int32, C.c_int32, C.c_int64]
    L.gcrnn_graph_create_csr.argtypes = [C.POINTER(_P), C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(_P),
                                         C.POINTER(_P), C.c_int32]
    L.gcrnn_graph_create_dense.argtypes = [C.POINTER(_P), C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32]
    L.gcrnn_graph_destroy.argtypes = [_P]
    L.gcrnn_graph_info.argtypes = [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64)]
    L.gcrnn_lsigf_forward.argtypes = [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P,
                                      C.c_size_t, _P]
    L.gcrnn_lsigf_backward.argtypes = [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P,
                                       C.c_size_t, _P]
    L.gcrnn_gat_forward.argtypes = [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int64, _P, C.c_size_t, _P]
    L.gcrnn_gat_backward.argtypes = [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int64, _P,
                                     C.c_size_t, _P]
    L.gcrnn_cell_create.argtypes = [C.POINTER(_P), C.POINTER(CellDesc), _P]
    L.gcrnn_cell_destroy.argtypes = [_P]
    L.gcrnn_cell_set_option.argtypes = [_P, C.c_char_p, C.c_int32]
    L.gcrnn_cell_get_option.argtypes = [_P, C.c_char_p, C.POINTER(C.c_int32)]
    L.gcrnn_cell_workspace_bytes.argtypes = [_P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.gcrnn_cell_forward.argtypes = [_P, C.POINTER(CellParams), _P, _P, _P, _P, C.c_size_t, _P, C.c_size_t,
                                     C.c_int64, C.c_int64, _P]
    L.gcrnn_cell_backward.argtypes = [_P, C.POINTER(CellParams), _P, _P, _P, _P, _P, C.c_size_t,
                                      C.POINTER(CellParams), _P, _P, _P, C.c_size_t, C.c_int64, C.c_int64, _P]
    L.gcrnn_comm_unique_id.argtypes = [_P]
    L.gcrnn_comm_create.argtypes = [C.POINTER(_P), _P, C.c_int32, C.c_int32, C.c_int32]
    L.gcrnn_comm_destroy.argtypes = [_P]
    L.gcrnn_allreduce_sum.argtypes = [_P, _P, C.c_int64, _P]
    L.gcrnn_build_knn_csr.argtypes = [C.c_int32, C.c_int32, _P, C.c_float, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.gcrnn_data_diffusion.argtypes = [_P, _P, _P, _P, C.c_int64, C.c_int32, _P]
    if L.gcrnn_abi_version() != 2:
        raise GcrnnError('libgcrnn_b200.so ABI version mismatch')
    _lib = L
    return L


def check(rc, what=''):
    if rc != 0:
        msg = lib().gcrnn_last_error().decode(errors='replace')
        raise GcrnnError(f'{what} failed ({rc}): {msg}')
