"""CPU: the C-ABI library loads and exports every symbol include/gcrnn_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from gated_gcrnns_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'gcrnn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gcrnn_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/gcrnn_b200.h but not exported'
    assert sorted(_lib.SYMBOLS) == names


def test_loads_and_reports_version():
    L = _lib.lib()
    assert L.gcrnn_abi_version() == 1
    assert isinstance(L.gcrnn_last_error(), bytes)


def test_struct_layout_matches_header():
    # 3 + 5*2 + 5*2 + 2*2 pointers, 9 int32 in the descriptor
    assert ctypes.sizeof(_lib.CellParams) == 8 * (3 + 10 + 10 + 4)
    assert ctypes.sizeof(_lib.CellDesc) == 4 * 9


def test_workspace_query_is_pure_host_arithmetic():
    # NULL graph -> error code, not a crash; message available
    L = _lib.lib()
    rc = L.gcrnn_graph_info(None, None, None, None, None)
    assert rc != 0 and b'null graph' in L.gcrnn_last_error()
