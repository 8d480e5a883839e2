"""Tuning switches of the CUDA library, host side.

The library itself keeps no process-wide mutable state: every switch lives on a cell (or graph) handle
(``gcrnn_cell_set_option`` / ``gcrnn_graph_set_option`` in include/gcrnn_b200.h).  This module is the Python-level
default table: ``set(name, value)`` changes what every ``CellHandle`` applies to its own handle before its next call.
They exist for tests and A/B measurements; the defaults are the fast paths.
"""
from __future__ import annotations

DEFAULTS = {
    'bwd_fused': 1, 'sparse_fused': 1, 'sparse_v2': 63, 'sparse_v2_rows_bps': 2, 'sparse_v2_fuse_dpre': 1,
    'sparse_v2_tc': 1, 'sparse_v2_bps': 2, 'graph_capture': 1, 'gate_fq8': 2, 'gemm_pair': 1, 'fwd_fused': 0, 'persist': 1,
}
_values = dict(DEFAULTS)
_version = 0


def set(name: str, value: int) -> int:      # noqa: A001  (mirrors the C entry point's name)
    """Set a switch for every handle's next call; returns the previous value."""
    global _version
    if name not in DEFAULTS:
        raise KeyError(f'unknown option {name!r}; known: {sorted(DEFAULTS)}')
    old = _values[name]
    _values[name] = int(value)
    _version += 1
    return old


def get(name: str) -> int:
    return _values[name]


def reset():
    global _version
    _values.update(DEFAULTS)
    _version += 1


def version() -> int:
    return _version


def items():
    return list(_values.items())
