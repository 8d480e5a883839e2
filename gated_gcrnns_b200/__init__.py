"""gated_gcrnns_b200 — B200-native gated GCRNN recurrence behind the reference's module API.

    import gated_gcrnns_b200 as gg
    gg.install()                       # rebinds Utils.graphML.{LSIGF,GraphFilter,GraphAttentional,GGCRNNCell}
    # ... the reference's Modules/architectures.py, train_rnn.py, train_rnn_quake.py now run on the CUDA path

See DESIGN.md / INTEGRATION.md.  The C ABI is include/gcrnn_b200.h.
"""
from .functional import LSIGF, graph_attention_relu, gated_gcrnn, set_precision, get_precision, cell_param_slots
from .modules import GraphFilter, GraphAttentional, GGCRNNCell
from . import dist, graph, graphs, options, readout, train
from ._lib import GcrnnError, LIB_PATH

__all__ = ['LSIGF', 'GraphFilter', 'GraphAttentional', 'GGCRNNCell', 'install', 'uninstall', 'set_precision',
           'get_precision', 'dist', 'graph', 'graphs', 'options', 'readout', 'train', 'GcrnnError']

_PATCHED = ('LSIGF', 'GraphFilter', 'GraphAttentional', 'GGCRNNCell')
_originals = {}
_arch_undo = {}


def install(graphML_module=None, architectures_module=None):
    """Rebind the four hot-path names on the reference's ``Utils.graphML`` module.

    ``Modules/architectures.py`` looks them up as ``gml.<Name>`` at call time (architectures.py:6, :1521), so
    nothing in the reference needs editing.  Everything else in ``Utils.graphML`` is left untouched.

    ``architectures_module`` (the imported ``Modules.architectures``, or ``True`` to import it) additionally applies the
    fused readouts of ``gated_gcrnns_b200.readout``: last-state-only recurrence for the classifier and the batched per-node
    MLP for the regression architecture.  Their outputs are identical to the reference's."""
    if graphML_module is None:
        import Utils.graphML as graphML_module  # the reference must be importable (sys.path)
    import sys
    me = sys.modules[__name__]
    for n in _PATCHED:
        _originals.setdefault((id(graphML_module), n), getattr(graphML_module, n))
        setattr(graphML_module, n, getattr(me, n))
    if architectures_module is not None and architectures_module is not False:
        if architectures_module is True:
            import Modules.architectures as architectures_module
        if id(architectures_module) not in _arch_undo:
            _arch_undo[id(architectures_module)] = readout.patch_architectures(architectures_module)
    return graphML_module


def uninstall(graphML_module=None, architectures_module=None):
    if graphML_module is None:
        import Utils.graphML as graphML_module
    for n in _PATCHED:
        o = _originals.pop((id(graphML_module), n), None)
        if o is not None:
            setattr(graphML_module, n, o)
    if architectures_module is not None:
        for cls, name, orig in _arch_undo.pop(id(architectures_module), []):
            setattr(cls, name, orig)
