"""Import the UNMODIFIED reference from /root/reference (build container only).

Test infrastructure.  The reference tree does not exist on the GPU box, so
nothing that runs there may import this module; it is used by
``oracle/make_golden.py`` (fixture generation), by CPU-side tests that are
skipped when the tree is absent, and by ``bench.py --impl reference`` when a
copy of the reference was installed under ``baseline/_ref``.

Shims: gensim stub (Utils/dataTools.py:1001 imports it),
``np.int`` alias (Modules/train_rnn.py:134), no bytecode writes (read-only tree).
"""
import os
import sys
import types
import warnings

REF_ROOTS = [os.environ.get('GCRNN_REFERENCE_ROOT', ''), '/root/reference',
             os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')]


def reference_root():
    for r in REF_ROOTS:
        if r and os.path.isfile(os.path.join(r, 'Utils', 'graphML.py')):
            return r
    return None


def available() -> bool:
    return reference_root() is not None


def load():
    """Return the reference's ``Utils.graphML`` module."""
    root = reference_root()
    if root is None:
        raise ImportError('reference tree not found (looked in %s)' % REF_ROOTS)
    sys.dont_write_bytecode = True
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    sys.modules.setdefault('gensim', types.ModuleType('gensim'))
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import Utils.graphML as gml
    return gml


def load_architectures():
    load()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import Modules.architectures as archit
    return archit
