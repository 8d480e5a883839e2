"""Loader for tests/golden/*.npz (fixtures written by oracle/make_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SG = {0: None, 1: 'node', 2: 'edge'}


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    out = {}
    for k in z.files:
        if '::' in k:
            grp, kk = k.split('::', 1)
            out.setdefault(grp, {})[kk] = z[k]
        else:
            out[k] = z[k]
    return out


def cell_meta(case):
    G, F, Kin, Kst, T, B, tg, sg, bias, seed, E = [int(v) for v in case['meta']]
    return dict(G=G, F=F, Kin=Kin, Kst=Kst, T=T, B=B, time_gating=bool(tg), spatial_gating=SG[sg],
                bias=bool(bias), seed=seed, E=E)


def t64(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)
