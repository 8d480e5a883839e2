"""Readouts fused onto the recurrence — host-side drop-ins for the two architecture ``forward``s.

``install(..., architectures_module)`` applies them to the reference's ``Modules/architectures.py`` classes without editing the
reference:

  * ``GatedGCRNNforClassification`` (architectures.py:1647-1858) uses only ``H.select(1, -1)`` (:1844): its cell is switched to
    ``last_state_only`` — the library still computes every state (the reverse sweep needs them) but neither the ``[B,T,F,N]``
    output nor its gradient tensor is ever handed to autograd, and the backward kernels read one L2-resident zero slab instead
    of ``dH[:, t]`` for t < T-1.
  * ``GatedGCRNNforRegression.forward`` with ``mlpType='multipMlp'`` (:1613-1636) applies the SAME per-node MLP in a Python
    loop over the N nodes and concatenates (and starts from a CPU ``torch.empty(0)``, so it cannot even run on a GPU);
    ``regression_forward`` is the same function as ONE batched contraction over ``[B*T, N, F_h]``.
"""
from __future__ import annotations

import torch


def regression_forward(self, x, h0):
    """Drop-in for ``GatedGCRNNforRegression.forward`` (architectures.py:1607-1645): identical outputs; the per-node readout
    loop becomes one batched call."""
    batchSize, seqLength = x.shape[0], x.shape[1]
    H = self.stateGCRNN(x, h0)
    flatH = H.reshape(-1, self.F_h, self.N)                    # merge batch and time (architectures.py:1613)
    if self.F_o is None:                                       # MLP readout
        if self.mlpType == 'multipMlp':
            # reference: for i in range(N): y_i = outputNN(flatH[:, :, i]); cat over nodes; transpose; squeeze  (:1620-1632)
            flatY = self.outputNN(flatH.transpose(1, 2))       # [B*T, N, out]: nn.Linear broadcasts over the node dimension
            flatY = flatY.transpose(1, 2).squeeze()
        elif self.mlpType == 'oneMlp':
            flatY = self.outputNN(flatH.reshape(-1, self.F_h * self.N))
        else:
            raise ValueError(f'unknown mlpType {self.mlpType!r}')
    else:                                                      # GNN readout (SelectionGNN / AggregationGNN on our GraphFilter)
        flatY = self.outputNN(flatH)
    y = flatY.reshape(batchSize, seqLength, -1)
    return torch.unsqueeze(y, 2)


def patch_architectures(archit):
    """Apply both readout drop-ins to a ``Modules.architectures`` module; returns the undo list."""
    undo = []
    cls_c = getattr(archit, 'GatedGCRNNforClassification', None)
    if cls_c is not None:
        orig_init = cls_c.__init__

        def init(self, *a, **k):
            orig_init(self, *a, **k)
            cell = getattr(self, 'stateGCRNN', None)
            if cell is not None and hasattr(cell, 'last_state_only'):
                cell.last_state_only = True
        init.__wrapped__ = orig_init
        cls_c.__init__ = init
        undo.append((cls_c, '__init__', orig_init))
    cls_r = getattr(archit, 'GatedGCRNNforRegression', None)
    if cls_r is not None:
        undo.append((cls_r, 'forward', cls_r.forward))
        cls_r.forward = regression_forward
    return undo
