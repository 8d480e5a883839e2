"""Whole-step CUDA-graph capture for the small (launch-bound) configurations — SURVEY.md §8f rank 3.

At the reference's own sizes (N = 59 / 80 nodes, B = 100: cfg1 / cfg2) a training step is hundreds of tiny kernels: the
recurrence, the readout, the loss (``Utils/miscTools.py:112-119`` / cross entropy), autograd's backward and the optimiser
update (``Modules/train_rnn.py:231-276``).  The library already replays the CELL's forward / backward as CUDA graphs
(``csrc/api.cu``); ``GraphedStep`` captures the WHOLE step — node-reordering gather, forward, loss, backward, gradient all-reduce
if attached, optimiser — into one graph and replays it with new data copied into static buffers:

    step = GraphedStep(model, loss_fn, optimizer, example_x, example_h0, example_y)
    for x, y in batches:
        loss = step(x, h0, y)          # one cudaGraphLaunch; `loss` is a device scalar (no host sync)

Everything inside the step must be capturable: an optimiser built with ``capturable=True`` (torch.optim.Adam / AdamW / SGD), device-resident inputs (the reference loop creates ``h0`` on the CPU every
step — pass a device tensor here), no ``.item()`` / host reads, shapes fixed.  The library's own kernels are launched on the
capturing stream (its internal per-call graph cache steps aside while a capture is in progress).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedStep:
    """One training step (zero_grad, forward, loss, backward, optimizer.step) as a single replayable CUDA graph."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer, *example_inputs: torch.Tensor,
                 target: torch.Tensor, order: Optional[Sequence[int]] = None, warmup: int = 3):
        assert all(t.is_cuda for t in example_inputs) and target.is_cuda, 'GraphedStep needs device-resident example tensors'
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        dev = target.device
        self.static_in = [t.detach().clone() for t in example_inputs]
        self.static_tgt = target.detach().clone()
        # node re-ordering of the reference loop (xTrain[:, :, order], train_rnn.py:234) as part of the captured step
        self.order = None if order is None else torch.as_tensor(list(order), dtype=torch.long, device=dev)
        self.loss = None
        self.output = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                 # allocator warm-up, lazy handle creation, optimizer state
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def _body(self):
        self.optimizer.zero_grad(set_to_none=True)
        ins = list(self.static_in)
        if self.order is not None:
            ins[0] = ins[0].index_select(-1, self.order)
        self.output = self.model(*ins)
        self.loss = self.loss_fn(self.output, self.static_tgt)
        self.loss.backward()
        self.optimizer.step()

    def __call__(self, *inputs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.static_tgt.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.loss
