"""CUDA-graph capture of one training step of the hot path.

At the reference's batch sizes (128-1024 spots) a forward+backward of the path is ~300 small
kernel launches and is launch-latency bound ("tiny-problem regime").
Capturing it once and replaying removes the per-launch host cost; the kernels are unchanged.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch


class GraphedTrainStep:
    """``step(batch) -> loss`` with forward + backward replayed from a CUDA graph.

    ``model`` is a ``mclstexp_b200.model.mclSTExp_Attention`` (or any module built from this
    package's layers) whose ``forward(batch)`` returns the loss.  Gradients land in the
    parameters' ``.grad`` exactly as with eager ``loss.backward()``; call ``optimizer.step()``
    after ``step()`` as usual.  Batch shapes are fixed at capture time.  Drop references to any
    loss tensor of an earlier eager step before constructing this (a live autograd graph
    created on the default stream invalidates the capture)."""

    def __init__(self, model: torch.nn.Module, example_batch: Dict[str, torch.Tensor], warmup: int = 3,
                 defer_weight_grads: bool = True):
        from .model import deferred_weight_grads
        # position tables owned by optim.LazyEmbeddingAdam (TrainOptimizer): the captured forward
        # contains their row catch-up (which follows the optimiser's device-side step counter) and the
        # captured backward leaves (position, d_out) in static buffers for ``optimizer.step()``
        self.lazy = getattr(getattr(getattr(model, "x_embed", None), "weight", None), "_mclst_lazy", None)
        self.model = model
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.table_rows = getattr(getattr(model, "x_embed", None), "num_embeddings", None)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # weight / bias gradients of the Linear layers leave the dX chain: a parallel branch of the
        # graph that joins at the end of the backward pass (model.deferred_weight_grads)
        with torch.cuda.stream(side), deferred_weight_grads(defer_weight_grads):   # warm-up off the default stream
            for _ in range(warmup):
                for p in model.parameters():
                    p.grad = None
                if self.lazy is not None:
                    self.lazy.zero_grad()
                model(self.static).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        for p in model.parameters():       # backward inside the capture then ASSIGNS fresh .grad
            p.grad = None                  # buffers from the graph's pool: every replay overwrites them
        if self.lazy is not None:
            self.lazy.zero_grad()
        with torch.cuda.graph(self.graph), deferred_weight_grads(defer_weight_grads):
            self.loss = model(self.static)
            self.loss.backward()
        self._lazy_pending = self.lazy._pending if self.lazy is not None else None
        # the replay writes into THESE buffers; ``optimizer.zero_grad()`` (set_to_none=True is the
        # default, train.py:37) drops them from the parameters, so they are re-attached after
        # every replay -- otherwise optimizer.step() would silently skip every parameter
        self.grads = {p: p.grad for p in model.parameters() if p.grad is not None}

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        pos = batch.get("position")
        if pos is not None and self.table_rows is not None:
            lo, hi = pos.min(), pos.max()                  # nn.Embedding would raise here
            if bool((lo < 0) | (hi >= self.table_rows)):
                raise IndexError("position index out of range for x_embed / y_embed")
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        for p, g in self.grads.items():
            if p.grad is not g:
                p.grad = g
        if self.lazy is not None:                          # the replay refreshed these static buffers
            self.lazy._pending = self._lazy_pending
        return self.loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Provided for loops that call it on the step object; the replay overwrites every captured
        gradient anyway, so nothing needs clearing."""
