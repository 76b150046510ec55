"""Optimiser side of the training step (reference: train.py:36-38, :118-120 --
``torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-3)``) -- SURVEY.md 8(f) rank 3.

The two position tables ``x_embed`` / ``y_embed`` (model.py:204-205) hold 2 x 65536 x G of the
model's parameters but a step reads at most B rows of each.  Dense Adam still has to rewrite all
of them every step (weight decay moves every row), which makes the tables the dominant non-CNN
cost of a step: 7 arrays x 411 MB of HBM traffic at G = 785 plus a dense gradient that is zero
almost everywhere.  ``LazyEmbeddingAdam`` keeps the exact trajectory of dense Adam but defers
every row until it is next read: ``csrc/optim.cu`` replays the missed steps in registers, with
the arithmetic of its own dense kernel, so the tables equal the dense result bit for bit at any
point where they are observed (forward gathers, ``flush()``, ``state_dict()``).

    opt = TrainOptimizer(model, lr=1e-4, weight_decay=1e-3)     # stock Adam + lazy tables
    opt.zero_grad(); loss = model(batch); loss.backward(); opt.step()

No CPU fallback.  Works with ``graphs.GraphedTrainStep``: the catch-up inside the captured forward
reads the step counter from device memory, and the graphed step re-arms the recorded
(position, gradient) pair after every replay; ``step()`` itself runs outside the graph.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check, load, ptr, stream_ptr

__all__ = ["LazyEmbeddingAdam", "TrainOptimizer", "adam_dense_step"]


class _CoefTable:
    """Per-step Adam scalars on the device (one 32-byte record per step)."""

    def __init__(self, device, max_steps: int):
        self.max_steps = int(max_steps)
        self.buf = torch.zeros(load().mclst_adam_coef_bytes(self.max_steps), dtype=torch.uint8, device=device)

    def set_step(self, step: int, lr, betas, eps, weight_decay):
        with torch.cuda.device(self.buf.device):
            check(load().mclst_adam_set_step(ptr(self.buf), self.max_steps, int(step), float(lr),
                                             float(betas[0]), float(betas[1]), float(eps),
                                             float(weight_decay), stream_ptr()), "adam_set_step")


def adam_dense_step(param: torch.Tensor, grad: Optional[torch.Tensor], exp_avg: torch.Tensor,
                    exp_avg_sq: torch.Tensor, coef: _CoefTable, step: int) -> None:
    """One dense Adam step with the scalars recorded for ``step`` (testing / small tensors)."""
    assert param.is_cuda and param.dtype == torch.float32 and param.is_contiguous()
    with torch.cuda.device(param.device):
        check(load().mclst_adam_dense(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq),
                                      param.numel(), ptr(coef.buf), int(step), stream_ptr()),
              "adam_dense")


class LazyEmbeddingAdam:
    """Exact Adam(+L2 weight decay) for embedding tables indexed by the columns of ``position``
    (column i of ``position`` [B, len(tables)] selects the rows of ``tables[i]``)."""

    def __init__(self, tables: Sequence[nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, max_steps: int = 1 << 20):
        if not tables or len(tables) > 2:
            raise ValueError("LazyEmbeddingAdam handles the one or two position tables of the model")
        for t in tables:
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
                raise _lib.MclstError("LazyEmbeddingAdam: tables must be contiguous float32 CUDA matrices")
        self.tables = list(tables)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        dev = tables[0].device
        self.coef = _CoefTable(dev, max_steps)
        self.exp_avg = [torch.zeros_like(t) for t in tables]
        self.exp_avg_sq = [torch.zeros_like(t) for t in tables]
        self.last = [torch.zeros(t.shape[0], dtype=torch.int32, device=dev) for t in tables]
        self._first = torch.empty(max(t.shape[0] for t in tables), dtype=torch.int32, device=dev)
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.steps_done = 0
        # the same counter on the device: kernels captured in a CUDA graph (the catch-up inside a
        # graphed forward) read it at replay time
        self._steps_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._pending: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        for i, t in enumerate(self.tables):
            t._mclst_lazy = self                                  # picked up by model.embed_add

    # -- the three moments a row is touched
    def _rows(self, position: torch.Tensor, d_out: Optional[torch.Tensor]):
        position = position.detach().float()
        if position.dim() != 2 or position.shape[1] < len(self.tables):
            raise ValueError("position must be [B, 2]")
        if position.stride(1) != 1:
            position = position.contiguous()
        if d_out is not None:
            d_out = d_out.detach().float()
            if d_out.stride(-1) != 1:
                d_out = d_out.contiguous()
        lib = load()
        with torch.cuda.device(position.device), torch.no_grad():
            for i, t in enumerate(self.tables):
                check(lib.mclst_adam_lazy_rows(ptr(t), ptr(self.exp_avg[i]), ptr(self.exp_avg_sq[i]),
                                               ptr(self.last[i]), ptr(self._first), t.shape[0], t.shape[1],
                                               ptr(position), position.stride(0), i, position.shape[0],
                                               ptr(d_out), d_out.stride(0) if d_out is not None else 0,
                                               ptr(self.coef.buf), self.steps_done, ptr(self._steps_dev),
                                               ptr(self._err), stream_ptr()),
                      "adam_lazy_rows")

    def catch_up(self, position: torch.Tensor) -> None:
        """Bring the rows a batch is about to read to the current step (called by embed_add)."""
        # (under graph capture the launch must exist even at step 0: the replay follows the device counter)
        if self.steps_done > 0 or torch.cuda.is_current_stream_capturing():
            self._rows(position, None)

    def record(self, position: torch.Tensor, d_out: torch.Tensor) -> None:
        """Gradient of the embed-add output for the batch at ``position`` (called from backward)."""
        if self._pending is not None:
            raise _lib.MclstError("LazyEmbeddingAdam: two backward passes without a step in between "
                                  "(gradient accumulation is not supported for the lazy tables)")
        self._pending = (position.detach(), d_out.detach())

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._pending = None

    def set_hyper(self, lr=None, betas=None, eps=None, weight_decay=None) -> None:
        """Hyper-parameters of the NEXT steps (an LR scheduler changing ``param_groups`` ends up here)."""
        if lr is not None:
            self.lr = float(lr)
        if betas is not None:
            self.betas = (float(betas[0]), float(betas[1]))
        if eps is not None:
            self.eps = float(eps)
        if weight_decay is not None:
            self.weight_decay = float(weight_decay)

    def step(self) -> None:
        if self._pending is None:
            raise _lib.MclstError("LazyEmbeddingAdam.step(): no gradient recorded (run forward + backward first)")
        position, d_out = self._pending
        self.coef.set_step(self.steps_done + 1, self.lr, self.betas, self.eps, self.weight_decay)
        self._rows(position, d_out)
        self.steps_done += 1
        self._steps_dev.add_(1)
        self._pending = None

    def flush(self) -> None:
        """Every row of every table to the current step (before the tables are read as a whole)."""
        if self.steps_done == 0:
            return
        lib = load()
        with torch.cuda.device(self.tables[0].device), torch.no_grad():
            for i, t in enumerate(self.tables):
                check(lib.mclst_adam_lazy_flush(ptr(t), ptr(self.exp_avg[i]), ptr(self.exp_avg_sq[i]),
                                                ptr(self.last[i]), t.shape[0], t.shape[1], ptr(self.coef.buf),
                                                self.steps_done, stream_ptr()), "adam_lazy_flush")

    def check_positions(self) -> None:
        """Raise if any recorded position was outside the tables (synchronises)."""
        if int(self._err.item()):
            raise IndexError("position index out of range for x_embed / y_embed")

    def mark_all_current(self) -> None:
        """Declare every row current (after the table VALUES were replaced from outside, e.g.
        ``load_state_dict``): no missed step may be replayed onto the new values."""
        for l in self.last:
            l.fill_(self.steps_done)

    def state_dict(self) -> dict:
        self.flush()
        return {"step": self.steps_done, "exp_avg": [t.clone() for t in self.exp_avg],
                "exp_avg_sq": [t.clone() for t in self.exp_avg_sq],
                "hyper": {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}}

    def load_state_dict(self, sd: dict) -> None:
        """Moments and step counter of ``state_dict()``; the per-step scalars of the steps already
        done are never read again once every row is current, so only the counter is restored."""
        if len(sd["exp_avg"]) != len(self.tables):
            raise ValueError("LazyEmbeddingAdam.load_state_dict: table count differs")
        self.flush()
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)
        self.steps_done = int(sd["step"])
        self._steps_dev.fill_(self.steps_done)
        if self.steps_done >= self.coef.max_steps:
            raise ValueError("LazyEmbeddingAdam.load_state_dict: step beyond max_steps")
        self.set_hyper(**sd.get("hyper", {}))
        self._pending = None
        self.mark_all_current()


class TrainOptimizer:
    """``torch.optim.Adam`` (stock) for every parameter except the position tables, which go to
    ``LazyEmbeddingAdam`` with the same hyper-parameters; the calls of train.py:36-38 unchanged.
    ``model.state_dict()`` and direct calls of ``model.x_embed`` / ``model.y_embed`` flush the tables
    first; the model's own forward only catches up the rows it reads."""

    def __init__(self, model: nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-3):
        tables = [model.x_embed.weight, model.y_embed.weight]
        ids = {id(t) for t in tables}
        rest: List[nn.Parameter] = [p for p in model.parameters() if id(p) not in ids]
        self.dense = torch.optim.Adam(rest, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.lazy = LazyEmbeddingAdam(tables, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        model.register_state_dict_pre_hook(lambda module, prefix, keep_vars: self.lazy.flush())
        # ``model.load_state_dict`` after training has started (resume, load-best): apply every
        # deferred step to the OLD values first, so that the loaded rows are current by construction
        # (``last == steps_done``) and no stale step is ever replayed onto them -- stock Adam leaves
        # loaded weights untouched too
        model.register_load_state_dict_pre_hook(self._before_model_load)
        # reference-style evaluation calls the embedding modules directly (evel_her2st.py:52-57:
        # ``model.x_embed(x)``), bypassing embed_add's per-row catch-up: bring everything current first
        for emb in (model.x_embed, model.y_embed):
            emb.register_forward_pre_hook(lambda module, args: self.lazy.flush())

        # the tables appear as their own param group so that ``get_lr(optimizer)`` (train.py:41) and
        # LR schedulers see ONE optimizer; stock Adam never steps them (their .grad stays None)
        g0 = self.dense.param_groups[0]
        self.dense.add_param_group({"params": tables, **{k: g0[k] for k in ("lr", "betas", "eps", "weight_decay")}})

    def _before_model_load(self, module, state_dict, prefix, *args) -> None:
        self.lazy.flush()
        self.lazy.mark_all_current()

    @property
    def param_groups(self):
        return self.dense.param_groups

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.dense.zero_grad(set_to_none=set_to_none)
        self.lazy.zero_grad()

    def step(self) -> None:
        g = self.dense.param_groups[-1]                  # the tables' group: scheduler-visible lr
        self.lazy.set_hyper(g["lr"], g["betas"], g["eps"], g["weight_decay"])
        self.dense.step()
        self.lazy.step()

    def flush(self) -> None:
        self.lazy.flush()

    def state_dict(self) -> dict:
        return {"dense": self.dense.state_dict(), "lazy": self.lazy.state_dict()}

    def load_state_dict(self, sd: dict) -> None:
        self.dense.load_state_dict(sd["dense"])
        self.lazy.load_state_dict(sd["lazy"])
