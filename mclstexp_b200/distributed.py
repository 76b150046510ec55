"""Multi-GPU forms of the two shardable pieces of the path.

One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for the exchanges.
The reference is single-process (train.py:107) -- these are new capabilities, specified by
the north star, whose results must equal the single-GPU results on the concatenated inputs.

Retrieval -- the bank (spot embeddings AND expression rows) is split into contiguous row
ranges, queries are replicated:
    local   : find_matches on the shard (+ index offset) and the distances of its own winners
    exchange: all-gather of (similarity, global index, distance) candidate lists  [R, Q, k]
    merge   : global top-k by (value desc, index asc) -- identical on every rank
    average : each rank accumulates w_j * expression[idx_j] for the winners it owns,
              all-reduce of the [Q, G] partial sums
Contrastive loss -- rank r owns batch rows [r*B/R, (r+1)*B/R):
    all-gather of the embeddings, three phases on the local rows separated by all-gathers of
    the per-row statistics (5 floats per row); gradients of the local rows need no exchange.

The host logic is backend-agnostic: ``CudaBackend`` drives libmclst_b200.so; the CPU test
suite injects an oracle-backed backend to exercise the sharding / gather / merge plumbing
under ``gloo`` with world_size 2 (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import WEIGHT_MODES, check, load, ptr, stream_ptr

__all__ = ["shard_bounds", "BankShard", "CudaBackend", "retrieve_sharded", "contrastive_loss_sharded",
           "RetrievalGrid", "make_retrieval_grid"]


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row ranges [(start, end)] by global index."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


@dataclass
class BankShard:
    spot_key: torch.Tensor          # [n_local, D] float32
    expression_key: torch.Tensor    # [n_local, G] float32 / float64
    index_offset: int               # global index of local row 0
    n_total: int
    expr_ready: Optional["torch.cuda.Event"] = None   # set when expression_key is still uploading

    @classmethod
    def from_full(cls, spot_key, expression_key, rank: int, world: int) -> "BankShard":
        lo, hi = shard_bounds(spot_key.shape[0], world)[rank]
        return cls(spot_key[lo:hi].contiguous(), expression_key[lo:hi].contiguous(), lo,
                   spot_key.shape[0])

    @classmethod
    def from_host(cls, spot_key: torch.Tensor, expression_key: torch.Tensor, index_offset: int,
                  n_total: int, device) -> "BankShard":
        """Upload one shard from (pinned) host tensors.  The expression rows -- most of the bytes,
        not needed before the average -- go on a side stream so that the copy runs underneath the
        shard's top-k kernels; ``retrieve_sharded`` waits for ``expr_ready``."""
        from .retrieval import _side_stream
        device = torch.device(device)
        sk = spot_key.to(device, non_blocking=True)
        side = _side_stream(device)
        with torch.cuda.stream(side):
            ek = expression_key.to(device, non_blocking=True)
            ready = side.record_event()
        ek.record_stream(torch.cuda.current_stream(device))
        return cls(sk, ek, index_offset, n_total, ready)


class CudaBackend:
    """libmclst_b200.so on the local device."""

    def local_topk(self, shard: BankShard, query: torch.Tensor, k: int, p: int, need_dist: bool):
        from .retrieval import find_matches_device
        n_loc = shard.spot_key.shape[0]
        kk = min(k, n_loc)
        Q = query.shape[0]
        dev = query.device
        val = torch.full((Q, k), float("-inf"), dtype=torch.float32, device=dev)
        idx = torch.full((Q, k), 2 ** 31 - 1, dtype=torch.int64, device=dev)
        dst = torch.full((Q, k), float("inf"), dtype=torch.float32, device=dev) if need_dist else None
        if kk == k:
            r = find_matches_device(shard.spot_key, query, k, index_offset=shard.index_offset,
                                    dist_p=p if need_dist else None)
            return (r[0], r[1], r[2]) if need_dist else (r[0], r[1], None)
        if kk > 0:
            r = find_matches_device(shard.spot_key, query, kk, index_offset=shard.index_offset,
                                    dist_p=p if need_dist else None)
            val[:, :kk], idx[:, :kk] = r[0], r[1]
            if need_dist:
                dst[:, :kk] = r[2]
        return val, idx, dst

    def merge(self, vals, idx, dst, k: int):
        R, Q, _ = vals.shape
        dev = vals.device
        ov = torch.empty((Q, k), dtype=torch.float32, device=dev)
        oi = torch.empty((Q, k), dtype=torch.int64, device=dev)
        od = torch.empty((Q, k), dtype=torch.float32, device=dev) if dst is not None else None
        with torch.cuda.device(dev):
            check(load().mclst_merge_topk(ptr(vals), ptr(idx), ptr(dst), R, Q, k, ptr(ov), ptr(oi), ptr(od),
                                          stream_ptr()), "merge_topk")
        return ov, oi, od

    def weights(self, dst, val, mode: str):
        Q, k = val.shape
        w = torch.empty((Q, k), dtype=torch.float32, device=val.device)
        with torch.cuda.device(val.device):
            check(load().mclst_neighbor_weights(ptr(dst), ptr(val), Q, k, WEIGHT_MODES[mode], ptr(w),
                                                stream_ptr()), "neighbor_weights")
        return w

    def partial_average(self, rows: torch.Tensor, index_offset: int, idx, w):
        n_loc, G = rows.shape
        Q, k = idx.shape
        out = torch.empty((Q, G), dtype=torch.float32, device=idx.device)
        with torch.cuda.device(idx.device):
            check(load().mclst_weighted_gather(ptr(rows), n_loc, rows.stride(0), G,
                                               int(rows.dtype == torch.float64), ptr(idx), ptr(w), Q, k,
                                               index_offset, ptr(out), stream_ptr()), "weighted_gather")
        return out


def _all_gather_stack(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous(), group=group)
    return torch.stack(out)


def retrieve_sharded(shard: BankShard, query: torch.Tensor, top_k: int = 50, mode: str = "inv_sq_l2",
                     group=None, backend=None, want_emb: bool = False):
    """(indices int64 [Q,k], values f32 [Q,k], emb_pred | None, expr_pred f32 [Q,G]) -- identical on
    every rank and equal to the single-GPU ``retrieve_device`` on the concatenated bank."""
    backend = backend or CudaBackend()
    need_dist = mode in ("inv_sq_l1", "inv_sq_l2", "bleep_exp")
    p = 1 if mode == "inv_sq_l1" else 2
    val, idx, dst = backend.local_topk(shard, query, top_k, p, need_dist)
    world = dist.get_world_size(group) if (dist.is_initialized() and shard.n_total != shard.spot_key.shape[0]) else 1
    if world > 1:
        vals = _all_gather_stack(val, group)
        idxs = _all_gather_stack(idx, group)
        dsts = _all_gather_stack(dst, group) if need_dist else None
        val, idx, dst = backend.merge(vals, idxs, dsts, top_k)
    w = backend.weights(dst, val, mode)
    if shard.expr_ready is not None:
        torch.cuda.current_stream(shard.expression_key.device).wait_event(shard.expr_ready)
    expr = backend.partial_average(shard.expression_key, shard.index_offset, idx, w)
    emb = backend.partial_average(shard.spot_key, shard.index_offset, idx, w) if want_emb else None
    if world > 1:
        dist.all_reduce(expr, group=group)
        if emb is not None:
            dist.all_reduce(emb, group=group)
    return idx, val, emb, expr


@dataclass
class RetrievalGrid:
    """2-D decomposition of a retrieval job over world = query_groups x bank_shards ranks.

    Ranks [g*bank_shards, (g+1)*bank_shards) form query group g: they share one slice of the
    queries and hold one bank shard each (candidate all-gather + merge inside the group).
    Sharding the BANK multiplies the streaming top-k / re-rank work (every shard produces its
    own k + band candidates for every query); sharding the QUERIES does not and needs no
    exchange, but replicates the bank.  bank_shards is therefore a capacity knob: 1 when the
    bank fits on a GPU, larger when it does not."""
    query_groups: int
    bank_shards: int
    q_index: int
    b_index: int
    group: object          # process group of this rank's query group (None when bank_shards == 1)

    def query_slice(self, n_query: int) -> Tuple[int, int]:
        return shard_bounds(n_query, self.query_groups)[self.q_index]


def make_retrieval_grid(bank_shards: int, world: Optional[int] = None, rank: Optional[int] = None) -> RetrievalGrid:
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    if world % bank_shards != 0:
        raise ValueError(f"bank_shards={bank_shards} does not divide world={world}")
    qg = world // bank_shards
    mine = None
    if bank_shards > 1:
        for g in range(qg):                     # every rank must take part in every new_group call
            h = dist.new_group(list(range(g * bank_shards, (g + 1) * bank_shards)))
            if g == rank // bank_shards:
                mine = h
    return RetrievalGrid(qg, bank_shards, rank // bank_shards, rank % bank_shards, mine)


# ----------------------------------------------------------------------------- loss
class _ShardedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_loc, i_loc, temperature, mode, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        rows, D = s_loc.shape
        B = rows * world
        if rows % 128 != 0:
            raise _lib.MclstError("contrastive_loss_sharded: rows per rank must be a multiple of 128")
        S = torch.empty((B, D), dtype=torch.float32, device=s_loc.device)
        I = torch.empty((B, D), dtype=torch.float32, device=s_loc.device)
        dist.all_gather_into_tensor(S, s_loc.detach().contiguous(), group=group)
        dist.all_gather_into_tensor(I, i_loc.detach().contiguous(), group=group)
        lib = load()
        nbytes = C.c_size_t()
        check(lib.mclst_contrastive_loss_workspace_bytes(B, D, mode, rows, C.byref(nbytes)), "loss workspace")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=S.device)
        stats = torch.zeros((6, B), dtype=torch.float32, device=S.device)
        loss = torch.zeros((), dtype=torch.float32, device=S.device)
        dS = torch.empty((rows, D), dtype=torch.float32, device=S.device)
        dI = torch.empty((rows, D), dtype=torch.float32, device=S.device)
        row0 = rank * rows

        def phase(ph):
            with torch.cuda.device(S.device):
                check(lib.mclst_contrastive_loss_phase(ptr(S), S.stride(0), ptr(I), I.stride(0), B, D,
                                                       float(temperature), mode, row0, rows, ph, ptr(stats),
                                                       ptr(loss), ptr(dS), dS.stride(0), ptr(dI),
                                                       dI.stride(0), ptr(ws), ws.numel(), stream_ptr()),
                      f"contrastive_loss_phase {ph}")

        def gather_rows(sel):
            loc = stats[sel, row0:row0 + rows].contiguous()
            # (output as the dim-0 concatenation of the inputs: the one shape both NCCL and gloo accept)
            full = torch.empty((world * len(sel), rows), dtype=torch.float32, device=S.device)
            dist.all_gather_into_tensor(full, loc, group=group)
            stats[sel] = full.view(world, len(sel), rows).permute(1, 0, 2).reshape(len(sel), B)

        phase(1)
        gather_rows([0, 1, 2, 5])
        if mode != _lib.T_EYE:
            phase(2)
            gather_rows([3, 4])
        phase(3)
        dist.all_reduce(loss, group=group)
        ctx.save_for_backward(dS, dI)
        return loss

    @staticmethod
    def backward(ctx, g):
        dS, dI = ctx.saved_tensors
        return g * dS, g * dI, None, None, None


def contrastive_loss_sharded(spot_emb_local: torch.Tensor, image_emb_local: torch.Tensor,
                             temperature: float = 1.0, targets: str = "eye", soft_scale: str = "div",
                             group=None) -> torch.Tensor:
    """Global-batch contrastive loss over all ranks' rows (each rank passes its [B/R, D] slice).
    Returns the full-batch loss (same value on every rank); ``backward`` yields the gradients
    of that loss w.r.t. the local slices."""
    from .loss import TARGET_MODES
    return _ShardedLoss.apply(spot_emb_local, image_emb_local, float(temperature),
                              TARGET_MODES[(targets, soft_scale)], group)
