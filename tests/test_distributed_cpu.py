"""World-size-2 gloo test of the sharded-retrieval host logic (sharding, candidate all-gather,
merge, owner-only partial sums, all-reduce) with an oracle-backed backend injected in place of
the CUDA kernels.  The kernels themselves are covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Same interface as mclstexp_b200.distributed.CudaBackend, numpy/oracle inside."""

    def local_topk(self, shard, query, k, p, need_dist):
        from oracle import oracle
        bank = shard.spot_key.numpy()
        q = query.numpy()
        kk = min(k, bank.shape[0])
        val = np.full((q.shape[0], k), -np.inf, np.float32)
        idx = np.full((q.shape[0], k), 2 ** 31 - 1, np.int64)
        dst = np.full((q.shape[0], k), np.inf, np.float32)
        v, i = oracle.find_matches_spec(bank, q, kk)
        val[:, :kk], idx[:, :kk] = v, i + shard.index_offset
        diff = bank[i].astype(np.float64) - q[:, None, :].astype(np.float64)
        d = np.abs(diff).sum(-1) if p == 1 else np.sqrt((diff ** 2).sum(-1))
        dst[:, :kk] = d.astype(np.float32)
        return torch.from_numpy(val), torch.from_numpy(idx), (torch.from_numpy(dst) if need_dist else None)

    def merge(self, vals, idx, dst, k):
        R, Q, kk = vals.shape
        v = vals.permute(1, 0, 2).reshape(Q, R * kk).numpy()
        i = idx.permute(1, 0, 2).reshape(Q, R * kk).numpy()
        order = np.lexsort((i, -v), axis=1)[:, :k]                 # value desc, index asc
        take = lambda a: torch.from_numpy(np.take_along_axis(a, order, axis=1).copy())
        d = None if dst is None else take(dst.permute(1, 0, 2).reshape(Q, R * kk).numpy())
        return take(v), take(i), d

    def weights(self, dst, val, mode):
        if mode in ("inv_sq_l1", "inv_sq_l2"):
            d = dst.numpy().astype(np.float64)
            zero = d == 0
            with np.errstate(divide="ignore"):
                w = 1.0 / d ** 2
            anyz = zero.any(1)
            w[anyz] = zero[anyz]
        elif mode == "similarity":
            w = val.numpy().astype(np.float64)
        elif mode == "uniform":
            w = np.ones(val.shape)
        else:
            d2 = dst.numpy().astype(np.float64) ** 2
            w = np.exp(-(d2 - d2[:, :1] + 1.0))
        return torch.from_numpy((w / w.sum(1, keepdims=True)).astype(np.float32))

    def partial_average(self, rows, index_offset, idx, w):
        r = rows.numpy().astype(np.float64)
        loc = idx.numpy() - index_offset
        own = (loc >= 0) & (loc < r.shape[0])
        g = r[np.where(own, loc, 0)] * (w.numpy().astype(np.float64) * own)[:, :, None]
        return torch.from_numpy(g.sum(1).astype(np.float32))


def _worker(rank, world, port, mode, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mclstexp_b200 import synth
    from mclstexp_b200.distributed import BankShard, retrieve_sharded, shard_bounds
    N, Q, D, G = 515, 23, 64, 12
    bank = torch.from_numpy(synth.embeddings(N, D, 5, "clustered"))
    bank[300] = bank[100]                                    # exact tie across the shard boundary
    expr = torch.from_numpy(synth.expression(N, G, 6))
    qry = torch.from_numpy(synth.embeddings(Q, D, 7, "clustered"))
    assert shard_bounds(N, world)[0] == (0, N // world) and shard_bounds(N, world)[-1][1] == N
    shard = BankShard.from_full(bank, expr, rank, world)
    idx, val, emb, ex = retrieve_sharded(shard, qry, k, mode, backend=OracleBackend(), want_emb=True)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), idx=idx.numpy(), val=val.numpy(), emb=emb.numpy(),
             ex=ex.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,k", [("inv_sq_l2", 50), ("inv_sq_l1", 7), ("similarity", 300), ("uniform", 1)])
def test_sharded_retrieval_host_logic_world2(tmp_path, mode, k):
    from mclstexp_b200 import synth
    from oracle import oracle
    port = 29500 + (os.getpid() + hash(mode)) % 2000
    mp.spawn(_worker, args=(2, port, mode, k, str(tmp_path)), nprocs=2, join=True)
    N, Q, D, G = 515, 23, 64, 12
    bank = synth.embeddings(N, D, 5, "clustered")
    bank[300] = bank[100]
    expr = synth.expression(N, G, 6)
    qry = synth.embeddings(Q, D, 7, "clustered")
    sval, sidx = oracle.find_matches_spec(bank, qry, k)
    emb64, ex64 = oracle.weighted_average_spec(bank, expr, qry, sidx, mode, sval)
    r0, r1 = (np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(2))
    for r in (r0, r1):
        np.testing.assert_array_equal(r["idx"], sidx)          # index-exact vs the single-process result
        np.testing.assert_array_equal(r["val"], sval)
        np.testing.assert_allclose(r["ex"], ex64, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r["emb"], emb64, rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(r0["ex"], r1["ex"])          # every rank holds the same answer


def test_shard_bounds_cover_everything():
    from mclstexp_b200.distributed import shard_bounds
    for n in (0, 1, 7, 1000, 1_000_000):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


# --------------------------------------------------------------------------------------------
# Row-sharded contrastive loss: the exchange pattern of distributed._ShardedLoss (embedding
# all-gather, three phases separated by all-gathers of the per-row statistics, loss all-reduce)
# with the three CUDA phases replaced by a NumPy restatement working on the same raw buffers.
class _PhaseOracleLib:
    """Stands in for libmclst_b200.so: mclst_contrastive_loss_phase on host pointers, float64 inside.
    stats rows: 0 rl, 1 cl, 2 za, 3 wbar, 4 cs, 5 diag, 6-8 lo halves of rl / cl / za (csrc/loss.cu)."""

    @staticmethod
    def _view(p, n):
        import ctypes as C
        addr = p.value if hasattr(p, "value") else int(p)
        return np.ctypeslib.as_array((C.c_float * n).from_address(addr))

    def mclst_contrastive_loss_workspace_bytes(self, B, D, mode, rows, out):
        out._obj.value = 256
        return 0

    def mclst_contrastive_loss_phase(self, s, ld_s, i, ld_i, B, D, T, mode, row0, rows, phase, stats, loss,
                                     d_s, ld_ds, d_i, ld_di, ws, ws_bytes, stream):
        S = self._view(s, B * ld_s).reshape(B, ld_s)[:, :D].astype(np.float64)
        I = self._view(i, B * ld_i).reshape(B, ld_i)[:, :D].astype(np.float64)
        st = self._view(stats, 9 * B).reshape(9, B)
        soft = mode != 0
        a_scale = 0.0 if not soft else (1.0 / (2 * T) if mode == 1 else T / 2.0)
        loc = slice(row0, row0 + rows)

        def lse(x):
            m = x.max(1, keepdims=True)
            return (m + np.log(np.exp(x - m).sum(1, keepdims=True)))[:, 0]

        Lg_loc = S[loc] @ I.T / T                      # Lg[r, j], r local
        LgT_loc = I[loc] @ S.T / T                     # Lg[j, r] as [r, j]
        A_loc = (I[loc] @ I.T + S[loc] @ S.T) * a_scale if soft else None
        if phase == 1:
            def put(k, x):                             # float pair: hi in row k, lo in row 6 + k
                st[k, loc] = x.astype(np.float32)
                st[6 + k, loc] = (x - st[k, loc].astype(np.float64)).astype(np.float32)
            put(0, lse(Lg_loc))
            put(1, lse(LgT_loc))
            st[5, loc] = Lg_loc[np.arange(rows), np.arange(row0, row0 + rows)]
            if soft:
                put(2, lse(A_loc))
            return 0
        rl, cl, za = (st[k].astype(np.float64) + st[6 + k].astype(np.float64) for k in range(3))
        wbar, cs = st[3].astype(np.float64), st[4].astype(np.float64)
        if phase == 2:
            Pt = np.exp(A_loc - za[loc, None])
            st[3, loc] = (Pt * (rl[loc, None] + cl[None, :] - 2 * Lg_loc)).sum(1)
            st[4, loc] = np.exp(A_loc - za[None, :]).sum(1)
            return 0
        term = wbar[loc] if soft else rl[loc] + cl[loc] - 2 * st[5, loc].astype(np.float64)
        self._view(loss, 1)[0] = term.sum() / (2 * B)
        if d_s is None or (hasattr(d_s, "value") and not d_s.value):
            return 0
        if soft:
            Pt_rj = np.exp(A_loc - za[loc, None])
            Pt_jr = np.exp(A_loc - za[None, :])
            cs_j, cs_r = cs[None, :], cs[loc, None]
        else:
            Pt_rj = Pt_jr = (np.arange(row0, row0 + rows)[:, None] == np.arange(B)[None, :]).astype(np.float64)
            cs_j, cs_r = 1.0, 1.0
        dLg_rj = (np.exp(Lg_loc - rl[loc, None]) + cs_j * np.exp(Lg_loc - cl[None, :]) - 2 * Pt_rj) / (2 * B)
        dLg_jr = (np.exp(LgT_loc - rl[None, :]) + cs_r * np.exp(LgT_loc - cl[loc, None]) - 2 * Pt_jr) / (2 * B)
        dS = dLg_rj @ I / T
        dI = dLg_jr @ S / T
        if soft:
            W_rj = rl[loc, None] + cl[None, :] - 2 * Lg_loc
            W_jr = rl[None, :] + cl[loc, None] - 2 * LgT_loc
            sym = (Pt_rj * (W_rj - wbar[loc, None]) + Pt_jr * (W_jr - wbar[None, :])) / (2 * B)
            dS += sym @ S * a_scale
            dI += sym @ I * a_scale
        self._view(d_s, rows * ld_ds).reshape(rows, ld_ds)[:, :D] = dS
        self._view(d_i, rows * ld_di).reshape(rows, ld_di)[:, :D] = dI
        return 0


def _loss_worker(rank, world, port, targets, out_dir):
    import contextlib
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mclstexp_b200.distributed as md
    from mclstexp_b200 import synth
    md.load = lambda: _PhaseOracleLib()                     # no CUDA in this process
    md.stream_ptr = lambda: None
    torch.cuda.device = lambda dev: contextlib.nullcontext()
    B, D = 128 * world, 32
    S = torch.from_numpy(synth.embeddings(B, D, 21, "clustered", centres=5) * 0.3)
    I = torch.from_numpy(synth.embeddings(B, D, 22, "clustered", centres=5) * 0.3)
    rows = B // world
    Sl = S[rank * rows:(rank + 1) * rows].clone().requires_grad_(True)
    Il = I[rank * rows:(rank + 1) * rows].clone().requires_grad_(True)
    loss = md.contrastive_loss_sharded(Sl, Il, 0.7, targets)
    (2.0 * loss).backward()                                 # upstream gradient other than 1
    np.savez(os.path.join(out_dir, f"l{rank}.npz"), loss=loss.detach().numpy(), dS=Sl.grad.numpy(),
             dI=Il.grad.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("targets", ["eye", "soft"])
def test_sharded_loss_exchange_pattern_world2(tmp_path, targets):
    from mclstexp_b200 import synth
    from oracle import oracle
    world = 2
    port = 31500 + (os.getpid() + len(targets)) % 2000
    mp.spawn(_loss_worker, args=(world, port, targets, str(tmp_path)), nprocs=world, join=True)
    B, D = 128 * world, 32
    S = synth.embeddings(B, D, 21, "clustered", centres=5) * 0.3
    I = synth.embeddings(B, D, 22, "clustered", centres=5) * 0.3
    loss, dS, dI = oracle.contrastive_loss_closed_form(S, I, 0.7, targets)
    rows = B // world
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"l{r}.npz"))
        assert abs(float(z["loss"]) - loss) <= 1e-5 * abs(loss)           # same value on every rank
        np.testing.assert_allclose(z["dS"], 2.0 * dS[r * rows:(r + 1) * rows], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(z["dI"], 2.0 * dI[r * rows:(r + 1) * rows], rtol=2e-4, atol=1e-7)


# --------------------------------------------------------------------------------------------
# 2-D decomposition (query groups x bank shards) with per-group process groups, world size 4.
def _grid_worker(rank, world, port, bank_shards, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mclstexp_b200 import synth
    from mclstexp_b200.distributed import BankShard, make_retrieval_grid, retrieve_sharded
    N, Q, D, G, k = 403, 37, 48, 9, 11
    bank = torch.from_numpy(synth.embeddings(N, D, 15, "clustered"))
    expr = torch.from_numpy(synth.expression(N, G, 16))
    qry = torch.from_numpy(synth.embeddings(Q, D, 17, "clustered"))
    grid = make_retrieval_grid(bank_shards, world, rank)
    assert grid.query_groups * grid.bank_shards == world
    assert (grid.q_index, grid.b_index) == (rank // bank_shards, rank % bank_shards)
    shard = BankShard.from_full(bank, expr, grid.b_index, grid.bank_shards)
    q0, q1 = grid.query_slice(Q)
    idx, val, _, ex = retrieve_sharded(shard, qry[q0:q1].contiguous(), k, "inv_sq_l2", group=grid.group,
                                       backend=OracleBackend())
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), idx=idx.numpy(), val=val.numpy(), ex=ex.numpy(),
             q0=q0, q1=q1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bank_shards", [1, 2, 4])
def test_retrieval_grid_world4(tmp_path, bank_shards):
    from mclstexp_b200 import synth
    from mclstexp_b200.distributed import make_retrieval_grid
    from oracle import oracle
    world = 4
    port = 33500 + (os.getpid() + 7 * bank_shards) % 2000
    mp.spawn(_grid_worker, args=(world, port, bank_shards, str(tmp_path)), nprocs=world, join=True)
    N, Q, D, G, k = 403, 37, 48, 9, 11
    bank = synth.embeddings(N, D, 15, "clustered")
    expr = synth.expression(N, G, 16)
    qry = synth.embeddings(Q, D, 17, "clustered")
    sval, sidx = oracle.find_matches_spec(bank, qry, k)
    _, ex64 = oracle.weighted_average_spec(bank, expr, qry, sidx, "inv_sq_l2", sval)
    covered = np.zeros(Q, dtype=np.int32)
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"g{r}.npz"))
        q0, q1 = int(z["q0"]), int(z["q1"])
        np.testing.assert_array_equal(z["idx"], sidx[q0:q1])
        np.testing.assert_array_equal(z["val"], sval[q0:q1])
        np.testing.assert_allclose(z["ex"], ex64[q0:q1], rtol=1e-4, atol=1e-6)
        if r % bank_shards == 0:
            covered[q0:q1] += 1
    assert (covered == 1).all()                    # the query groups tile the queries exactly once
    with pytest.raises(ValueError):
        make_retrieval_grid(3, 4, 0)
