"""Parity of the CUDA retrieval path (through the C-ABI) with the oracle and with the
golden fixtures made from the reference's own code.  Needs a B200."""
import numpy as np
import pytest
import torch

from mclstexp_b200 import retrieval, synth
from oracle import oracle
from test_oracle_golden import _retrieval_inputs, LOOPS

pytestmark = pytest.mark.gpu

RTOL = 1e-3      # north star: predicted expression within 1e-3 relative error under FP32


def _check_spec(bank, qry, k, exact_only):
    val, idx = retrieval.find_matches(bank, qry, k, return_values=True, exact_only=exact_only)
    sval, sidx = oracle.find_matches_spec(bank, qry, k)
    if idx.ndim == 1:
        idx, val = idx[None], val[None]
    assert idx.dtype == np.int64 and val.dtype == np.float32
    np.testing.assert_array_equal(idx, sidx)           # bit-exact indices, ties -> lowest index
    np.testing.assert_array_equal(val, sval)           # bit-exact float32 similarities


@pytest.mark.parametrize("exact_only", [True, False])
@pytest.mark.parametrize("name", ["iid", "clustered", "pm1"])
@pytest.mark.parametrize("k", [1, 50, 200])
def test_find_matches_bit_exact_vs_spec(golden_retrieval, name, k, exact_only):
    z, meta = golden_retrieval
    bank, qry, _ = _retrieval_inputs(meta, name)
    _check_spec(bank, qry, k, exact_only)


@pytest.mark.parametrize("exact_only", [True, False])
@pytest.mark.parametrize("name", ["iid", "clustered"])
@pytest.mark.parametrize("k", [1, 50, 200])
def test_find_matches_vs_reference_golden(golden_retrieval, name, k, exact_only):
    """Against what the reference's own find_matches returned (evel_her2st.py:74-84)."""
    z, meta = golden_retrieval
    bank, qry, _ = _retrieval_inputs(meta, name)
    idx = retrieval.find_matches(bank, qry, k, exact_only=exact_only)
    ref = z[f"{name}/find_matches/k{k}/indices"]
    ok = oracle.decidable_rows(bank, qry, k)
    np.testing.assert_array_equal(idx[ok], ref[ok])
    val, _ = retrieval.find_matches(bank, qry, k, return_values=True, exact_only=exact_only)
    np.testing.assert_allclose(val, z[f"{name}/find_matches/k{k}/values"], rtol=0, atol=5e-7)


def test_pm1_known_answer_against_reference_values(golden_retrieval):
    z, meta = golden_retrieval
    bank, qry, _ = _retrieval_inputs(meta, "pm1")
    val, idx = retrieval.find_matches(bank, qry, 50, return_values=True)
    np.testing.assert_array_equal(val, z["pm1/find_matches/k50/values"])   # exact arithmetic


def test_q1_squeeze_quirk(golden_retrieval):
    z, meta = golden_retrieval
    bank, qry, _ = _retrieval_inputs(meta, "iid")
    idx = retrieval.find_matches(bank, qry[:1], 5)
    assert idx.shape == (5,)
    np.testing.assert_array_equal(idx, z["iid/find_matches/q1/indices"])


@pytest.mark.parametrize("exact_only", [True, False])
@pytest.mark.parametrize("N,Q,D,k", [(50, 3, 256, 50), (257, 130, 256, 7), (1000, 5, 100, 33),
                                     (513, 129, 64, 64), (3000, 300, 256, 600), (40, 1, 8, 40)])
def test_find_matches_ragged_shapes(N, Q, D, k, exact_only):
    bank = synth.embeddings(N, D, 900 + N)
    qry = synth.embeddings(Q, D, 901 + N)
    _check_spec(bank, qry, k, exact_only)


@pytest.mark.parametrize("exact_only", [True, False])
def test_find_matches_degenerate_rows(exact_only):
    bank = synth.embeddings(300, 256, 5)
    bank[7] = 0.0                       # zero-norm row: F.normalize eps keeps it at 0
    bank[11] = bank[3]                  # exact duplicate -> exact tie, lowest index first
    qry = synth.embeddings(9, 256, 6)
    qry[2] = 0.0                        # zero query: every similarity is 0 -> indices 0..k-1
    _check_spec(bank, qry, 20, exact_only)
    idx = retrieval.find_matches(bank, qry, 20, exact_only=exact_only)
    np.testing.assert_array_equal(idx[2], np.arange(20))


def test_find_matches_errors():
    bank = synth.embeddings(10, 16, 1)
    with pytest.raises(RuntimeError):
        retrieval.find_matches(bank, bank[:2], 11)       # k > N, torch.topk raises too


def test_empty_queries():
    bank = torch.tensor(synth.embeddings(10, 16, 1)).cuda()
    val, idx = retrieval.find_matches_device(bank, bank[:0], 3)
    assert idx.shape == (0, 3)


@pytest.mark.parametrize("name", ["iid", "clustered"])
@pytest.mark.parametrize("tag,mode", LOOPS)
def test_weighted_average_vs_reference_golden(golden_retrieval, name, tag, mode):
    z, meta = golden_retrieval
    bank, qry, expr = _retrieval_inputs(meta, name)
    idx = z[f"{name}/{tag}/indices"].astype(np.int64)
    emb, ex = retrieval.weighted_topk_average(bank, expr, qry, idx, mode=mode)
    assert emb.dtype == np.float64 and ex.dtype == np.float64
    # bleep_exp exponentiates float32 squared distances ~5e2: the reference's own float32
    # round-off is 1e-4-level there, so elements near zero get an absolute tolerance
    a = 2e-4 if mode == "bleep_exp" else 1e-5
    np.testing.assert_allclose(emb, z[f"{name}/{tag}/emb_pred"], rtol=RTOL, atol=a)
    np.testing.assert_allclose(ex, z[f"{name}/{tag}/expr_pred"], rtol=RTOL, atol=a * 0.1)


@pytest.mark.parametrize("mode", ["inv_sq_l1", "inv_sq_l2", "similarity", "uniform", "bleep_exp"])
@pytest.mark.parametrize("G,dtype", [(785, np.float32), (1000, np.float32), (171, np.float64), (4, np.float32)])
def test_weighted_average_modes_and_layouts(mode, G, dtype):
    N, Q, D, k = 600, 21, 256, 50
    bank = synth.embeddings(N, D, 41, "clustered")
    qry = synth.embeddings(Q, D, 42, "clustered")
    expr = synth.expression(N, G, 43, dtype=dtype)
    val, idx = oracle.find_matches_spec(bank, qry, k)
    emb, ex = retrieval.weighted_topk_average(bank, expr, qry, idx, mode=mode, values=val)
    emb64, ex64 = oracle.weighted_average_spec(bank, expr, qry, idx, mode, val)
    a = 2e-4 if mode == "bleep_exp" else 1e-5
    np.testing.assert_allclose(emb, emb64, rtol=RTOL, atol=a)
    np.testing.assert_allclose(ex, ex64, rtol=RTOL, atol=a * 0.1)


def test_weighted_average_zero_distance_defined():
    bank = synth.embeddings(64, 32, 1)
    qry = bank[[3, 7]].copy()
    expr = synth.expression(64, 12, 2)
    idx = np.array([[3, 5, 9], [1, 7, 2]], np.int64)
    for mode in ("inv_sq_l1", "inv_sq_l2"):
        _, ex = retrieval.weighted_topk_average(bank, expr, qry, idx, mode=mode)
        np.testing.assert_allclose(ex[0], expr[3], rtol=1e-6)
        np.testing.assert_allclose(ex[1], expr[7], rtol=1e-6)


@pytest.mark.parametrize("name,tag", [("iid", "loop_her2st"), ("clustered", "loop_visium"),
                                      ("iid", "loop_cscc")])
def test_retrieve_fold_body_vs_reference_golden(golden_retrieval, name, tag):
    """The whole fold-loop body (find_matches + loop) against the reference's verbatim run."""
    z, meta = golden_retrieval
    bank, qry, expr = _retrieval_inputs(meta, name)
    info = meta[name][tag]
    idx, emb, ex = retrieval.retrieve(bank, expr, qry, top_k=info["k"], mode=info["mode"])
    ok = oracle.decidable_rows(bank, qry, info["k"])
    ref_idx = z[f"{name}/{tag}/indices"]
    np.testing.assert_array_equal(idx[ok], ref_idx[ok])
    np.testing.assert_allclose(ex, z[f"{name}/{tag}/expr_pred"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(emb, z[f"{name}/{tag}/emb_pred"], rtol=RTOL, atol=1e-5)


def test_cfg3_shape_properties():
    """BASELINE cfg3 size (N=30000, Q=4000): size-independent properties + sampled rows."""
    c = synth.CONFIGS["cfg3"]
    bank = synth.embeddings(c["N"], c["D"], 1237, "clustered")
    qry = synth.embeddings(c["Q"], c["D"], 1238, "clustered")
    expr = synth.expression(c["N"], c["G"], 1239)
    tb, tq, te = (torch.tensor(a).cuda() for a in (bank, qry, expr))
    idx, val, _, ex = retrieval.retrieve_device(tb, te, tq, c["k"], "inv_sq_l2")
    idx_c, val_c, ex_c = idx.cpu().numpy(), val.cpu().numpy(), ex.cpu().numpy()
    assert (np.diff(val_c, axis=1) <= 0).all()                       # sorted descending
    assert all(len(set(r)) == c["k"] for r in idx_c[::97])            # no duplicates
    assert idx_c.min() >= 0 and idx_c.max() < c["N"]
    rows = np.arange(0, c["Q"], 131)
    sval, sidx = oracle.find_matches_spec(bank, qry[rows], c["k"])
    np.testing.assert_array_equal(idx_c[rows], sidx)
    np.testing.assert_array_equal(val_c[rows], sval)
    _, ex64 = oracle.weighted_average_spec(bank, expr, qry[rows], sidx, "inv_sq_l2")
    np.testing.assert_allclose(ex_c[rows], ex64, rtol=RTOL, atol=1e-6)
    # convexity: a weighted average of rows lies inside their per-gene range
    g = expr[idx_c[rows]]
    assert (ex_c[rows] <= g.max(1) + 1e-5).all() and (ex_c[rows] >= g.min(1) - 1e-5).all()
    # the exact path and the default path agree everywhere
    val_e, idx_e = retrieval.find_matches_device(tb, tq, c["k"], exact_only=True)
    assert torch.equal(idx_e, idx) and torch.equal(val_e, val)


def test_resident_bank_and_pinned_io_equal_one_shot():
    """retrieve() with pageable NumPy inputs, with pinned tensors, and Bank.retrieve() with the
    bank resident give byte-identical results (the upload overlap and the pinned result buffers
    of the host layer change nothing numerically); results large enough to take the pinned path."""
    N, Q, D, G, k = 6000, 1500, 256, 300, 50
    bank = synth.embeddings(N, D, 4101, "clustered")
    qry = synth.embeddings(Q, D, 4102, "clustered")
    expr = synth.expression(N, G, 4103)
    a = retrieval.retrieve(bank, expr, qry, top_k=k, p=2)
    pb, pe, pq = (torch.from_numpy(x).pin_memory() for x in (bank, expr, qry))
    b = retrieval.retrieve(pb, pe, pq, top_k=k, p=2)
    resident = retrieval.Bank(bank, expr)
    assert len(resident) == N
    c = resident.retrieve(qry, top_k=k, p=2)
    c2 = resident.retrieve(pq, top_k=k, p=2)                     # second call reuses the pinned pool
    for other in (b, c, c2):
        for x, y in zip(a, other):
            assert x.dtype == y.dtype and x.shape == y.shape
            np.testing.assert_array_equal(x, y)
    assert a[0].dtype == np.int64 and a[1].dtype == np.float64 and a[2].dtype == np.float64
    sval, sidx = oracle.find_matches_spec(bank, qry[::50], k)
    np.testing.assert_array_equal(a[0][::50], sidx)


def test_bank_shard_from_host_single_rank():
    """BankShard.from_host (expression rows uploaded on a side stream) + retrieve_sharded without a
    process group equals retrieve_device on the same data."""
    from mclstexp_b200.distributed import BankShard, retrieve_sharded
    N, Q, D, G, k = 9000, 700, 256, 200, 50
    bank = synth.embeddings(N, D, 5101, "clustered")
    qry = synth.embeddings(Q, D, 5102, "clustered")
    expr = synth.expression(N, G, 5103)
    dev = torch.device("cuda", 0)
    hb, he = torch.from_numpy(bank).pin_memory(), torch.from_numpy(expr).pin_memory()
    tq = torch.from_numpy(qry).to(dev)
    shard = BankShard.from_host(hb, he, 0, N, dev)
    idx, val, emb, ex = retrieve_sharded(shard, tq, k, "inv_sq_l2", want_emb=True)
    idx1, val1, emb1, ex1 = retrieval.retrieve_device(hb.to(dev), he.to(dev), tq, k, "inv_sq_l2", want_emb=True)
    assert torch.equal(idx, idx1) and torch.equal(val, val1)
    torch.testing.assert_close(ex, ex1, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(emb, emb1, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N,Q", [(70000, 19000), (70000, 16384), (5000, 40000), (300, 19100)])
def test_lane_rounds_bit_exact(N, Q):
    """Shapes whose query-block count exercises every branch of the persistent kernel's unit lists
    (full rounds + main/tail last round; many rounds over a short bank; tiny bank): default path
    == exact path == spec on sampled rows."""
    k = 50
    bank = synth.embeddings(N, 256, 6001 + N, "clustered")
    qry = synth.embeddings(Q, 256, 6002 + Q, "clustered")
    tb, tq = torch.from_numpy(bank).cuda(), torch.from_numpy(qry).cuda()
    val, idx = retrieval.find_matches_device(tb, tq, k)
    val_e, idx_e = retrieval.find_matches_device(tb, tq, k, exact_only=True)
    assert torch.equal(idx, idx_e) and torch.equal(val, val_e)
    rows = np.arange(0, Q, 997)
    sval, sidx = oracle.find_matches_spec(bank, qry[rows], k)
    np.testing.assert_array_equal(idx.cpu().numpy()[rows], sidx)


# ------------------------------------------------------------------ staged find_matches / resident bank
def _merge(vals, idxs, k):
    import ctypes as C
    from mclstexp_b200._lib import check, load, ptr, stream_ptr
    R, Q, _ = vals.shape
    ov = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    oi = torch.empty((Q, k), dtype=torch.int64, device="cuda")
    check(load().mclst_merge_topk(ptr(vals), ptr(idxs), None, R, Q, k, ptr(ov), ptr(oi), None, stream_ptr()),
          "merge_topk")
    return ov, oi


@pytest.mark.parametrize("N,Q,k,shards", [(40000, 700, 50, 3), (70000, 9000, 50, 8), (9000, 300, 200, 2),
                                          (5000, 130, 50, 4)])
def test_bank_shards_with_bound_exchange_equal_whole_bank(N, Q, k, shards):
    """The multi-GPU bank-shard protocol replayed on ONE GPU: every shard's seed pass offers its
    bounds, they are combined exactly as distributed.retrieve_sharded combines them (max of the k-th
    best bounds, min of the ceil(k/R)-th best bounds), every shard's candidate pass filters against
    the global bound and reports what its converged thresholds imply, the maximum of those decides
    what each shard still re-ranks (lists come back short and padded), and the merged result must be
    bit-identical to find_matches on the whole bank -- and to the oracle's spec."""
    bank = torch.tensor(synth.embeddings(N, 256, 61, "clustered"), device="cuda")
    qry = torch.tensor(synth.embeddings(Q, 256, 62, "clustered"), device="cuda")
    wval, widx = retrieval.find_matches_device(bank, qry, k)
    cuts = [N * s // shards for s in range(shards + 1)]
    parts = [bank[cuts[s]:cuts[s + 1]].contiguous() for s in range(shards)]
    k_part = -(-k // shards)
    wss, bounds = [], []
    for pb in parts:
        ws = retrieval.fm_workspace(pb.shape[0], Q, 256, k, pb.device)
        wss.append(ws)
        bounds.append(retrieval.fm_seed(pb, qry, k, ws, k_part, want_bounds=True))
    b = torch.stack(bounds)                                     # [R, 2, Q]
    ext = torch.maximum(b[:, 0].max(0).values, b[:, 1].min(0).values).contiguous()
    # second exchange: the converged thresholds of every shard's candidate pass
    b2 = torch.stack([retrieval.fm_candidates(pb, qry, k, wss[s], ext, bank_packed=True, k_part=k_part)
                      for s, pb in enumerate(parts)])                                # [R, 2, Q]
    ext2 = torch.maximum(torch.maximum(b2[:, 0].max(0).values, b2[:, 1].min(0).values), ext).contiguous()
    assert bool((ext2 <= wval[:, -1] + 1e-6).all())             # a LOWER bound of the true k-th best
    vals, idxs, short = [], [], 0
    for s, pb in enumerate(parts):
        v, i, _ = retrieval.fm_main(pb, qry, k, wss[s], index_offset=cuts[s], ext_bound=ext2, bank_packed=True,
                                    finish_only=True)
        short += int((i == 0x7fffffff).any(1).sum())
        vals.append(v)
        idxs.append(i)
    mv, mi = _merge(torch.stack(vals), torch.stack(idxs), k)
    assert torch.equal(mi, widx) and torch.equal(mv, wval)
    if N >= 40000:
        assert short > 0          # the bound really lets shards drop rows that cannot win
    sval, sidx = oracle.find_matches_spec(bank.cpu().numpy(), qry[:64].cpu().numpy(), k)
    np.testing.assert_array_equal(mi[:64].cpu().numpy(), sidx)
    np.testing.assert_array_equal(mv[:64].cpu().numpy(), sval)


def test_staged_without_bound_equals_single_call_and_extreme_bounds():
    bank = torch.tensor(synth.embeddings(20000, 256, 71, "clustered"), device="cuda")
    qry = torch.tensor(synth.embeddings(500, 256, 72, "clustered"), device="cuda")
    k = 50
    wval, widx, wdst = retrieval.find_matches_device(bank, qry, k, dist_p=2)
    ws = retrieval.fm_workspace(20000, 500, 256, k, bank.device)
    retrieval.fm_seed(bank, qry, k, ws)
    v, i, d = retrieval.fm_main(bank, qry, k, ws, dist_p=2)
    assert torch.equal(i, widx) and torch.equal(v, wval) and torch.equal(d, wdst)
    # a bound just below the true k-th best changes nothing
    retrieval.fm_seed(bank, qry, k, ws, bank_packed=True)
    v, i, d = retrieval.fm_main(bank, qry, k, ws, dist_p=2, ext_bound=(wval[:, -1] - 1e-6).contiguous(),
                                bank_packed=True)
    assert torch.equal(i, widx) and torch.equal(v, wval) and torch.equal(d, wdst)
    # a bound above every score: nothing of this bank can be among the global winners
    retrieval.fm_seed(bank, qry, k, ws, bank_packed=True)
    v, i, d = retrieval.fm_main(bank, qry, k, ws, dist_p=2,
                                ext_bound=torch.full((500,), 2.0, device="cuda"), bank_packed=True)
    assert bool((i == 0x7fffffff).all()) and bool(torch.isinf(v).all()) and bool((v < 0).all())
    assert bool(torch.isinf(d).all())
    # -inf / NaN bounds mean "nothing known"
    bad = torch.full((500,), float("-inf"), device="cuda")
    bad[::2] = float("nan")
    retrieval.fm_seed(bank, qry, k, ws, bank_packed=True)
    v, i, _ = retrieval.fm_main(bank, qry, k, ws, ext_bound=bad, bank_packed=True)
    assert torch.equal(i, widx) and torch.equal(v, wval)


def test_resident_bank_equals_one_shot_retrieve():
    """retrieval.Bank keeps the packed image of the bank across calls (query batches of different
    sizes, capacity growth, two top_k classes) and must return what the one-shot path returns."""
    N, G = 30000, 300
    bank = synth.embeddings(N, 256, 81, "clustered")
    expr = synth.expression(N, G, 82)
    b = retrieval.Bank(bank, expr)
    for Q, k, p in ((100, 50, 2), (3000, 50, 1), (700, 50, 2), (64, 200, 2)):
        qry = synth.embeddings(Q, 256, 90 + Q, "clustered")
        idx, emb, ex = b.retrieve(qry, top_k=k, p=p)
        idx1, emb1, ex1 = retrieval.retrieve(bank, expr, qry, top_k=k, p=p)
        np.testing.assert_array_equal(idx, idx1)
        np.testing.assert_array_equal(ex, ex1)
        np.testing.assert_array_equal(emb, emb1)
    assert set(b._packed) == {50, 200}


@pytest.mark.parametrize("mode,want_emb", [("inv_sq_l2", True), ("inv_sq_l1", False), ("similarity", False)])
def test_blocked_retrieve_equals_unblocked(mode, want_emb):
    """mclst_retrieve sends query sets beyond eight lane rounds (> 8 x 128 x SMs rows) through in
    blocks of four, the average of one block on a side stream next to the top-k pass of the next: results must
    be byte-identical to the same queries retrieved in unblocked pieces, the path counters must add
    up over the blocks, and a second call on the same stream must not disturb the first."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    N, D, G, k = 2000, 256, 24, 50
    Q = 8 * 128 * sms + 777                                      # blocks: four rounds, four rounds, 777 rows
    bank = torch.tensor(synth.embeddings(N, D, 5101, "clustered")).cuda()
    expr = torch.tensor(synth.expression(N, G, 5102)).cuda()
    qry = torch.tensor(synth.embeddings(Q, D, 5103, "clustered")).cuda()
    idx, val, emb, ex = retrieval.retrieve_device(bank, expr, qry, k, mode, want_emb=want_emb)
    counters = retrieval.last_counters()
    assert counters["tensor_core"] + counters["exact_fallback"] == Q, counters
    again = retrieval.retrieve_device(bank, expr, qry, k, mode, want_emb=want_emb)   # right behind it
    torch.cuda.synchronize()
    cut = Q // 2 + 5                                             # both pieces below the blocking threshold
    parts = [retrieval.retrieve_device(bank, expr, qry[a:b], k, mode, want_emb=want_emb)
             for a, b in ((0, cut), (cut, Q))]
    for i, got in enumerate((idx, val, emb, ex)):
        if got is None:
            continue
        want = torch.cat([p[i] for p in parts])
        assert torch.equal(got, want), i
        assert torch.equal(again[i], want), i
    rows = np.arange(0, Q, 997)
    sval, sidx = oracle.find_matches_spec(bank.cpu().numpy(), qry.cpu().numpy()[rows], k)
    np.testing.assert_array_equal(idx.cpu().numpy()[rows], sidx)
    # resident bank: same blocks against the packed image kept in the workspace
    res = retrieval.Bank(bank, expr)
    r_idx, r_val, _, r_ex = res.retrieve_device(qry, k, mode)
    assert torch.equal(r_idx, idx) and torch.equal(r_val, val) and torch.equal(r_ex, ex)
