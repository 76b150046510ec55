"""Bring-up and soundness checks of the tcgen05 candidate pass (sim_topk.cu)."""
import numpy as np
import pytest
import torch

from mclstexp_b200 import retrieval, synth
from oracle import oracle

pytestmark = pytest.mark.gpu

E_ACC = 2.0e-6      # the accumulation bound the kernel assumes (sim_topk.cu)


def _fp16_operands(x):
    # what pack_rows builds: float32(x * float32(1 / float64 norm)), rounded to fp16
    x = np.ascontiguousarray(x, np.float32)
    inv = (1.0 / oracle.norms_spec(x)).astype(np.float32)
    xn = (x * inv[:, None]).astype(np.float32)
    xh = xn.astype(np.float16)
    resid = np.sqrt(((xn.astype(np.float64) - xh.astype(np.float64)) ** 2).sum(1))
    return xn, xh, resid


@pytest.mark.parametrize("N,Q,D,flavour", [(700, 37, 256, "iid"), (256, 128, 256, "clustered"),
                                           (1025, 129, 256, "clustered"), (300, 5, 64, "iid"),
                                           (530, 140, 100, "iid"), (5000, 260, 256, "clustered")])
def test_raw_tensor_core_similarities(N, Q, D, flavour):
    bank = synth.embeddings(N, D, 17 + N, flavour)
    qry = synth.embeddings(Q, D, 18 + N, flavour)
    got = retrieval.debug_similarity(bank, qry).cpu().numpy()
    assert not np.isnan(got).any(), "some similarities were never written"
    bn, bh, rb = _fp16_operands(bank)
    qn, qh, rq = _fp16_operands(qry)
    want16 = qh.astype(np.float64) @ bh.astype(np.float64).T       # what the MMA should compute
    acc_err = np.abs(got - want16).max()
    assert acc_err < E_ACC / 4, f"tensor-core accumulation error {acc_err:.3e} vs bound {E_ACC:.1e}"
    # the a-priori bound the exactness argument rests on: |approx - exact| <= r_q + r_s + E_ACC
    exact = oracle.similarity_spec(bank, qry).astype(np.float64)      # the float64 cosine (ranking key)
    bound = 1.002 * (rq[:, None] + rb[None, :]) + E_ACC + 1e-6
    assert (np.abs(got - exact) <= bound).all()
    print(f"N={N} Q={Q} D={D}: max acc err {acc_err:.2e}, max |approx-exact| "
          f"{np.abs(got - exact).max():.2e}, typical bound {bound.mean():.2e}")


def test_pm1_raw_similarities_are_exact():
    bank = synth.pm1_embeddings(600, 256, 5)
    qry = synth.pm1_embeddings(40, 256, 6)
    got = retrieval.debug_similarity(bank, qry).cpu().numpy()
    want = (qry.astype(np.float64) / 16) @ (bank.astype(np.float64) / 16).T
    np.testing.assert_array_equal(got, want.astype(np.float32))


def test_tensor_core_path_is_the_one_that_runs():
    bank = synth.embeddings(20000, 256, 1, "clustered")
    qry = synth.embeddings(700, 256, 2, "clustered")
    val, idx = retrieval.find_matches(bank, qry, 50, return_values=True)
    c = retrieval.last_counters()
    assert c["tensor_core"] + c["exact_fallback"] == 700
    assert c["tensor_core"] >= 690, c
    sval, sidx = oracle.find_matches_spec(bank, qry, 50)
    np.testing.assert_array_equal(idx, sidx)
    np.testing.assert_array_equal(val, sval)


def test_massive_ties_fall_back_and_stay_exact():
    bank = synth.pm1_embeddings(40000, 256, 9)
    qry = synth.pm1_embeddings(64, 256, 10)
    val, idx = retrieval.find_matches(bank, qry, 50, return_values=True)
    sval, sidx = oracle.find_matches_spec(bank, qry, 50)
    np.testing.assert_array_equal(idx, sidx)
    np.testing.assert_array_equal(val, sval)
    print(retrieval.last_counters())


def test_nonfinite_inputs_take_the_exact_path():
    bank = synth.embeddings(900, 256, 3)
    qry = synth.embeddings(20, 256, 4)
    bank[5, 7] = np.inf
    val, idx = retrieval.find_matches(bank, qry, 10, return_values=True)
    val_e, idx_e = retrieval.find_matches(bank, qry, 10, return_values=True, exact_only=True)
    np.testing.assert_array_equal(idx, idx_e)
    assert retrieval.last_counters() is not None
