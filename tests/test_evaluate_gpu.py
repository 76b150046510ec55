"""The whole fold-loop body from files (io.load_fold -> retrieval.retrieve -> metrics.evaluate) against
the oracle pipeline on the same temporary files (evel_her2st.py:145-221)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mclstexp_b200 import evaluate as mev, io as mio, synth      # noqa: E402
from oracle import oracle                                        # noqa: E402


def _write_dataset(root, sizes, genes, seed):
    n = sum(sizes)
    img = synth.embeddings(n, 256, seed, "clustered")
    spot = synth.embeddings(n, 256, seed + 1, "clustered")
    expr = synth.expression(n, genes, seed + 2)                     # [n, G]
    paths = []
    for fold in range(len(sizes)):
        mio.save_fold_embeddings(os.path.join(root, f"embeddings_{fold}"), img, spot, sizes)
    start = 0
    for i, s in enumerate(sizes):
        d = os.path.join(root, f"slide{i}")
        os.makedirs(d)
        np.save(os.path.join(d, "preprocessed_matrix.npy"), expr[start:start + s].T.astype(np.float64))
        paths.append(os.path.join(d, "preprocessed_matrix.npy"))
        start += s
    return paths


@pytest.mark.parametrize("p,k", [(1, 50), (2, 20)])
def test_evaluate_fold_matches_oracle_pipeline(tmp_path, p, k):
    sizes = [310, 190, 260]
    paths = _write_dataset(str(tmp_path), sizes, genes=64, seed=900 + p)
    fold = 1
    got = mev.evaluate_fold(os.path.join(str(tmp_path), f"embeddings_{fold}"), paths, fold, top_k=k, p=p,
                            return_prediction=True)
    fd = mio.load_fold(os.path.join(str(tmp_path), f"embeddings_{fold}"), paths, fold)
    mode = {1: "inv_sq_l1", 2: "inv_sq_l2"}[p]
    ref_idx, ref_emb, ref_expr = oracle.retrieve_ref(fd.spot_key, fd.expression_key, fd.image_query, k, mode)
    ok = oracle.decidable_rows(fd.spot_key, fd.image_query, k)
    assert ok.mean() > 0.5
    np.testing.assert_array_equal(got["indices"][ok], ref_idx[ok])
    np.testing.assert_allclose(got["expr_pred"][ok], ref_expr[ok], rtol=1e-3, atol=1e-6)
    ref_scores = oracle.metrics_ref(fd.expression_gt, got["expr_pred"])
    for key in ("heg_pcc", "hvg_pcc", "mse", "mae"):
        assert abs(got[key] - ref_scores[key]) <= 1e-6 * max(1.0, abs(ref_scores[key])), key
    if ok.all():
        end_to_end = oracle.metrics_ref(fd.expression_gt, ref_expr)
        for key in ("heg_pcc", "hvg_pcc", "mse", "mae"):
            assert abs(got[key] - end_to_end[key]) <= 1e-3 * max(1.0, abs(end_to_end[key])), key


def test_evaluate_folds_summary(tmp_path):
    sizes = [120, 150, 140]
    paths = _write_dataset(str(tmp_path), sizes, genes=40, seed=77)
    out = mev.evaluate_folds(os.path.join(str(tmp_path), "embeddings_{fold}"), paths, top_k=10, p=2)
    assert out["folds"] == [0, 1, 2] and len(out["per_fold"]) == 3
    for key in ("heg_pcc", "hvg_pcc", "mse", "mae"):
        vals = [r[key] for r in out["per_fold"]]
        assert abs(out[f"{key}_mean"] - float(np.mean(vals))) < 1e-12
        assert np.isfinite(out[f"{key}_mean"])
