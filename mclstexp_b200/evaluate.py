"""The body of the reference's fold loop as one call: load the held-out slide and the bank from the
reference's on-disk layout, retrieve, score (evel_her2st.py:145-221; evel_visium.py:165-239;
evel_cscc.py:168-250 differ only in the dataset paths, top-k and the L1 / L2 distance).

    for fold in range(32):                                       # evel_her2st.py:143
        scores = evaluate_fold(f"./embedding_result/her2st_result/embeddings_{fold}/",
                               expression_paths, fold, top_k=200, p=1)

Host glue only: the numbers come from ``retrieval.retrieve`` and ``metrics.evaluate`` (CUDA)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import io as mio
from . import metrics, retrieval

__all__ = ["evaluate_fold", "evaluate_folds"]


def evaluate_fold(embedding_dir: str, expression_paths: Sequence[str], fold: int, top_k: int = 50,
                  p: int = 2, mode: Optional[str] = None, dim: int = 256, top_genes: int = 50,
                  return_prediction: bool = False) -> Dict[str, object]:
    """{'heg_pcc', 'hvg_pcc', 'mse', 'mae'} of held-out slide ``fold`` (+ 'indices', 'expr_pred',
    'emb_pred' with ``return_prediction``): evel_her2st.py:145-221."""
    fd = mio.load_fold(embedding_dir, expression_paths, fold, dim=dim)
    if fd.n_total < top_k:
        raise ValueError(f"bank of {fd.n_total} spots is smaller than top_k={top_k}")
    idx, emb, expr = retrieval.retrieve(fd.spot_key, fd.expression_key, fd.image_query, top_k=top_k, p=p,
                                        mode=mode, want_emb=return_prediction)
    scores: Dict[str, object] = dict(metrics.evaluate(fd.expression_gt, expr, top_genes=top_genes))
    if return_prediction:
        scores.update({"indices": idx, "expr_pred": expr, "emb_pred": emb})
    return scores


def evaluate_folds(embedding_dir_pattern: str, expression_paths: Sequence[str], top_k: int = 50, p: int = 2,
                   mode: Optional[str] = None, folds: Optional[Sequence[int]] = None) -> Dict[str, object]:
    """All folds (``embedding_dir_pattern.format(fold=fold)``) and the averages the scripts print at
    the end (evel_her2st.py:223-226: ``np.mean`` over folds; the variances are an extra)."""
    folds = list(range(len(expression_paths))) if folds is None else list(folds)
    per_fold = [evaluate_fold(embedding_dir_pattern.format(fold=f), expression_paths, f, top_k, p, mode)
                for f in folds]
    out: Dict[str, object] = {"folds": folds, "per_fold": per_fold}
    for key in ("heg_pcc", "hvg_pcc", "mse", "mae"):
        vals = np.array([r[key] for r in per_fold], dtype=np.float64)
        out[f"{key}_mean"] = float(np.mean(vals))
        out[f"{key}_var"] = float(np.var(vals))
    return out
