"""Evaluation metrics of the fold loop on the GPU (reference evel_her2st.py:190-221 with
utils.py:52-65): mean per-gene Pearson correlation over all HVGs (NaN columns dropped) and over
the 50 most highly expressed genes, MSE and MAE."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from ._lib import check, load, ptr, require_cuda, stream_ptr


def gene_metrics_device(true: torch.Tensor, pred: torch.Tensor):
    """(mean_true, pcc, sq_err, abs_err) float64 [G] CUDA tensors for [Q,G] matrices."""
    require_cuda(true, pred)
    assert true.shape == pred.shape and true.dim() == 2
    assert true.dtype in (torch.float32, torch.float64) and pred.dtype in (torch.float32, torch.float64)
    assert true.stride(1) == 1 and pred.stride(1) == 1
    Q, G = true.shape
    lib = load()
    n = C.c_size_t()
    check(lib.mclst_gene_metrics_scratch_doubles(G, C.byref(n)), "gene_metrics_scratch")
    scratch = torch.empty(n.value, dtype=torch.float64, device=true.device)
    outs = [torch.empty(G, dtype=torch.float64, device=true.device) for _ in range(4)]
    with torch.cuda.device(true.device):
        check(lib.mclst_gene_metrics(ptr(true), true.stride(0), int(true.dtype == torch.float64), ptr(pred),
                                     pred.stride(0), int(pred.dtype == torch.float64), Q, G, ptr(outs[0]),
                                     ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), ptr(scratch), stream_ptr()),
              "gene_metrics")
    return tuple(outs)


def evaluate(true, pred, top_genes: int = 50) -> Dict[str, float]:
    """{'heg_pcc', 'hvg_pcc', 'mse', 'mae'} exactly as printed per fold at evel_her2st.py:207-221."""
    t = true if isinstance(true, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(true))
    p = pred if isinstance(pred, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pred))
    if not torch.cuda.is_available():
        from ._lib import MclstError
        raise MclstError("no CUDA device: metrics have no CPU fallback")
    t, p = t.cuda(), p.cuda()
    if t.dtype not in (torch.float32, torch.float64):
        t = t.float()
    if p.dtype not in (torch.float32, torch.float64):
        p = p.float()
    mean_true, pcc, sq, ab = gene_metrics_device(t.contiguous(), p.contiguous())
    Q, G = t.shape
    # np.argsort(mean)[::-1][:50]  (evel_her2st.py:202): descending, ties in reversed stable order
    order = torch.flip(torch.argsort(mean_true, stable=True), dims=[0])[:top_genes]
    heg = pcc[order]
    hvg = pcc[~torch.isnan(pcc)]
    return {"heg_pcc": float(heg.mean()), "hvg_pcc": float(hvg.mean()),
            "mse": float(sq.sum() / (Q * G)), "mae": float(ab.sum() / (Q * G))}
