"""On-disk formats of the reference's evaluation scripts, loaded straight into the (sharded)
retrieval bank -- SURVEY.md section 8(f) rank 4.  Pure host-side IO: NumPy only, no CUDA.

Reference layout (paths relative to /root/reference):
  * ``save_embeddings`` (evel_her2st.py:87-117, same in evel_visium.py / evel_cscc.py) writes, per
    fold directory, ``img_embeddings_{i+1}.npy`` and ``spot_embeddings_{i+1}.npy`` for slide i,
    each TRANSPOSED: ``[256, n_i]`` float32;
  * ``preprocessed_matrix.npy`` per slide is ``[G, n_i]`` (hvg_her2st.py:123; read at
    evel_her2st.py:126-127, :136-137);
  * the fold loop (evel_her2st.py:145-172) drops slide ``fold`` from the bank, concatenates the
    rest along axis 1, and transposes whatever does not already have 256 columns / matching rows.

``load_fold`` reproduces that result; with ``rank``/``world`` it materialises only this rank's
contiguous range of bank rows (the split of ``distributed.shard_bounds``), reading the slices it
needs from memory-mapped files, so an 8-GPU job never holds the whole bank on one host buffer.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

__all__ = ["FoldData", "save_fold_embeddings", "load_fold", "slide_sizes"]


@dataclass
class FoldData:
    spot_key: np.ndarray          # [n_local, dim]  bank embeddings (rows [row0, row0 + n_local))
    expression_key: np.ndarray    # [n_local, G]    bank expression
    image_query: np.ndarray       # [Q, dim]        held-out slide's image embeddings
    expression_gt: np.ndarray     # [Q, G]          held-out slide's measured expression
    index_offset: int             # global bank row of local row 0
    n_total: int                  # bank rows over all shards


def save_fold_embeddings(save_path: str, img_embeddings_all: np.ndarray, spot_embeddings_all: np.ndarray,
                         datasize: Sequence[int]) -> None:
    """The file layout of ``save_embeddings`` (evel_her2st.py:109-117): rows
    ``sum(datasize[:i]) .. sum(datasize[:i+1])`` of both arrays go to slide i+1's files, transposed."""
    img = np.asarray(img_embeddings_all)
    spot = np.asarray(spot_embeddings_all)
    if img.shape[0] != sum(datasize) or spot.shape[0] != sum(datasize):
        raise ValueError("datasize does not add up to the number of embedded spots")
    os.makedirs(save_path, exist_ok=True)
    start = 0
    for i, n in enumerate(datasize):
        np.save(os.path.join(save_path, f"img_embeddings_{i + 1}.npy"), img[start:start + n].T)
        np.save(os.path.join(save_path, f"spot_embeddings_{i + 1}.npy"), spot[start:start + n].T)
        start += n


def _rows_view(a: np.ndarray, cols: int) -> np.ndarray:
    """``[n, cols]`` view of a per-slide array.  Files are written transposed (``[cols, n]``,
    evel_her2st.py:116-117); an array that is already ``[n, cols]`` is accepted too.  The reference
    decides by shape AFTER concatenation (evel_her2st.py:161-172: transpose unless shape[1] is
    already 256), which silently skips the transpose for a slide / bank of exactly 256 spots; here
    a square array follows the file convention instead."""
    if a.ndim != 2:
        raise ValueError(f"expected a 2-D array, got shape {a.shape}")
    if a.shape[0] == cols:
        return a.T
    if a.shape[1] == cols:
        return a
    raise ValueError(f"array of shape {a.shape} has no axis of length {cols}")


def slide_sizes(expression_paths: Sequence[str]) -> List[int]:
    """``datasize`` of evel_her2st.py:126-127: spots per slide = shape[1] of the [G, n] matrices."""
    return [int(np.load(p, mmap_mode="r").shape[1]) for p in expression_paths]


def load_fold(embedding_dir: str, expression_paths: Sequence[str], fold: int, dim: int = 256,
              rank: int = 0, world: int = 1, mmap: bool = True,
              expression_dtype: Optional[np.dtype] = None) -> FoldData:
    """Everything the fold-loop body needs (evel_her2st.py:145-172) for held-out slide ``fold``
    (0-based, like the reference's loop variable).

    embedding_dir     directory holding ``spot_embeddings_{i+1}.npy`` / ``img_embeddings_{i+1}.npy``
    expression_paths  one ``preprocessed_matrix.npy`` ([G, n_i]) per slide, in slide order
    rank, world       materialise only bank rows ``shard_bounds(N, world)[rank]``
    expression_dtype  cast the expression rows (default: keep the file's dtype)
    """
    n_slides = len(expression_paths)
    if not 0 <= fold < n_slides:
        raise ValueError(f"fold {fold} outside 0..{n_slides - 1}")
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    mode = "r" if mmap else None
    expr = [np.load(p, mmap_mode=mode) for p in expression_paths]
    genes = int(expr[fold].shape[0])
    sizes = [int(e.shape[1]) for e in expr]
    for i, e in enumerate(expr):
        if e.ndim != 2 or e.shape[0] != genes:
            raise ValueError(f"{expression_paths[i]}: expected [{genes}, n], got {e.shape}")
    rest = [i for i in range(n_slides) if i != fold]
    n_total = sum(sizes[i] for i in rest)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    e_dtype = np.dtype(expression_dtype) if expression_dtype is not None else expr[fold].dtype
    spot_key = np.empty((hi - lo, dim), dtype=np.float32)
    expression_key = np.empty((hi - lo, genes), dtype=e_dtype)
    start = 0
    for i in rest:
        a, b = max(lo, start), min(hi, start + sizes[i])
        if a < b:
            emb = np.load(os.path.join(embedding_dir, f"spot_embeddings_{i + 1}.npy"), mmap_mode=mode)
            rows = _rows_view(emb, dim)
            if rows.shape[0] != sizes[i]:
                raise ValueError(f"slide {i + 1}: {rows.shape[0]} embeddings for {sizes[i]} spots")
            spot_key[a - lo:b - lo] = rows[a - start:b - start]
            expression_key[a - lo:b - lo] = expr[i][:, a - start:b - start].T
        start += sizes[i]
    img = np.load(os.path.join(embedding_dir, f"img_embeddings_{fold + 1}.npy"), mmap_mode=mode)
    image_query = np.ascontiguousarray(_rows_view(img, dim), dtype=np.float32)
    if image_query.shape[0] != sizes[fold]:
        raise ValueError(f"slide {fold + 1}: {image_query.shape[0]} image embeddings for {sizes[fold]} spots")
    expression_gt = np.ascontiguousarray(np.asarray(expr[fold]).T)
    return FoldData(spot_key, expression_key, image_query, expression_gt, lo, n_total)
