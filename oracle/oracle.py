"""CPU oracle for the mclSTExp contrastive-alignment + retrieval hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``mclstexp_b200/`` imports it and
the product path raises when the CUDA library is missing.

What it is: a restatement, in the reference's own arithmetic libraries (PyTorch
CPU ops and NumPy -- the reference has no native code), of
every function on the hot path.  Each function cites the reference lines it
follows (paths relative to /root/reference).

Parity pinning: the reference ships no tests, golden vectors or fixtures for
this path.  The oracle is therefore pinned against
OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD CONTAINER: ``oracle/make_golden.py``
imports /root/reference/model.py (stub ``timm``), AST-extracts ``find_matches``
from evel_her2st.py / evel_cscc.py and ``cross_entropy`` from
baselines/Bleep/models.py, executes them on seeded synthetic inputs and commits
the results under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every
oracle function against those fixtures.

Two flavours exist for the order-sensitive pieces:

* ``*_ref``  -- the literal restatement (same calls, same dtypes, same order);
* ``*_spec`` -- an implementation-independent float64 statement with the tie
  rule the north star fixes (top-k ties -> lowest index first).  The CUDA path
  is compared bit-for-bit against the spec; the spec is compared against the
  literal restatement on every row whose rank boundaries are separated by more
  than ``DECIDABLE_GAP`` (two float32 GEMMs with different summation order already
  disagree below that).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DECIDABLE_GAP = 1e-6

# --------------------------------------------------------------------------
# a9  find_matches
# --------------------------------------------------------------------------


def find_matches_ref(spot_embeddings, query_embeddings, top_k=1, return_values=False):
    """Literal restatement of evel_her2st.py:74-84 (== evel_visium.py:94-104);
    with ``return_values`` the cSCC flavour evel_cscc.py:74-84.  The reference's
    ``print(dot_similarity.shape)`` (evel_her2st.py:81) is omitted."""
    spot_embeddings = torch.tensor(spot_embeddings)            # :76
    query_embeddings = torch.tensor(query_embeddings)          # :77
    query_embeddings = F.normalize(query_embeddings, p=2, dim=-1)   # :78
    spot_embeddings = F.normalize(spot_embeddings, p=2, dim=-1)     # :79
    dot_similarity = query_embeddings @ spot_embeddings.T      # :80
    values, indices = torch.topk(dot_similarity.squeeze(0), k=top_k)  # :82
    if return_values:
        return values.cpu().numpy(), indices.cpu().numpy()     # evel_cscc.py:84
    return indices.cpu().numpy()                               # :84


def normalize_spec(x: np.ndarray) -> np.ndarray:
    """float32 statement of ``F.normalize(x, p=2, dim=-1)`` (evel_her2st.py:78-79)
    that does not depend on a reduction order: the squared norm is accumulated in
    float64 (exact products, ~1e-16 relative sum error), rounded once to float32,
    clamped at eps=1e-12 like ATen, and every element is divided by it with one
    IEEE float32 division."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    ss = np.einsum("ij,ij->i", x.astype(np.float64), x.astype(np.float64))
    nrm = np.sqrt(ss).astype(np.float32)
    nrm = np.maximum(nrm, np.float32(1e-12))
    return (x / nrm[:, None]).astype(np.float32)


def norms_spec(x: np.ndarray) -> np.ndarray:
    """float64 L2 norms of float32 rows, clamped at F.normalize's eps 1e-12."""
    x64 = np.ascontiguousarray(x, dtype=np.float32).astype(np.float64)
    return np.maximum(np.sqrt(np.einsum("ij,ij->i", x64, x64)), 1e-12)


def similarity_spec(spot_embeddings, query_embeddings) -> np.ndarray:
    """[Q,N] float32 similarities: the float64 cosine of the raw float32 rows --
    dot product accumulated in float64, divided by the float64 norm product
    (norms clamped at 1e-12 like F.normalize, evel_her2st.py:78-79) -- rounded once
    to float32.  This is the ranking key of the CUDA path and the value the cSCC
    flavour returns (evel_cscc.py:82-84); it is independent of summation order up
    to ~1e-16 before the single float32 rounding."""
    q64 = np.ascontiguousarray(query_embeddings, dtype=np.float32).astype(np.float64)
    s64 = np.ascontiguousarray(spot_embeddings, dtype=np.float32).astype(np.float64)
    if q64.ndim == 1:
        q64 = q64[None]
    return ((q64 @ s64.T) / (norms_spec(q64)[:, None] * norms_spec(s64)[None, :])).astype(np.float32)


def find_matches_spec(spot_embeddings, query_embeddings, top_k=1,
                      chunk: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """Tie-defined statement of evel_her2st.py:74-84: rank by (similarity
    descending, index ascending).  ``torch.topk`` leaves tie order unspecified
    (probe [1,3,3,2,3,0,3], k=2 -> [1,6]); the north star
    fixes it to lowest index first, i.e. a stable descending sort.
    Returns (values float32 [Q,k], indices int64 [Q,k])."""
    q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
    if q.ndim == 1:
        q = q[None]
    Q = q.shape[0]
    vals = np.empty((Q, top_k), np.float32)
    idx = np.empty((Q, top_k), np.int64)
    for q0 in range(0, Q, chunk):
        sim = similarity_spec(spot_embeddings, q[q0:q0 + chunk])
        order = np.argsort(-sim, axis=1, kind="stable")[:, :top_k]
        idx[q0:q0 + chunk] = order
        vals[q0:q0 + chunk] = np.take_along_axis(sim, order, axis=1)
    return vals, idx


def find_matches_spec_rows(spot_embeddings, query_embeddings, top_k: int,
                           bank_chunk: int = 1 << 17) -> Tuple[np.ndarray, np.ndarray]:
    """``find_matches_spec`` for a FEW query rows against a bank too large to hold as one float64
    array (the 1M-spot bank of BASELINE cfg4): the similarity is evaluated per bank chunk with
    ``similarity_spec`` -- identical arithmetic, every (query, spot) pair is independent of the
    chunking -- and the order is (similarity descending, index ascending)."""
    q = np.ascontiguousarray(query_embeddings, np.float32)
    if q.ndim == 1:
        q = q[None]
    n = spot_embeddings.shape[0]
    sims = np.concatenate([similarity_spec(spot_embeddings[c0:c0 + bank_chunk], q)
                           for c0 in range(0, n, bank_chunk)], axis=1)
    order = np.argsort(-sims, axis=1, kind="stable")[:, :top_k]
    return np.take_along_axis(sims, order, axis=1), order.astype(np.int64)


def decidable_rows(spot_embeddings, query_embeddings, top_k, gap=DECIDABLE_GAP) -> np.ndarray:
    """Boolean [Q]: rows whose top-(k+1) float64 similarities are pairwise
    separated by more than ``gap`` -- on those rows ANY correct float32
    implementation (the reference's MKL sgemm + torch.topk included) must return
    exactly the spec's ordered indices."""
    q64 = np.ascontiguousarray(query_embeddings, dtype=np.float32).astype(np.float64)
    s64 = np.ascontiguousarray(spot_embeddings, dtype=np.float32).astype(np.float64)
    sim = (q64 @ s64.T) / (norms_spec(q64)[:, None] * norms_spec(s64)[None, :])
    kk = min(top_k + 1, sim.shape[1])
    top = -np.sort(-sim, axis=1)[:, :kk]
    return (np.diff(-top, axis=1) > gap).all(axis=1)


# --------------------------------------------------------------------------
# a10 / a11  top-k weighted expression average
# --------------------------------------------------------------------------

WEIGHT_MODES = ("inv_sq_l1", "inv_sq_l2", "similarity", "uniform", "bleep_exp")


def weighted_average_ref(spot_key, expression_key, image_query, indices, mode="inv_sq_l1",
                         values=None):
    """Literal restatement of the per-query loops.

    ``inv_sq_l1``  evel_her2st.py:175-187 (``ord=1``)
    ``inv_sq_l2``  evel_visium.py:194-205 and evel_cscc.py:198-215 (default ord=2)
    ``similarity`` the commented-out variant evel_cscc.py:201-205
                   (``weights = value[i] / np.sum(value[i])``)
    ``uniform``    baselines/Bleep/BLEEP_inference.ipynb cell 5, ``average``
                   (``simple`` is ``uniform`` with k == 1)
    ``bleep_exp``  same cell, ``weighted_average``:
                   ``exp(-(d2_j - d2_best + 1))`` with squared L2 distances.

    Returns (emb_pred float64 [Q,D], expr_pred float64 [Q,G]) exactly like the
    ``np.zeros`` (float64) result arrays at evel_her2st.py:175-176.
    A zero distance makes the reference produce inf/inf = NaN; see
    ``weighted_average_spec`` for the defined behaviour."""
    indices = np.asarray(indices)
    if indices.ndim == 1:
        indices = indices[None]
    emb = np.zeros((indices.shape[0], spot_key.shape[1]))          # :175
    expr = np.zeros((indices.shape[0], expression_key.shape[1]))   # :176
    for i in range(indices.shape[0]):                              # :177
        rows = spot_key[indices[i, :], :]
        if mode == "inv_sq_l1":
            a = np.linalg.norm(rows - image_query[i, :], axis=1, ord=1)   # :178
            r = np.reciprocal(a ** 2)                                     # :182
            w = (r / np.sum(r)).flatten()                                 # :183-184
        elif mode == "inv_sq_l2":
            a = np.linalg.norm(rows - image_query[i, :], axis=1)   # evel_visium.py:197
            r = np.reciprocal(a ** 2)
            w = (r / np.sum(r)).flatten()
        elif mode == "similarity":
            w = values[i] / np.sum(values[i])                      # evel_cscc.py:201
        elif mode == "uniform":
            w = None                                               # nb cell 5 l.31-32
        elif mode == "bleep_exp":
            a = np.sum((spot_key[indices[i, 0], :] - image_query[i, :]) ** 2)   # nb cell 5 l.41
            w = np.exp(-(np.sum((rows - image_query[i, :]) ** 2, axis=1) - a + 1))  # l.42
        else:
            raise ValueError(mode)
        emb[i, :] = np.average(rows, axis=0, weights=w)                           # :185
        expr[i, :] = np.average(expression_key[indices[i, :], :], axis=0, weights=w)  # :186
    return emb, expr


def weights_spec(spot_key, image_query, indices, mode, values=None) -> np.ndarray:
    """float64 normalised weights [Q,k] with the zero-distance case DEFINED
    (SURVEY.md section 8c): neighbours at distance exactly 0 share weight 1 and
    every other neighbour gets 0 (the reference yields NaN there)."""
    sk = np.asarray(spot_key, np.float64)
    iq = np.asarray(image_query, np.float64)
    indices = np.asarray(indices)
    diff = sk[indices] - iq[:, None, :]
    if mode == "inv_sq_l1":
        d = np.abs(diff).sum(-1)
    elif mode in ("inv_sq_l2", "bleep_exp"):
        d = np.sqrt((diff ** 2).sum(-1))
    if mode in ("inv_sq_l1", "inv_sq_l2"):
        zero = d == 0
        with np.errstate(divide="ignore"):
            w = 1.0 / d ** 2
        anyz = zero.any(axis=1)
        w[anyz] = zero[anyz].astype(np.float64)
    elif mode == "similarity":
        w = np.asarray(values, np.float64).copy()
    elif mode == "uniform":
        w = np.ones(indices.shape, np.float64)
    elif mode == "bleep_exp":
        d2 = (diff ** 2).sum(-1)
        w = np.exp(-(d2 - d2[:, :1] + 1.0))
    else:
        raise ValueError(mode)
    return w / w.sum(axis=1, keepdims=True)


def weighted_average_spec(spot_key, expression_key, image_query, indices, mode="inv_sq_l1",
                          values=None) -> Tuple[np.ndarray, np.ndarray]:
    """float64 statement of ``weighted_average_ref`` (vectorised)."""
    indices = np.asarray(indices)
    w = weights_spec(spot_key, image_query, indices, mode, values)
    emb = np.einsum("qk,qkd->qd", w, np.asarray(spot_key, np.float64)[indices])
    Q = indices.shape[0]
    G = expression_key.shape[1]
    expr = np.empty((Q, G), np.float64)
    step = max(1, (1 << 24) // (indices.shape[1] * G))
    for q0 in range(0, Q, step):
        sl = slice(q0, q0 + step)
        expr[sl] = np.einsum("qk,qkg->qg", w[sl],
                             np.asarray(expression_key[indices[sl]], np.float64))
    return emb, expr


def retrieve_ref(spot_key, expression_key, image_query, top_k, mode="inv_sq_l2"):
    """The fold-loop body evel_her2st.py:174-187 (find_matches + loop)."""
    if mode == "similarity":
        values, indices = find_matches_ref(spot_key, image_query, top_k, return_values=True)
    else:
        values, indices = None, find_matches_ref(spot_key, image_query, top_k)
    if indices.ndim == 1:
        indices = indices[None]
        values = None if values is None else values[None]
    emb, expr = weighted_average_ref(spot_key, expression_key, image_query, indices, mode, values)
    return indices, emb, expr


# --------------------------------------------------------------------------
# downstream metrics
# --------------------------------------------------------------------------


def metrics_ref(true: np.ndarray, pred: np.ndarray, top_genes: int = 50) -> Dict[str, float]:
    """evel_her2st.py:201-221 with utils.py:52-65 (get_R): per-gene scipy pearsonr, the top-50
    genes by mean true expression, NaN columns dropped from the HVG mean, sklearn MSE / MAE
    (plain means over all entries)."""
    from scipy.stats import pearsonr
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = np.array([pearsonr(pred[:, g], true[:, g])[0] for g in range(true.shape[1])])   # utils.py:56-59
    gene_mean_expression = np.mean(true, axis=0)                                  # evel_her2st.py:201
    top = np.argsort(gene_mean_expression)[::-1][:top_genes]                       # :202
    heg = r[top]                                                                   # :207
    hvg = r[~np.isnan(r)]                                                          # :209
    return {"heg_pcc": float(np.mean(heg)), "hvg_pcc": float(np.mean(hvg)),       # :211-212
            "mse": float(np.mean((true - pred) ** 2)),                             # :216
            "mae": float(np.mean(np.abs(true - pred))), "pcc": r}                  # :219


# --------------------------------------------------------------------------
# a7 / a8  contrastive losses
# --------------------------------------------------------------------------


def eye_loss_ref(spot_embeddings: torch.Tensor, image_embeddings: torch.Tensor,
                 temperature: float) -> torch.Tensor:
    """model.py:242-247 with the hard-coded ``.cuda()`` (:243) dropped."""
    cos_smi = (spot_embeddings @ image_embeddings.T) / temperature       # :242
    label = torch.eye(cos_smi.shape[0], cos_smi.shape[1], dtype=cos_smi.dtype)  # :243
    spots_loss = F.cross_entropy(cos_smi, label)                         # :244
    images_loss = F.cross_entropy(cos_smi.T, label.T)                    # :245
    loss = (images_loss + spots_loss) / 2.0                              # :246
    return loss.mean()                                                   # :247


def _bleep_cross_entropy(preds, targets, reduction="none"):
    """baselines/Bleep/models.py:228-234."""
    log_softmax = torch.nn.LogSoftmax(dim=-1)                            # :229
    loss = (-targets * log_softmax(preds)).sum(1)                        # :230
    if reduction == "none":
        return loss
    elif reduction == "mean":
        return loss.mean()


def soft_loss_ref(spot_embeddings: torch.Tensor, image_embeddings: torch.Tensor,
                  temperature: float, soft_scale: str = "div") -> torch.Tensor:
    """baselines/Bleep/models.py:34-43 (``soft_scale='div'``: ``/ 2 / T``) and the
    ViT/CLIP/ResNet101/152 variant :70-79 (``'mul'``: ``/ 2 * T``).  The targets
    stay in the autograd graph exactly as in the reference (not detached)."""
    logits = (spot_embeddings @ image_embeddings.T) / temperature        # :34
    images_similarity = image_embeddings @ image_embeddings.T            # :35
    spots_similarity = spot_embeddings @ spot_embeddings.T               # :36
    if soft_scale == "div":
        targets = F.softmax(((images_similarity + spots_similarity) / 2) / temperature, dim=-1)  # :37-39
    elif soft_scale == "mul":
        targets = F.softmax((images_similarity + spots_similarity) / 2 * temperature, dim=-1)    # :73-75
    else:
        raise ValueError(soft_scale)
    spots_loss = _bleep_cross_entropy(logits, targets, reduction="none")      # :40
    images_loss = _bleep_cross_entropy(logits.T, targets.T, reduction="none")  # :41
    loss = (images_loss + spots_loss) / 2.0                              # :42
    return loss.mean()                                                   # :43


def contrastive_loss_ref(S, I, temperature, targets="eye", soft_scale="div", dtype=torch.float32):
    """loss + (dS, dI) through autograd of the literal restatements."""
    S = torch.as_tensor(S).detach().to(dtype).requires_grad_(True)
    I = torch.as_tensor(I).detach().to(dtype).requires_grad_(True)
    if targets == "eye":
        loss = eye_loss_ref(S, I, temperature)
    else:
        loss = soft_loss_ref(S, I, temperature, soft_scale)
    loss.backward()
    return loss.detach(), S.grad.detach(), I.grad.detach()


def contrastive_loss_closed_form(S, I, temperature, targets="eye", soft_scale="div"):
    """float64 closed form used by the fused kernels:
    W_ij = rl_i + cl_j - 2 Lg_ij; loss = (1/2B) sum_ij Pt_ij W_ij;
    dLg = (softmax_row + cs_j softmax_col - 2 Pt)/(2B); dA = Pt (W - wbar_i)/(2B)."""
    S = np.asarray(S, np.float64)
    I = np.asarray(I, np.float64)
    B = S.shape[0]
    T = float(temperature)
    Lg = S @ I.T / T
    rl = _lse(Lg, 1)
    cl = _lse(Lg, 0)
    W = rl[:, None] + cl[None, :] - 2 * Lg
    if targets == "eye":
        Pt = np.eye(B)
        a_scale = 0.0
    else:
        a_scale = 1.0 / (2 * T) if soft_scale == "div" else T / 2.0
        A = (I @ I.T + S @ S.T) * a_scale
        Pt = np.exp(A - _lse(A, 1)[:, None])
    loss = (Pt * W).sum() / (2 * B)
    cs = Pt.sum(0)
    dLg = (np.exp(Lg - rl[:, None]) + cs[None, :] * np.exp(Lg - cl[None, :]) - 2 * Pt) / (2 * B)
    dS = dLg @ I / T
    dI = dLg.T @ S / T
    if targets != "eye":
        wbar = (Pt * W).sum(1)
        dA = Pt * (W - wbar[:, None]) / (2 * B)
        sym = dA + dA.T
        dS += sym @ S * a_scale
        dI += sym @ I * a_scale
    return loss, dS, dI


def _lse(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


# --------------------------------------------------------------------------
# a1-a6  spot encoder + projection heads (functional, from a state_dict)
# --------------------------------------------------------------------------


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)       # nn.LayerNorm default eps


def attention_ref(x, sd: Dict[str, torch.Tensor], prefix: str, heads: int) -> torch.Tensor:
    """model.py:49-57 (``Attention.forward``) on x [b,n,dim]."""
    b, n, _ = x.shape
    qkv = F.linear(x, sd[prefix + "to_qkv.weight"]).chunk(3, dim=-1)     # :51
    dh = qkv[0].shape[-1] // heads
    q, k, v = [t.reshape(b, n, heads, dh).permute(0, 2, 1, 3) for t in qkv]   # :52
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * (dh ** -0.5)          # :53
    attn = dots.softmax(dim=-1)                                          # :54
    out = torch.einsum("bhij,bhjd->bhid", attn, v)                       # :55
    out = out.permute(0, 2, 1, 3).reshape(b, n, heads * dh)              # :56
    if prefix + "to_out.0.weight" in sd:
        out = F.linear(out, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])  # :57
    return out


def feed_forward_ref(x, sd, prefix: str) -> torch.Tensor:
    """model.py:23-32: Linear -> GELU(erf) -> Dropout(0) -> Linear -> Dropout(0)."""
    h = F.linear(x, sd[prefix + "net.0.weight"], sd[prefix + "net.0.bias"])
    h = F.gelu(h)
    return F.linear(h, sd[prefix + "net.3.weight"], sd[prefix + "net.3.bias"])


def attn_block_ref(x, sd, prefix: str, heads: int) -> torch.Tensor:
    """model.py:66-69 with PreNorm model.py:16-17."""
    h = _ln(x, sd[prefix + "attn.norm.weight"], sd[prefix + "attn.norm.bias"])
    x = attention_ref(h, sd, prefix + "attn.fn.", heads) + x             # :67
    h = _ln(x, sd[prefix + "ff.norm.weight"], sd[prefix + "ff.norm.bias"])
    x = feed_forward_ref(h, sd, prefix + "ff.fn.") + x                   # :68
    return x


def projection_head_ref(x, sd, prefix: str) -> torch.Tensor:
    """model.py:160-168."""
    projected = F.linear(x, sd[prefix + "projection.weight"], sd[prefix + "projection.bias"])  # :161
    h = F.gelu(projected)                                                # :162
    h = F.linear(h, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])    # :163
    h = h + projected                                                    # :165
    return _ln(h, sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"])  # :166


def spot_embedding_ref(sd, expression, position, heads: int, layers: int) -> torch.Tensor:
    """model.py:230-240 (== evel_her2st.py:52-69): position gathers + add,
    unsqueeze, ``layers`` attn_blocks, spot projection, squeeze."""
    x = position[:, 0].long()                                            # :230
    y = position[:, 1].long()                                            # :231
    centers_x = F.embedding(x, sd["x_embed.weight"])                     # :232
    centers_y = F.embedding(y, sd["y_embed.weight"])                     # :233
    h = expression + centers_x + centers_y                               # :235
    h = h.unsqueeze(dim=0)                                               # :236
    for l in range(layers):                                              # :238
        h = attn_block_ref(h, sd, f"spot_encoder.{l}.", heads)
    h = projection_head_ref(h, sd, "spot_projection.")                   # :239
    return h.squeeze(dim=0)                                              # :240


def path_loss_ref(sd, image_features, expression, position, temperature, heads, layers,
                  targets="eye", soft_scale="div") -> torch.Tensor:
    """model.py:225-247 with the stock CNN replaced by its output
    ``image_features`` (the image encoder is outside the path, SURVEY.md
    section 2 row 10)."""
    image_embeddings = projection_head_ref(image_features, sd, "image_projection.")   # :228
    spot_embeddings = spot_embedding_ref(sd, expression, position, heads, layers)
    if targets == "eye":
        return eye_loss_ref(spot_embeddings, image_embeddings, temperature)
    return soft_loss_ref(spot_embeddings, image_embeddings, temperature, soft_scale)


def make_state_dict(G: int, E: int = 1024, P: int = 256, heads: int = 8, dim_head: int = 64,
                    layers: int = 2, seed: int = 0, table_rows: int = 65536,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic (numpy PCG64) parameters with the reference's state_dict
    keys and shapes and PyTorch-default-like scales:
    Linear ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)), Embedding ~ N(0,1), LayerNorm
    weight ~ 1 + 0.1 N(0,1), bias ~ 0.1 N(0,1) (perturbed so that affine terms
    are exercised)."""
    g = np.random.default_rng(np.random.PCG64(seed))
    inner = heads * dim_head

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return torch.tensor(g.uniform(-b, b, size=(o, i)), dtype=dtype)

    def vec(n, scale, shift=0.0):
        return torch.tensor(shift + scale * g.standard_normal(n), dtype=dtype)

    sd = {
        "x_embed.weight": torch.tensor(g.standard_normal((table_rows, G)), dtype=dtype),
        "y_embed.weight": torch.tensor(g.standard_normal((table_rows, G)), dtype=dtype),
    }
    for l in range(layers):
        p = f"spot_encoder.{l}."
        sd[p + "attn.norm.weight"] = vec(G, 0.1, 1.0)
        sd[p + "attn.norm.bias"] = vec(G, 0.1)
        sd[p + "attn.fn.to_qkv.weight"] = lin(3 * inner, G)
        sd[p + "attn.fn.to_out.0.weight"] = lin(G, inner)
        sd[p + "attn.fn.to_out.0.bias"] = vec(G, 1.0 / math.sqrt(inner))
        sd[p + "ff.norm.weight"] = vec(G, 0.1, 1.0)
        sd[p + "ff.norm.bias"] = vec(G, 0.1)
        sd[p + "ff.fn.net.0.weight"] = lin(G, G)
        sd[p + "ff.fn.net.0.bias"] = vec(G, 1.0 / math.sqrt(G))
        sd[p + "ff.fn.net.3.weight"] = lin(G, G)
        sd[p + "ff.fn.net.3.bias"] = vec(G, 1.0 / math.sqrt(G))
    for name, e in (("image_projection.", E), ("spot_projection.", G)):
        sd[name + "projection.weight"] = lin(P, e)
        sd[name + "projection.bias"] = vec(P, 1.0 / math.sqrt(e))
        sd[name + "fc.weight"] = lin(P, P)
        sd[name + "fc.bias"] = vec(P, 1.0 / math.sqrt(P))
        sd[name + "layer_norm.weight"] = vec(P, 0.1, 1.0)
        sd[name + "layer_norm.bias"] = vec(P, 0.1)
    return sd


# --------------------------------------------------------------------------
# Adam on the position tables (train.py:118-120)
# --------------------------------------------------------------------------


def adam_step_ref(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3):
    """One step of torch.optim.Adam (L2 weight decay, no amsgrad) on float32 arrays, in the
    operation order of torch/optim/adam.py::_single_tensor_adam -- the order csrc/optim.cu follows.
    Returns the new (p, m, v)."""
    f32 = np.float32
    p, g, m, v = (np.asarray(a, dtype=f32) for a in (p, g, m, v))
    b1, b2 = betas
    g = g + f32(weight_decay) * p                                  # grad.add(param, alpha=wd)
    m = m + (g - m) * f32(1.0 - b1)                                # exp_avg.lerp_(grad, 1 - beta1)
    v = v * f32(b2) + f32(1.0 - b2) * g * g                        # mul_(beta2).addcmul_(grad, grad, 1-beta2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f32(bc2 ** 0.5) + f32(eps)
    p = p + f32(-step_size) * (m / denom)                          # addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(f32), m.astype(f32), v.astype(f32)


def adam_lazy_rows_ref(table, m, v, last, rows, row_grads, step, **hyper):
    """The deferral of optim.LazyEmbeddingAdam restated on the host: bring ``rows`` (distinct) from
    ``last[row]`` to ``step - 1`` with a zero data gradient, then apply ``step`` with ``row_grads``.
    Operates in place on float32 arrays; ``last`` is an int array of steps applied per row."""
    for r, g in zip(rows, row_grads):
        pr, mr, vr = table[r], m[r], v[r]
        for s in range(int(last[r]) + 1, step):
            pr, mr, vr = adam_step_ref(pr, np.zeros_like(pr), mr, vr, s, **hyper)
        pr, mr, vr = adam_step_ref(pr, g, mr, vr, step, **hyper)
        table[r], m[r], v[r], last[r] = pr, mr, vr, step


def adam_lazy_flush_ref(table, m, v, last, step, **hyper):
    for r in range(table.shape[0]):
        pr, mr, vr = table[r], m[r], v[r]
        for s in range(int(last[r]) + 1, step + 1):
            pr, mr, vr = adam_step_ref(pr, np.zeros_like(pr), mr, vr, s, **hyper)
        table[r], m[r], v[r], last[r] = pr, mr, vr, step
