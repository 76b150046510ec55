"""Fused bank / query embedding forward for evaluation.

The reference's ``get_embeddings`` (evel_her2st.py:30-71) walks the spots in un-shuffled batches
of 32 and runs ~30 small launches per batch; the spot "sequence" of the self-attention is the
batch itself (model.py:236), so tokens only attend inside their batch.  Here every per-row
operation (position embeddings, LayerNorm, the linear layers, GELU, the projection heads) runs
once over all spots, and the attention runs block-diagonally: scores are computed on tiles of
128 consecutive tokens with one batched tensor-core product per head and a block-diagonal softmax
keeps each token inside its own batch of ``group`` -- the same numbers as the per-batch loop.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from ._lib import check, load, ptr, require_cuda, stream_ptr
from .model import embed_add, gelu, layer_norm, linear

TILE = 128


def _grouped_attention(attn, x: torch.Tensor, group: int, n_valid: int) -> torch.Tensor:
    """model.py:49-57 for tokens that attend only inside consecutive groups.  x: [Np, dim], Np a
    multiple of TILE; returns the attention output before ``to_out`` ([Np, heads*dh])."""
    heads = attn.heads
    qkv = linear(x, attn.to_qkv.weight)                        # [Np, 3*inner]
    Np, three_inner = qkv.shape
    inner = three_inner // 3
    dh = inner // heads
    nt = Np // TILE
    out = torch.empty((Np, inner), dtype=torch.float32, device=x.device)
    scores = torch.empty((nt, TILE, TILE), dtype=torch.float32, device=x.device)
    lib = load()
    for h in range(heads):
        q, k, v = (qkv[:, i * inner + h * dh: i * inner + (h + 1) * dh].unflatten(0, (nt, TILE))
                   for i in range(3))
        ops.matmul(q, k, alpha=attn.scale, out=scores)                 # [nt, 128, 128]
        with torch.cuda.device(x.device):
            check(lib.mclst_softmax_blockdiag(ptr(scores), TILE, nt * TILE, TILE, group, n_valid,
                                              stream_ptr()), "softmax_blockdiag")
        ops.matmul(scores, v, b_trans=True, out=out[:, h * dh:(h + 1) * dh].unflatten(0, (nt, TILE)))
    return out


@torch.no_grad()
def embed_spots(model, expression: torch.Tensor, position: torch.Tensor, group: int = 32,
                chunk: int = 1 << 16) -> torch.Tensor:
    """Spot embeddings [N, projection_dim] exactly as evel_her2st.py:52-70 produces them with
    ``DataLoader(batch_size=group, shuffle=False)``: spot i is embedded together with the spots of
    its batch i // group (the last batch may be shorter)."""
    require_cuda(expression, position)
    if TILE % group != 0:
        raise ValueError(f"group {group} must divide {TILE}")
    N = expression.shape[0]
    chunk = max(TILE, chunk // TILE * TILE)
    outs = []
    for c0 in range(0, N, chunk):
        n = min(chunk, N - c0)
        npad = (n + TILE - 1) // TILE * TILE
        h = embed_add(expression[c0:c0 + n], position[c0:c0 + n], model.x_embed.weight, model.y_embed.weight)
        if npad != n:                                              # dummy tokens, masked out of every softmax
            h = torch.cat([h, h.new_zeros(npad - n, h.shape[1])])
        for blk in model.spot_encoder:
            a = blk.attn
            y = layer_norm(h, a.norm.weight, a.norm.bias, a.norm.eps)
            o = _grouped_attention(a.fn, y, group, n)
            if isinstance(a.fn.to_out, torch.nn.Identity):
                h = o + h
            else:
                h = linear(o, a.fn.to_out[0].weight, a.fn.to_out[0].bias, h)
            f = blk.ff
            y = layer_norm(h, f.norm.weight, f.norm.bias, f.norm.eps)
            t = gelu(linear(y, f.fn.net[0].weight, f.fn.net[0].bias))
            h = linear(t, f.fn.net[3].weight, f.fn.net[3].bias, h)
        outs.append(model.spot_projection(h)[:n])
    return torch.cat(outs) if len(outs) > 1 else outs[0]


@torch.no_grad()
def embed_bank(model, expression: torch.Tensor, position: torch.Tensor,
               image_features: Optional[torch.Tensor] = None, group: int = 32
               ) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """(image_embeddings | None, spot_embeddings): what ``get_embeddings`` (evel_her2st.py:30-71)
    concatenates over its loader; ``image_features`` are the CNN outputs (the CNN stays stock)."""
    img = model.image_projection(image_features) if image_features is not None else None
    return img, embed_spots(model, expression, position, group)
