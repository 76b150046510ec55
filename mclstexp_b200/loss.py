"""Contrastive loss of mclSTExp / BLEEP as one fused forward+backward call.

Mirrors reference model.py:242-247 (identity targets) and
baselines/Bleep/models.py:34-43 / :70-79 (soft targets, kept in the autograd graph).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, load, ptr, require_cuda, stream_ptr
from .ops import _workspace

TARGET_MODES = {("eye", "div"): _lib.T_EYE, ("eye", "mul"): _lib.T_EYE,
                ("soft", "div"): _lib.T_SOFT_DIV, ("soft", "mul"): _lib.T_SOFT_MUL}


def contrastive_loss_fwd_bwd(spot_emb: torch.Tensor, image_emb: torch.Tensor, temperature: float,
                             mode: int, want_grad: bool = True):
    """(loss 0-d f32, dS | None, dI | None) straight from libmclst_b200.so."""
    require_cuda(spot_emb, image_emb)
    assert spot_emb.dtype == torch.float32 and image_emb.dtype == torch.float32
    assert spot_emb.dim() == 2 and spot_emb.shape == image_emb.shape
    S = spot_emb if spot_emb.stride(1) == 1 else spot_emb.contiguous()
    I = image_emb if image_emb.stride(1) == 1 else image_emb.contiguous()
    B, D = S.shape
    lib = load()
    nbytes = C.c_size_t()
    check(lib.mclst_contrastive_loss_workspace_bytes(B, D, mode, B, C.byref(nbytes)),
          "contrastive_loss_workspace_bytes")
    ws = _workspace(nbytes.value, S.device)
    loss = torch.empty((), dtype=torch.float32, device=S.device)
    dS = torch.empty_like(S, memory_format=torch.contiguous_format) if want_grad else None
    dI = torch.empty_like(I, memory_format=torch.contiguous_format) if want_grad else None
    with torch.cuda.device(S.device):
        check(lib.mclst_contrastive_loss(ptr(S), S.stride(0), ptr(I), I.stride(0), B, D,
                                         float(temperature), mode, ptr(loss), ptr(dS),
                                         dS.stride(0) if want_grad else 0, ptr(dI),
                                         dI.stride(0) if want_grad else 0, ptr(ws), ws.numel(),
                                         stream_ptr()), "contrastive_loss")
    return loss, dS, dI


class _ContrastiveLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spot_emb, image_emb, temperature, mode):
        need = spot_emb.requires_grad or image_emb.requires_grad
        loss, dS, dI = contrastive_loss_fwd_bwd(spot_emb.detach(), image_emb.detach(), temperature,
                                                mode, want_grad=need)
        if need:
            ctx.save_for_backward(dS, dI)
        return loss

    @staticmethod
    def backward(ctx, g):
        dS, dI = ctx.saved_tensors
        return g * dS, g * dI, None, None


def contrastive_loss(spot_emb: torch.Tensor, image_emb: torch.Tensor, temperature: float = 1.0,
                     targets: str = "eye", soft_scale: str = "div") -> torch.Tensor:
    """0-d float32 loss with autograd.

    ``targets='eye'``  -- model.py:242-247 (``F.cross_entropy`` with identity probability targets,
                          both directions, averaged);
    ``targets='soft'`` -- baselines/Bleep/models.py:34-43 (``soft_scale='div'``) or :70-79
                          (``'mul'``): softmax of the averaged image-image / spot-spot similarity."""
    return _ContrastiveLoss.apply(spot_emb, image_emb, float(temperature),
                                  TARGET_MODES[(targets, soft_scale)])
