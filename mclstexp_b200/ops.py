"""Thin Python wrappers over the dense-contraction entry points of libmclst_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import check, load, ptr, require_cuda, stream_ptr

_ws_cache = {}
WS_KEEP_MAX = 1 << 30      # scratch above this size is not kept between calls (a B = 32k loss needs ~20 GB)


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Per (device, stream) scratch for the C-ABI calls.  Requests up to WS_KEEP_MAX reuse one cached
    buffer (the training-step sizes; also what a captured CUDA graph keeps pointing at); larger ones
    get a fresh tensor that the caching allocator takes back as soon as the caller drops it."""
    if nbytes > WS_KEEP_MAX:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def matmul(a: torch.Tensor, b: torch.Tensor, a_trans: bool = False, b_trans: bool = False,
           alpha: float = 1.0, bias: Optional[torch.Tensor] = None, act: str = "none",
           residual: Optional[torch.Tensor] = None, precise: bool = True,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(alpha * op(a) @ op(b).T + bias) + residual on the tensor cores (float32).

    2-D: op(a) is [M,K] (``a`` stored [K,M] when ``a_trans``), op(b) is [N,K] (``b`` stored
    [K,N] when ``b_trans``).  3-D inputs are batched over dim 0 (any strides with unit inner
    stride, e.g. head slices of a qkv matrix)."""
    require_cuda(a, b, bias, residual)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == b.dim()
    batched = a.dim() == 3
    a3 = a if batched else a[None]
    b3 = b if batched else b[None]
    assert a3.stride(2) == 1 and b3.stride(2) == 1 and a3.shape[0] == b3.shape[0]
    nb = a3.shape[0]
    K, M = (a3.shape[1], a3.shape[2]) if a_trans else (a3.shape[2], a3.shape[1])
    Kb, N = (b3.shape[1], b3.shape[2]) if b_trans else (b3.shape[2], b3.shape[1])
    assert K == Kb, (a.shape, b.shape, a_trans, b_trans)
    lib = load()
    if out is None:
        out = torch.empty((nb, M, N) if batched else (M, N), dtype=torch.float32, device=a.device)
    o3 = out if batched else out[None]
    assert o3.shape == (nb, M, N) and o3.stride(2) == 1
    r3 = None
    if residual is not None:
        r3 = residual if batched else residual[None]
        assert r3.shape == o3.shape and r3.stride() == o3.stride()
    if M == 0 or N == 0 or nb == 0:
        return out
    if K == 0:
        raise ValueError("matmul: empty contraction")
    nbytes = C.c_size_t()
    check(lib.mclst_matmul_workspace_bytes(M, N, K, nb, C.byref(nbytes)), "matmul_workspace_bytes")
    ws = _workspace(nbytes.value, a.device)
    with torch.cuda.device(a.device):
        check(lib.mclst_matmul(ptr(a3), a3.stride(1), int(a_trans), a3.stride(0) if nb > 1 else 0,
                               ptr(b3), b3.stride(1), int(b_trans), b3.stride(0) if nb > 1 else 0,
                               ptr(o3), o3.stride(1), o3.stride(0) if nb > 1 else 0,
                               M, N, K, nb, float(alpha), ptr(bias), {"none": 0, "gelu": 1}[act],
                               ptr(r3), int(precise), ptr(ws), ws.numel(), stream_ptr()), "matmul")
    return out


def matmul_nt(a, b, alpha=1.0, bias=None, act="none", residual=None, precise=True, out=None):
    """act(alpha * a @ b.T + bias) + residual; a [M,K], b [N,K]."""
    return matmul(a, b, False, False, alpha, bias, act, residual, precise, out)
