"""Split-precision tcgen05 GEMM against a float64 product."""
import numpy as np
import pytest
import torch

from mclstexp_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1, 1, 1), (37, 300, 171), (1024, 1536, 785),
                                   (129, 257, 1000), (1024, 256, 2048), (300, 1000, 512)])
def test_matmul_nt_precise(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g)
    want = a.double() @ b.double().T
    got = ops.matmul_nt(a.cuda(), b.cuda()).cpu().double()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 3e-5 * scale + 1e-6, (err, scale)     # fp32-GEMM-level: far inside the 1e-3 budget
    fast = ops.matmul_nt(a.cuda(), b.cuda(), precise=False).cpu().double()
    assert (fast - want).abs().max().item() <= 2e-3 * scale + 1e-3


def test_matmul_nt_epilogue():
    g = torch.Generator().manual_seed(5)
    M, N, K = 200, 333, 96
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.2
    bias, res = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    want = torch.nn.functional.gelu(0.5 * a.double() @ b.double().T + bias.double()) + res.double()
    got = ops.matmul_nt(a.cuda(), b.cuda(), alpha=0.5, bias=bias.cuda(), act="gelu",
                        residual=res.cuda()).cpu().double()
    assert (got - want).abs().max().item() <= 5e-6 * want.abs().max().item()


def test_matmul_nt_strided_views():
    g = torch.Generator().manual_seed(6)
    big_a, big_b = torch.randn(70, 300, generator=g).cuda(), torch.randn(90, 300, generator=g).cuda()
    a, b = big_a[:, 10:138], big_b[:, 20:148]
    out = torch.zeros(70, 128, device="cuda")[:, :90]
    ops.matmul_nt(a, b, out=out)
    want = a.double() @ b.double().T
    assert (out.double() - want).abs().max().item() <= 1e-5 * want.abs().max().item()


@pytest.mark.parametrize("a_trans,b_trans", [(False, True), (True, False), (True, True)])
def test_matmul_transposes(a_trans, b_trans):
    g = torch.Generator().manual_seed(11)
    M, N, K = 150, 200, 333
    a = torch.randn((K, M) if a_trans else (M, K), generator=g).cuda()
    b = torch.randn((K, N) if b_trans else (N, K), generator=g).cuda()
    opa = a.T if a_trans else a
    opb = b.T if b_trans else b
    want = opa.double() @ opb.double().T
    got = ops.matmul(a, b, a_trans, b_trans)
    assert (got.double() - want).abs().max().item() <= 3e-5 * want.abs().max().item()


def test_matmul_batched_head_slices():
    """The attention pattern: per-head [B,64] slices of a [B,1536] qkv matrix."""
    g = torch.Generator().manual_seed(12)
    Bt, H, dh = 300, 8, 64
    qkv = torch.randn(Bt, 3 * H * dh, generator=g).cuda()
    q = qkv[:, :H * dh].view(Bt, H, dh).permute(1, 0, 2)          # [H, B, dh] strided view
    k = qkv[:, H * dh:2 * H * dh].view(Bt, H, dh).permute(1, 0, 2)
    v = qkv[:, 2 * H * dh:].view(Bt, H, dh).permute(1, 0, 2)
    dots = ops.matmul(q, k, alpha=0.125)
    want = torch.einsum("hid,hjd->hij", q.double(), k.double()) * 0.125
    assert (dots.double() - want).abs().max().item() <= 3e-5 * want.abs().max().item()
    p = torch.softmax(dots, -1)
    out = torch.empty(Bt, H * dh, device="cuda")
    ov = out.view(Bt, H, dh).permute(1, 0, 2)
    ops.matmul(p, v, b_trans=True, out=ov)                        # [H,B,B] x [H,B(K),dh(N)]
    want = torch.einsum("hij,hjd->hid", p.double(), v.double()).permute(1, 0, 2).reshape(Bt, H * dh)
    assert (out.double() - want).abs().max().item() <= 3e-5 * want.abs().max().item()
