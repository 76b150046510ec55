"""Contrastive loss fwd+bwd through the C-ABI against the reference's golden outputs and the
oracle (literal autograd restatement + float64 closed form)."""
import numpy as np
import pytest
import torch

from mclstexp_b200 import loss as mloss, synth
from oracle import oracle
from conftest import golden_checksum

pytestmark = pytest.mark.gpu

RTOL = 1e-3      # north star: loss and gradients within 1e-3 relative error under FP32


def _check(S, I, T, targets, scale, ref=None, env=None):
    St = torch.tensor(S, device="cuda", requires_grad=True)
    It = torch.tensor(I, device="cuda", requires_grad=True)
    l = mloss.contrastive_loss(St, It, T, targets, scale)
    assert l.dim() == 0 and l.dtype == torch.float32
    (l * 1.0).backward()
    l64, dS64, dI64 = oracle.contrastive_loss_closed_form(S, I, T, targets, scale)
    np.testing.assert_allclose(l.item(), l64, rtol=RTOL)
    for got, want in ((St.grad, dS64), (It.grad, dI64)):
        got = got.cpu().numpy()
        # normwise (per-matrix) relative error, plus elementwise with an absolute floor
        assert np.linalg.norm(got - want) <= RTOL * np.linalg.norm(want)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=RTOL * np.abs(want).max())
    if ref is not None:
        np.testing.assert_allclose(l.item(), ref[0], rtol=RTOL)
        np.testing.assert_allclose(St.grad.cpu().numpy(), ref[1], rtol=RTOL, atol=RTOL * np.abs(ref[1]).max())
        np.testing.assert_allclose(It.grad.cpu().numpy(), ref[2], rtol=RTOL, atol=RTOL * np.abs(ref[2]).max())


@pytest.mark.parametrize("case", ["loss_b48", "loss_b65_t07", "loss_b16_t2"])
@pytest.mark.parametrize("tag", ["eye", "soft_div", "soft_mul"])
def test_loss_vs_reference_golden(golden_model, case, tag):
    z, meta = golden_model
    m = meta[case]
    S = synth.embeddings(m["B"], m["D"], m["seed"], "clustered", centres=5)
    I = synth.embeddings(m["B"], m["D"], m["seed"] + 1, "clustered", centres=5)
    if m["scaled"]:
        S *= 0.25
        I *= 0.25
    assert golden_checksum(S, I) == m["checksum"]
    targets = "eye" if tag == "eye" else "soft"
    scale = "mul" if tag.endswith("mul") else "div"
    _check(S, I, m["T"], targets, scale,
           ref=(z[f"{case}/{tag}/loss"], z[f"{case}/{tag}/dS"], z[f"{case}/{tag}/dI"]))


@pytest.mark.parametrize("B,D,T,scale_in", [(256, 256, 1.0, 1.0), (1024, 256, 1.0, 1.0), (130, 256, 0.5, 0.3),
                                            (257, 100, 2.0, 1.0), (1, 256, 1.0, 1.0), (2, 8, 1.0, 1.0)])
@pytest.mark.parametrize("targets", ["eye", "soft"])
def test_loss_shapes(B, D, T, scale_in, targets):
    # LayerNorm-like rows (norm 16): logits reach +-256 -- the regime the reference trains in
    S = synth.embeddings(B, D, 50 + B, "clustered", centres=7) * scale_in
    I = synth.embeddings(B, D, 51 + B, "clustered", centres=7) * scale_in
    _check(S, I, T, targets, "div")


@pytest.mark.parametrize("targets", ["eye", "soft"])
def test_loss_row_blocked_equals_single_block(targets, monkeypatch):
    """Force several row blocks (the B=32k path) on a small batch."""
    import subprocess, sys, os
    code = f"""
import numpy as np, torch
from mclstexp_b200 import loss as mloss, synth
from oracle import oracle
B, D = 700, 256
S = synth.embeddings(B, D, 1, 'clustered', centres=9) * 0.5
I = synth.embeddings(B, D, 2, 'clustered', centres=9) * 0.5
St = torch.tensor(S, device='cuda', requires_grad=True); It = torch.tensor(I, device='cuda', requires_grad=True)
l = mloss.contrastive_loss(St, It, 1.0, '{targets}'); l.backward()
l64, dS64, dI64 = oracle.contrastive_loss_closed_form(S, I, 1.0, '{targets}')
assert abs(l.item() - l64) <= 1e-3 * abs(l64), (l.item(), l64)
assert np.linalg.norm(St.grad.cpu().numpy() - dS64) <= 1e-3 * np.linalg.norm(dS64)
assert np.linalg.norm(It.grad.cpu().numpy() - dI64) <= 1e-3 * np.linalg.norm(dI64)
print('ok')
"""
    env = dict(os.environ, MCLST_LOSS_SCRATCH_MB="3")      # 3 MiB -> R = 128 rows per block
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_loss_no_grad_path():
    S = torch.tensor(synth.embeddings(64, 256, 3), device="cuda")
    I = torch.tensor(synth.embeddings(64, 256, 4), device="cuda")
    l = mloss.contrastive_loss(S, I, 1.0, "soft")
    l64, _, _ = oracle.contrastive_loss_closed_form(S.cpu().numpy(), I.cpu().numpy(), 1.0, "soft")
    np.testing.assert_allclose(l.item(), l64, rtol=RTOL)
