"""Seeded synthetic inputs shaped like the reference's data.

Everything here is host-side numpy so the same bytes can be regenerated on the
GPU box, in the oracle and in the golden-fixture script without shipping data.
The reference ships no expression/embedding fixtures (only gene-name lists in
/root/reference/data), so these generators stand in for:

  * spot / image embeddings: rows leaving ``ProjectionHead``'s LayerNorm(256)
    (reference model.py:151-168) have mean 0 and norm ~16;
  * expression rows: log-normalised counts as produced by ``scprep`` at
    reference dataset.py:188 (non-negative, many zeros, values in [0, ~4]);
  * positions: ST array coordinates (< 64, reference dataset.py:195) or Visium
    pixel coordinates (< ~2e4, reference dataset.py:339) stored as float32.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "embeddings", "expression", "positions", "image_features", "pm1_embeddings",
    "CONFIGS",
]

# Concrete shapes of the BASELINE.json configs.
CONFIGS = {
    "cfg1": dict(N=9000, Q=600, D=256, k=50, G=1000, p=1, B=256),
    "cfg2": dict(B=1024, G=1000, G_real=171, L=2, E=1024, D=256),
    "cfg3": dict(N=30000, Q=4000, D=256, k=50, G=1000, p=2),
    "cfg4": dict(N=1_000_000, Q=65536, D=256, k=50, G=1000, p=2),
    "cfg5": dict(B_sweep=[256, 512, 1024, 2048, 4096, 8192, 16384, 32768], D=256),
}


def _rng(seed: int) -> np.random.Generator:
    return np.random.default_rng(np.random.PCG64(seed))


def embeddings(rows: int, dim: int = 256, seed: int = 0, flavour: str = "iid",
               centres: int = 64) -> np.ndarray:
    """float32 [rows, dim] embeddings.

    ``iid``: standard normal.  ``clustered``: ``centres`` cluster centres
    ``4*randn`` plus unit noise, then each row standardised to mean 0 / unit
    variance (what LayerNorm(256) leaves: norm == sqrt(dim)).
    """
    g = _rng(seed)
    if flavour == "iid":
        return g.standard_normal((rows, dim), dtype=np.float32)
    if flavour == "clustered":
        c = 4.0 * g.standard_normal((centres, dim), dtype=np.float32)
        which = g.integers(0, centres, size=rows)
        x = c[which] + g.standard_normal((rows, dim), dtype=np.float32)
        x -= x.mean(axis=1, keepdims=True)
        x /= x.std(axis=1, keepdims=True) + 1e-6
        return x.astype(np.float32)
    raise ValueError(f"unknown flavour {flavour!r}")


def pm1_embeddings(rows: int, dim: int = 256, seed: int = 0) -> np.ndarray:
    """Exact-arithmetic known-answer inputs: entries are +-1, so the L2 norm is
    exactly sqrt(dim) (16 for dim 256), normalised entries are +-1/16 and every
    dot product is a multiple of 1/dim -- exact in fp16, bf16, TF32, fp32 and
    fp64 alike.  Similarities therefore tie massively and implementation
    independent, which pins the top-k tie rule (lowest index first)."""
    g = _rng(seed)
    return (g.integers(0, 2, size=(rows, dim)).astype(np.float32) * 2.0 - 1.0)


def expression(rows: int, genes: int, seed: int = 0, dtype=np.float32) -> np.ndarray:
    """Log-normalised synthetic counts: Poisson(Gamma(0.5, 2)) -> library-size
    normalise to 1e4 -> log10(1+x)  (mirrors reference dataset.py:188)."""
    g = _rng(seed)
    out = np.empty((rows, genes), dtype=dtype)
    step = 65536
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        lam = g.gamma(0.5, 2.0, size=(r1 - r0, genes))
        cnt = g.poisson(lam).astype(np.float64)
        lib = cnt.sum(axis=1, keepdims=True)
        lib[lib == 0] = 1.0
        out[r0:r1] = np.log10(1.0 + cnt / lib * 1e4).astype(dtype)
    return out


def positions(rows: int, seed: int = 0, kind: str = "st") -> np.ndarray:
    """float32 [rows, 2] coordinates; ``.long()`` truncation happens downstream
    exactly as at reference model.py:230-231."""
    g = _rng(seed)
    hi = 64 if kind == "st" else 20000
    return g.integers(0, hi, size=(rows, 2)).astype(np.float32)


def image_features(rows: int, dim: int = 1024, seed: int = 0) -> np.ndarray:
    """Stand-in for the stock CNN output (reference model.py:72-84 -> [B,1024])."""
    return _rng(seed).standard_normal((rows, dim), dtype=np.float32)
