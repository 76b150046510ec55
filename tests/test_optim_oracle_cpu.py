"""Oracle side of the embedding-table optimiser: the NumPy restatement of one
Adam step against torch.optim.Adam itself (the reference's optimiser, train.py:118-120) on CPU, and
the row-deferral argument -- replaying a row's missed steps later gives the dense trajectory."""
import numpy as np
import torch
from torch import nn

from oracle import oracle

HYPER = dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3)


def test_adam_step_restatement_matches_torch_optim_adam():
    rng = np.random.default_rng(0)
    p0 = (rng.standard_normal((40, 17)) * 0.3).astype(np.float32)
    ref = nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([ref], **HYPER)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 31):
        g = (rng.standard_normal(p0.shape) * (0.1 if step % 4 else 0.0)).astype(np.float32)
        ref.grad = torch.from_numpy(g.copy())
        opt.step()
        p, m, v = oracle.adam_step_ref(p, g, m, v, step, **HYPER)
    st = opt.state[ref]
    np.testing.assert_allclose(p, ref.detach().numpy(), rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(m, st["exp_avg"].numpy(), rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(v, st["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-12)


def test_deferred_rows_follow_the_dense_trajectory_exactly():
    """Same arithmetic applied later is the same arithmetic: bit-identical tables after a flush,
    and at every moment a row is read."""
    rng = np.random.default_rng(1)
    R, G, steps = 60, 9, 20
    t0 = (rng.standard_normal((R, G)) * 0.2).astype(np.float32)
    dense, dm, dv = t0.copy(), np.zeros_like(t0), np.zeros_like(t0)
    lazy, lm, lv, last = t0.copy(), np.zeros_like(t0), np.zeros_like(t0), np.zeros(R, dtype=np.int64)
    for step in range(1, steps + 1):
        rows = np.unique(rng.integers(0, 12 if step % 4 else R, size=7))
        grads = (rng.standard_normal((len(rows), G)) * 0.05).astype(np.float32)
        g = np.zeros_like(t0)
        g[rows] = grads
        for r in range(R):                                   # dense Adam: every row, every step
            dense[r], dm[r], dv[r] = oracle.adam_step_ref(dense[r], g[r], dm[r], dv[r], step, **HYPER)
        oracle.adam_lazy_rows_ref(lazy, lm, lv, last, rows, grads, step, **HYPER)
        assert np.array_equal(lazy[rows], dense[rows])       # what the next forward would read
    assert (last < steps).any() and not np.array_equal(lazy, dense)
    oracle.adam_lazy_flush_ref(lazy, lm, lv, last, steps, **HYPER)
    assert np.array_equal(lazy, dense) and np.array_equal(lm, dm) and np.array_equal(lv, dv)
