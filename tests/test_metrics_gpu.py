"""GPU evaluation metrics against the reference's own get_R / fold-loop lines (tests/golden/
metrics.npz) and the oracle."""
import numpy as np
import pytest
import torch

from mclstexp_b200 import metrics, synth
from oracle import oracle
from conftest import _load, golden_checksum

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_metrics_vs_reference_golden(name, dtype):
    z, meta = _load("metrics.npz")
    m = meta[name]
    true = synth.expression(m["Q"], m["G"], m["seed"]).astype(np.float64)
    pred = z[f"{name}/pred"].astype(np.float64)
    assert golden_checksum(true, pred) == m["checksum"]
    got = metrics.evaluate(true.astype(dtype), pred.astype(dtype))
    for key in ("heg_pcc", "hvg_pcc", "mse", "mae"):
        np.testing.assert_allclose(got[key], float(z[f"{name}/{key}"]), rtol=1e-5, err_msg=key)


def test_per_gene_values_and_nan():
    true = synth.expression(4000, 1000, 5)
    rng = np.random.default_rng(1)
    pred = (0.5 * true + 0.5 * rng.random(true.shape)).astype(np.float32)
    pred[:, 17] = 1.0
    mean_true, pcc, sq, ab = metrics.gene_metrics_device(torch.tensor(true).cuda(), torch.tensor(pred).cuda())
    ref = oracle.metrics_ref(true.astype(np.float64), pred.astype(np.float64))
    got = pcc.cpu().numpy()
    assert np.isnan(got[17]) and np.isnan(ref["pcc"][17])
    ok = ~np.isnan(ref["pcc"])
    np.testing.assert_allclose(got[ok], ref["pcc"][ok], rtol=0, atol=1e-9)
    np.testing.assert_allclose(mean_true.cpu().numpy(), true.astype(np.float64).mean(0), rtol=1e-12)
    np.testing.assert_allclose(float(sq.sum()) / true.size, ref["mse"], rtol=1e-10)
    np.testing.assert_allclose(float(ab.sum()) / true.size, ref["mae"], rtol=1e-10)
