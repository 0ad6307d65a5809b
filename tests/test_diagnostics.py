"""CPU: diagnostics on synthetic chains with known behaviour."""
import numpy as np

from vamp_b200 import diagnostics as DG


def _acc(chains):
    mean = chains.mean(axis=1)
    m2 = ((chains - mean[:, None, :]) ** 2).sum(axis=1)
    return mean, m2, np.full(chains.shape[0], chains.shape[1])


def test_rhat_mixed_and_unmixed():
    rng = np.random.default_rng(0)
    good = rng.normal(0, 1, (64, 500, 3))
    r = DG.rhat(*_acc(good))
    assert np.all(np.abs(r - 1.0) < 0.02)
    bad = good.copy()
    bad[:32, :, 1] += 3.0                                  # half of the chains sit in another mode
    r = DG.rhat(*_acc(bad))
    assert abs(r[0] - 1) < 0.02 and r[1] > 1.5 and abs(r[2] - 1) < 0.02
    # the textbook formula on the raw samples
    n = 500.0
    W = bad.var(axis=1, ddof=1).mean(axis=0)
    B = n * bad.mean(axis=1).var(axis=0, ddof=1)
    np.testing.assert_allclose(r, np.sqrt(((n - 1) / n * W + B / n) / W), rtol=1e-12)


def test_geweke_flags_drift():
    rng = np.random.default_rng(1)
    flat = rng.normal(0, 1, 4000)
    z = DG.geweke(flat)
    # PyMC 2 normalises by the slice variances (not the variance of the means): |z| << 1 for a flat chain
    assert z.shape == (20, 2) and np.all(np.abs(z[:, 1]) < 0.3)
    drift = flat + np.linspace(3, 0, 4000)
    assert np.abs(DG.geweke(drift)[0, 1]) > 1.0


def test_ess_and_dic():
    rng = np.random.default_rng(2)
    white = rng.normal(0, 1, 5000)
    assert 0.7 * 5000 < DG.effective_sample_size(white) <= 1.3 * 5000
    ar = np.zeros(5000)
    for i in range(1, 5000):
        ar[i] = 0.9 * ar[i - 1] + rng.normal()
    assert DG.effective_sample_size(ar) < 0.2 * 5000       # (1 - rho) / (1 + rho) = 0.053
    d, b = DG.dic(np.array([-500.0 * 100]), np.array([100]), np.array([-498.0]))
    assert d[0] == 2 * 1000.0 - 996.0 and b[0] == 3 * 1000.0 - 2 * 996.0
