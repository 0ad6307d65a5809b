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


def test_raftery_lewis_known_chains():
    """vpfits_intro.ipynb cell 18 calls ``mc.raftery_lewis(vpfit.mcmc, q=0.995, r=0.05)`` and prints
    ``(nmin, kthin, nburn, nprec, kmind)`` per variable: nmin = 8 for every one of them."""
    rng = np.random.default_rng(4)
    white = rng.normal(0, 1, 20000)
    nmin, kthin, nburn, nprec, kmind = DG.raftery_lewis(white, q=0.995, r=0.05)
    assert nmin == 8                                         # the notebook's value, a function of (q, r, s) only
    assert kthin == 1 and kmind == 1 and 0 < nburn <= 4 and 0 < nprec <= 12      # notebook: (8, 1, 2, 10, 1) for well-mixed variables
    # the quantile the textbook example uses
    assert DG.raftery_lewis(white, q=0.025, r=0.005)[0] == 3746
    # a two-state indicator chain with KNOWN transition probabilities: x in {0, 1}, cut at the median
    a, b = 0.02, 0.03                                        # P(0 -> 1), P(1 -> 0)
    zc = np.zeros(200000, dtype=np.int64)
    u = rng.random(zc.size)
    for i in range(1, zc.size):
        zc[i] = (1 if u[i] < a else 0) if zc[i - 1] == 0 else (0 if u[i] < b else 1)
    x = 1.0 - zc + 0.01 * rng.random(zc.size)                # indicator (x <= cut) == zc for q = P(z = 1) = a / (a + b)
    q = a / (a + b)
    nmin, kthin, nburn, nprec, kmind = DG.raftery_lewis(x, q=q, r=0.02, s=0.95, epsilon=0.001)
    phi = 1.959963984540054
    assert nmin == int(np.ceil(phi ** 2 * q * (1 - q) / 0.02 ** 2))
    if kthin == 1:                                           # a first-order chain by construction
        want_burn = np.log(0.001 * (a + b) / max(a, b)) / np.log(1 - a - b)
        want_prec = (2 - a - b) * a * b * phi ** 2 / ((a + b) ** 3 * 0.02 ** 2)
        assert abs(nburn - want_burn) < 0.15 * want_burn and abs(nprec - want_prec) < 0.15 * want_prec
    assert kmind > 5                                         # strongly dependent: heavy thinning for independence
    # strongly autocorrelated Gaussian chain: more iterations than under independence
    ar = np.zeros(50000)
    for i in range(1, ar.size):
        ar[i] = 0.95 * ar[i - 1] + rng.normal()
    r = DG.raftery_lewis(ar, q=0.5, r=0.05)
    assert r[3] > 5 * r[0] and r[4] > 1
    # a variable that never moves: PyMC 2 reports INT_MIN (the notebook's xexp_1)
    assert DG.raftery_lewis(np.ones(1000), q=0.995, r=0.05)[2] == -2147483648
