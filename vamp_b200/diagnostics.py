"""Convergence / model-comparison diagnostics on the sampler's outputs.

The reference's notebooks call PyMC 2's ``geweke`` and ``raftery_lewis`` on the single chain of a
fit and read ``mcmc.DIC`` / ``mcmc.BPIC`` (vpfits_intro.ipynb cells 17-22).  With thousands of chains
per model the natural diagnostic is the multi-chain potential scale reduction R-hat, which needs no
traces: it is formed from the per-chain Welford accumulators K2 keeps (``mean``, ``m2``, ``ntally``).
Host numpy on small arrays; nothing here is on the hot path.
"""
import numpy as np


def rhat(mean, m2, ntally):
    """Gelman-Rubin potential scale reduction per parameter from per-chain accumulators.

    ``mean, m2``: ``[chains, k]`` Welford mean and sum of squared deviations of the tallied samples,
    ``ntally``: ``[chains]`` (all equal for chains advanced together).  Returns ``[k]``; 1 = mixed.
    Parameters that never moved (variance 0 in every chain) give ``nan``."""
    mean, m2 = np.asarray(mean, dtype=np.float64), np.asarray(m2, dtype=np.float64)
    n = float(np.min(ntally))
    m = mean.shape[0]
    if m < 2 or n < 2:
        return np.full(mean.shape[1], np.nan)
    W = np.mean(m2 / (np.asarray(ntally, dtype=np.float64)[:, None] - 1.0), axis=0)      # mean within-chain variance
    B = n * np.var(mean, axis=0, ddof=1)                                                   # between-chain variance
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sqrt(((n - 1.0) / n * W + B / n) / W)


def geweke(x, first=0.1, last=0.5, intervals=20):
    """PyMC 2's ``geweke(x, first, last, intervals)``: z-scores of the mean of an early slice against
    the mean of the last ``last`` fraction, for ``intervals`` start points over the first half.
    Returns ``[intervals, 2]`` = (start index, z)."""
    x = np.asarray(x, dtype=np.float64)
    if not 0.0 < first + last <= 1.0:
        raise ValueError("invalid intervals for the Geweke diagnostic: first=%g last=%g" % (first, last))
    end = len(x) - 1
    zs = []
    for start in np.arange(0, end // 2, max(1, (end // 2) // intervals), dtype=int)[:intervals]:
        a = x[start: start + int(first * (end - start))]
        b = x[int(end - last * (end - start)):]
        z = (a.mean() - b.mean()) / np.sqrt(a.var() + b.var())
        zs.append([start, z])
    return np.array(zs)


def effective_sample_size(x, max_lag=None):
    """Initial-positive-sequence estimate of the effective sample size of one trace."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if n < 4:
        return float(n)
    x = x - x.mean()
    f = np.fft.rfft(x, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n] / (np.arange(n, 0, -1) * x.var())
    max_lag = n // 2 if max_lag is None else max_lag
    s = 0.0
    for k in range(1, max_lag - 1, 2):
        pair = acf[k] + acf[k + 1]
        if pair < 0:
            break
        s += pair
    return n / (1.0 + 2.0 * s)


def dic(sum_lnL, ntally, lnL_at_mean):
    """PyMC 2 ``MCMC.DIC`` = 2 mean(D) - D(theta_bar), deviance D = -2 lnL; ``sum_lnL`` is K2's running sum
    over tallies.  ``BPIC`` = 3 mean(D) - 2 D(theta_bar)."""
    dbar = -2.0 * np.asarray(sum_lnL, dtype=np.float64) / np.asarray(ntally, dtype=np.float64)
    dhat = -2.0 * np.asarray(lnL_at_mean, dtype=np.float64)
    return 2.0 * dbar - dhat, 3.0 * dbar - 2.0 * dhat


# ---------------------------------------------------------------------------------------------------------------
# Raftery-Lewis (vamp_1.0/vpfits_intro.ipynb cell 18: ``mc.raftery_lewis(vpfit.mcmc, q=0.995, r=0.05)``)
def _ppnd(p):
    """Normal quantile (scipy when available; PyMC 2's Fortran uses AS 111 ``ppnd`` - same value to 1e-7)."""
    from scipy.stats import norm
    return float(norm.ppf(p))


def _g2_bic_markov(z):
    """Second- against first-order Markov chain for the binary sequence ``z``: likelihood-ratio statistic G2 and
    ``BIC = G2 - 2 ln(n - 2)`` (``mctest`` of Raftery & Lewis's gibbsit)."""
    z = np.asarray(z, dtype=np.int64)
    n = z.size
    if n < 3:
        return 0.0, -np.inf
    tr = np.bincount(4 * z[:-2] + 2 * z[1:-1] + z[2:], minlength=8).reshape(2, 2, 2).astype(np.float64)
    g2 = 0.0
    for i in range(2):
        for j in range(2):
            for k in range(2):
                if tr[i, j, k] > 0:
                    fitted = (tr[i, j, 0] + tr[i, j, 1]) * (tr[0, j, k] + tr[1, j, k]) / tr[:, j, :].sum()
                    g2 += 2.0 * tr[i, j, k] * np.log(tr[i, j, k] / fitted)
    return g2, g2 - 2.0 * np.log(n - 2.0)


def _g2_bic_indep(z):
    """First-order Markov chain against independence (``indtest`` of gibbsit): ``BIC = G2 - ln(n - 1)``."""
    z = np.asarray(z, dtype=np.int64)
    n = z.size
    if n < 2:
        return 0.0, -np.inf
    tr = np.bincount(2 * z[:-1] + z[1:], minlength=4).reshape(2, 2).astype(np.float64)
    g2 = 0.0
    for i in range(2):
        for j in range(2):
            if tr[i, j] > 0:
                fitted = tr[i, :].sum() * tr[:, j].sum() / (n - 1.0)
                g2 += 2.0 * tr[i, j] * np.log(tr[i, j] / fitted)
    return g2, g2 - np.log(n - 1.0)


def raftery_lewis(x, q, r, s=0.95, epsilon=0.001):
    """PyMC 2's ``raftery_lewis(x, q, r, s=.95, epsilon=.001)`` (its Fortran ``gibbmain`` = Raftery & Lewis 1992,
    "How many iterations in the Gibbs sampler?") for one trace: returns ``(nmin, kthin, nburn, nprec, kmind)`` -
    iterations needed under independence to estimate the ``q`` quantile to within ``r`` with probability ``s``;
    thinning that makes the indicator chain ``x <= x_q`` first-order Markov; burn-in and further iterations at that
    thinning; thinning that makes it an independence chain.  The notebook's settings give ``nmin = 8``."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    phi = _ppnd(0.5 * (1.0 + s))
    nmin = int(np.ceil(phi * phi * q * (1.0 - q) / (r * r)))         # int(1 + ...) in gibbsit: the same unless integral
    cut = np.quantile(x, q)
    z = (x <= cut).astype(np.int64)
    kthin = 1
    while True:                                                       # smallest thinning with a first-order chain
        zt = z[::kthin]
        _, bic = _g2_bic_markov(zt)
        if bic <= 0.0 or zt.size < 6:
            break
        kthin += 1
    tr = np.bincount(2 * zt[:-1] + zt[1:], minlength=4).reshape(2, 2).astype(np.float64)
    alpha = tr[0, 1] / max(tr[0, 0] + tr[0, 1], 1.0)                  # P(0 -> 1)
    beta = tr[1, 0] / max(tr[1, 0] + tr[1, 1], 1.0)                   # P(1 -> 0)
    if alpha <= 0.0 or beta <= 0.0 or alpha + beta >= 2.0 or abs(1.0 - alpha - beta) in (0.0, 1.0):
        # the indicator never changes state (or flips every step): gibbsit's logarithms are undefined - PyMC 2 then
        # reports INT_MIN for nburn / nprec (cell 18 shows it for ``xexp_1``)
        nburn = nprec = -2147483648
    else:
        lam = abs(1.0 - alpha - beta)
        m = np.log(epsilon * (alpha + beta) / max(alpha, beta)) / np.log(lam)
        nburn = int(np.ceil(m)) * kthin
        nn = (2.0 - alpha - beta) * alpha * beta * phi * phi / ((alpha + beta) ** 3 * r * r)
        nprec = int(np.ceil(nn)) * kthin
    kmind = kthin
    while True:                                                       # smallest thinning with an independence chain
        zi = z[::kmind]
        _, bic = _g2_bic_indep(zi)
        if bic <= 0.0 or zi.size < 4:
            break
        kmind += 1
    return nmin, kthin, nburn, nprec, kmind
