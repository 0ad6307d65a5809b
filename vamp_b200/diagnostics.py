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
