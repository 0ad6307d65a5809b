"""Oracle: float64 replay of the DEVICE sampler (vamp_b200/csrc/vamp_k2.cu) on the CPU.

TEST INFRASTRUCTURE ONLY.  Same Philox4x32-10 stream, same proposal / bounds / accept /
tune / tally rules as the CUDA kernel, but every likelihood comes from the oracle density
(``oracle/density.py``: scipy wofz + numpy), in the kernels' scaled units.  The fp64 CUDA
chains must reproduce this replay step for step (tests/test_k2_gpu.py), which pins the
kernel's control flow, RNG use and cached-tau bookkeeping to an independent implementation;
the *target density* itself is pinned separately against the reference (test_k1_gpu.py).
"""
import math

import numpy as np

from . import density as D

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
FWHM_PER_SIGMA = 2.0 * math.sqrt(2.0 * math.log(2.0))


def philox4x32_10(ctr, key):
    """Salmon et al. 2011 (Random123), 10 rounds."""
    c0, c1, c2, c3 = [int(v) & MASK for v in ctr]
    k0, k1 = [int(v) & MASK for v in key]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(x):
    return (float(x) + 0.5) * 2.3283064365386962890625e-10


class WarpRng:
    """Per-chain stream of (normal, log-uniform) pairs: lane l of block b owns steps l and 32 + l."""

    def __init__(self, uid, seed):
        self.uid_lo, self.uid_hi = uid & MASK, (uid >> 32) & MASK
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.blk = None
        self.vals = None

    def get(self, q):
        blk = q >> 6
        if blk != self.blk:
            self.blk, self.vals = blk, [None] * 64
            for lane in range(32):
                x = philox4x32_10((self.uid_lo, self.uid_hi, blk & MASK, (((blk >> 32) & 0xFFFFFF) << 8) | lane), self.key)
                rad = math.sqrt(-2.0 * math.log(u01(x[0])))
                ang = 2.0 * u01(x[1])
                self.vals[lane] = (rad * math.cos(math.pi * ang), math.log(u01(x[2])))
                self.vals[32 + lane] = (rad * math.sin(math.pi * ang), math.log(u01(x[3])))
        return self.vals[q & 63]


def initial_state(uid, seed, P, n, voigt, sd_free, cmax):
    """``init_prior_kernel``: xexp = 0.5, Uniforms at a prior draw ("INIT" key stream)."""
    key = (seed & MASK, ((seed >> 32) & MASK) ^ 0x494E4954)
    c_hi = float(P - 1)
    w_hi = 0.5 * (P - 1) * (FWHM_PER_SIGMA if voigt else 1.0)
    th = np.zeros(4 * cmax + 1)
    for c in range(n):
        x = philox4x32_10((uid & MASK, (uid >> 32) & MASK, c, 0), key)
        th[4 * c + 0] = 0.5
        th[4 * c + 1] = u01(x[0]) * c_hi
        th[4 * c + 2] = u01(x[1]) * w_hi
        if voigt:
            th[4 * c + 3] = u01(x[2]) * w_hi
    if sd_free:
        x = philox4x32_10((uid & MASK, (uid >> 32) & MASK, 0xFFFFFFFF, 0), key)
        th[4 * cmax] = u01(x[0])
    return th


def _tune_factor(rate):
    rate = np.float32(rate)
    if rate < np.float32(0.001):
        return np.float32(0.1)
    if rate < np.float32(0.05):
        return np.float32(0.5)
    if rate < np.float32(0.2):
        return np.float32(0.9)
    if rate > np.float32(0.95):
        return np.float32(10.0)
    if rate > np.float32(0.75):
        return np.float32(2.0)
    if rate > np.float32(0.5):
        return np.float32(1.1)
    return np.float32(1.0)


def replay_chain(u, flux, noise, theta0, n, voigt, sd_free, uid, seed, iters, burn, thin, tune_interval,
                 tune_throughout, cmax, lsf=None, record_every_step=False, multiplet=None):
    """Run one chain exactly as ``mcmc_chain`` does.  ``u``: scaled pixel grid (0..P-1 for a
    uniform grid).  Returns dict(theta, trace, best_theta, best_logp, mean, m2, ntally, steps)."""
    NP = 4 if voigt else 3
    P = len(u)
    th = np.array(theta0, dtype=np.float64)
    ts = 4 * cmax + 1
    sdi = 4 * cmax
    k = NP * n + (1 if sd_free else 0)
    c_hi = float(P - 1)
    w_hi = 0.5 * (P - 1) * (FWHM_PER_SIGMA if voigt else 1.0)
    psd = np.where(th != 0.0, np.abs(th), 1.0).astype(np.float32)
    asf = np.ones(ts, dtype=np.float32)
    acc = np.zeros(ts, dtype=np.int64)
    rej = np.zeros(ts, dtype=np.int64)
    w = 1.0 / np.asarray(noise, dtype=np.float64)
    lognorm = float(np.sum(np.log(w)) - 0.5 * P * math.log(2.0 * math.pi))

    def S_of(theta):
        comps = theta[: 4 * n].reshape(n, 4)
        m = D.model_flux(u, comps, voigt, lsf, multiplet)      # multiplet in scaled units: (fr, ar, lr, shift)
        r = flux - m
        return float(np.sum((r * w) ** 2)) if not sd_free else float(np.sum(r * r))

    def lnL_of(S, sd):
        if sd_free:
            return -0.5 * S / (sd * sd) - P * (math.log(sd) + 0.91893853320467274178)
        return -0.5 * S + lognorm

    rng = WarpRng(uid, seed)
    Scur = S_of(th)
    sd = th[sdi] if sd_free else 1.0
    best_logp, best = -np.inf, th.copy()
    mean, m2 = np.zeros(ts), np.zeros(ts)
    ntally = 0
    trace, steps = [], []
    q = 0
    for it in range(iters):
        if it > 0 and it % tune_interval == 0 and (tune_throughout or it <= burn):
            for i in range(ts):
                if acc[i] + rej[i] > 0:
                    nf = np.float32(asf[i] * _tune_factor(np.float32(acc[i]) / np.float32(acc[i] + rej[i])))
                    if nf > 0:
                        asf[i] = nf
                acc[i] = rej[i] = 0
        for c in range(n):
            for i in range(NP):
                idx = 4 * c + i
                z, lu = rng.get(q)
                q += 1
                cur = th[idx]
                prop = cur + z * float(np.float32(asf[idx] * psd[idx]))
                if i == 0:
                    ok = prop > 0.0
                elif i == 1:
                    ok = 0.0 <= prop <= c_hi
                elif voigt and i == 2:
                    ok = 0.0 <= prop <= w_hi
                else:
                    ok = 0.0 < prop <= w_hi
                accept = False
                if ok:
                    t2 = th.copy()
                    t2[idx] = prop
                    Snew = S_of(t2)
                    dprior = (math.log(prop) - prop) - (math.log(cur) - cur) if i == 0 else 0.0
                    dl = -0.5 * (Snew - Scur) / (sd * sd) if sd_free else -0.5 * (Snew - Scur)
                    accept = lu < dl + dprior
                    if accept:
                        th[idx] = prop
                        Scur = Snew
                if accept:
                    acc[idx] += 1
                else:
                    rej[idx] += 1
                if record_every_step:
                    steps.append((idx, prop, accept, ok))
        if sd_free:
            z, lu = rng.get(q)
            q += 1
            prop = sd + z * float(np.float32(asf[sdi] * psd[sdi]))
            accept = False
            if 0.0 < prop <= 1.0:
                d = -0.5 * Scur * (1.0 / (prop * prop) - 1.0 / (sd * sd)) - P * (math.log(prop) - math.log(sd))
                accept = lu < d
            if accept:
                sd = prop
                th[sdi] = prop
                acc[sdi] += 1
            else:
                rej[sdi] += 1
            if record_every_step:
                steps.append((sdi, prop, accept, 0.0 < prop <= 1.0))
        lnL = lnL_of(Scur, sd)
        logp = lnL + sum(math.log(th[4 * c]) - th[4 * c] for c in range(n))
        if logp > best_logp:
            best_logp, best = logp, th.copy()
        if it >= burn and it % thin == 0:
            ntally += 1
            d = th - mean
            mean = mean + d / ntally
            m2 = m2 + d * (th - mean)
            trace.append(th.copy())
    return dict(theta=th, trace=np.array(trace), best_theta=best, best_logp=best_logp, mean=mean, m2=m2,
                ntally=ntally, steps=steps, asf=asf, lnL=lnL_of(Scur, sd))


# ---------------------------------------------------------------------------------------------------------------------
# Adaptive Metropolis (vamp_k2.cu:mcmc_chain_am; step_kind = VAMP_STEP_ADAPTIVE)
def _tri(i, j):
    return i * (i + 1) // 2 + j


def am_initial_cov(theta_c, NP):
    """PyMC 2 ``AdaptiveMetropolis.cov_from_value(100)`` [3P-memory]: C_0 = diag(|value| / 100), 1 for a zero value."""
    v = np.asarray(theta_c[:NP], dtype=np.float64)
    return np.diag(np.where(v != 0.0, np.abs(v), 1.0) / 100.0)


def replay_chain_am(u, flux, noise, theta0, n, voigt, sd_free, uid, seed, iters, burn, thin, tune_interval, tune_throughout,
                    cmax, delay=1000, interval=200, lsf=None):
    """Float64 replay of the device's adaptive-Metropolis chain: PyMC 2's ``AdaptiveMetropolis`` (Haario et al. 2001)
    installed per component - one joint proposal ``theta_c + L z`` per iteration and component, greedy tallying of the
    block history until ``delay``, recursive covariance update every ``interval`` iterations after it (scaling
    ``2.4^2 / d``, ``epsilon = 1e-5``), Cholesky refactorisation - with the scalar Metropolis + ``tune()`` step on
    ``sd``.  Same Philox stream as the device: a block consumes the ``NP`` draws of its parameters, using their
    normals and the first draw's log-uniform.  [3P-memory: PyMC 2.3 ``StepMethods.AdaptiveMetropolis``; the two stated
    deviations - previous covariance weighted by ``max(k - 1, 0)``, a non positive definite update keeps the old
    factor - are the device's.]"""
    NP = 4 if voigt else 3
    P = len(u)
    th = np.array(theta0, dtype=np.float64)
    ts, sdi = 4 * cmax + 1, 4 * cmax
    c_hi = float(P - 1)
    w_hi = 0.5 * (P - 1) * (FWHM_PER_SIGMA if voigt else 1.0)
    psd_sd = np.float32(abs(th[sdi]) if th[sdi] != 0.0 else 1.0)
    asf_sd = np.float32(1.0)
    acc_sd = tot_sd = 0
    w = 1.0 / np.asarray(noise, dtype=np.float64)
    lognorm = float(np.sum(np.log(w)) - 0.5 * P * math.log(2.0 * math.pi))

    def S_of(theta):
        m = D.model_flux(u, theta[: 4 * n].reshape(n, 4), voigt, lsf)
        r = flux - m
        return float(np.sum((r * w) ** 2)) if not sd_free else float(np.sum(r * r))

    def lnL_of(S, sd):
        if sd_free:
            return -0.5 * S / (sd * sd) - P * (math.log(sd) + 0.91893853320467274178)
        return -0.5 * S + lognorm

    C = [am_initial_cov(th[4 * c: 4 * c + 4], NP) for c in range(n)]
    L = [np.sqrt(Ci) for Ci in C]                      # diagonal: its own Cholesky factor
    cmean = [np.zeros(NP) for _ in range(n)]
    sumx = [np.zeros(NP) for _ in range(n)]
    sumxx = [np.zeros((NP, NP)) for _ in range(n)]
    kcount = [0.0] * n
    mcount = [0.0] * n
    acc_b = np.zeros(n, dtype=np.int64)
    rng = WarpRng(uid, seed)
    Scur = S_of(th)
    sd = th[sdi] if sd_free else 1.0
    best_logp, best = -np.inf, th.copy()
    mean, m2 = np.zeros(ts), np.zeros(ts)
    ntally = 0
    trace = []
    q = 0
    for it in range(iters):
        if sd_free and it > 0 and it % tune_interval == 0 and (tune_throughout or it <= burn):
            if tot_sd > 0:
                nf = np.float32(asf_sd * _tune_factor(np.float32(acc_sd) / np.float32(tot_sd)))
                if nf > 0:
                    asf_sd = nf
            acc_sd = tot_sd = 0
        for c in range(n):
            z = np.zeros(NP)
            lu = 0.0
            for i in range(NP):
                zi, lui = rng.get(q)
                q += 1
                z[i] = zi
                if i == 0:
                    lu = lui
            cur = th[4 * c: 4 * c + NP].copy()
            prop = cur + L[c] @ z
            ok = prop[0] > 0.0 and 0.0 <= prop[1] <= c_hi
            if voigt:
                ok = ok and 0.0 <= prop[2] <= w_hi and 0.0 < prop[3] <= w_hi
            else:
                ok = ok and 0.0 < prop[2] <= w_hi
            accept = False
            if ok:
                t2 = th.copy()
                t2[4 * c: 4 * c + NP] = prop
                Snew = S_of(t2)
                dprior = (math.log(prop[0]) - prop[0]) - (math.log(cur[0]) - cur[0])
                dl = -0.5 * (Snew - Scur) / (sd * sd) if sd_free else -0.5 * (Snew - Scur)
                accept = lu < dl + dprior
                if accept:
                    th = t2
                    Scur = Snew
                    acc_b[c] += 1
            if accept or it >= delay:
                x = th[4 * c: 4 * c + NP]
                sumx[c] += x
                sumxx[c] += np.outer(x, x)
                mcount[c] += 1.0
        if sd_free:
            zi, lui = rng.get(q)
            q += 1
            prop = sd + zi * float(np.float32(asf_sd * psd_sd))
            if 0.0 < prop <= 1.0:
                d = -0.5 * Scur * (1.0 / (prop * prop) - 1.0 / (sd * sd)) - P * (math.log(prop) - math.log(sd))
                if lui < d:
                    sd = prop
                    th[sdi] = prop
                    acc_sd += 1
            tot_sd += 1
        if it > delay and it % interval == 0:
            for c in range(n):
                kk, m = kcount[c], mcount[c]
                nn = kk + m
                if m > 0 and nn > 1:
                    nm = (kk * cmean[c] + sumx[c]) / nn
                    t = kk * np.outer(cmean[c], cmean[c]) + sumxx[c] - nn * np.outer(nm, nm) + 1.0e-5 * np.eye(NP)
                    C[c] = max(kk - 1.0, 0.0) / (nn - 1.0) * C[c] + (5.76 / NP) / (nn - 1.0) * t
                    cmean[c], kcount[c], mcount[c] = nm, nn, 0.0
                    sumx[c] = np.zeros(NP)
                    sumxx[c] = np.zeros((NP, NP))
                    try:
                        L[c] = np.linalg.cholesky(C[c])
                    except np.linalg.LinAlgError:
                        pass
        lnL = lnL_of(Scur, sd)
        logp = lnL + sum(math.log(th[4 * c]) - th[4 * c] for c in range(n))
        if logp > best_logp:
            best_logp, best = logp, th.copy()
        if it >= burn and it % thin == 0:
            ntally += 1
            d = th - mean
            mean = mean + d / ntally
            m2 = m2 + d * (th - mean)
            trace.append(th.copy())
    return dict(theta=th, trace=np.array(trace), best_theta=best, best_logp=best_logp, mean=mean, m2=m2, ntally=ntally,
                lnL=lnL_of(Scur, sd), cov=C, chol=L, chain_mean=cmean, accepted=acc_b)
