"""Generate ``tests/golden/posterior_c2.json``: reference-style posterior of N, b, v on the BASELINE C2 region.

TEST INFRASTRUCTURE ONLY.  north_star: "sampler output is validated statistically against the reference's PyMC
chains (posterior means and credible intervals of N, b and v within Monte Carlo error on synthetic absorbers with
known truth)".  PyMC 2 cannot be installed here, so the chains come from the oracle's restatement of its scalar
Metropolis sampler (``oracle/sampler.py:metropolis_sample``, PyMC 2.3 ``tune()`` every 1000 iterations) over the
reference density (``oracle/density.py``: the reference's own lines + scipy wofz): ``CHAINS`` independent chains of
the 3-component Voigt model on ``vamp_b200.synth.config_c2()`` (256 px, S/N 50), started at the truth.  N, b and v
are derived per sample the way the reference derives them after a fit (vamp_1.0/vpspectrum.py:377, 421-422:
``sigma = GaussianWidth(G_fwhm)``, ``ColumnDensity(amplitude, sigma)``, ``DopplerParameter(sigma, line)``;
``Freq2wave(centroid)`` -> velocity offset from the line).  Components are ordered by centroid per sample.

    python -m oracle.make_posterior_golden [chains] [iterations]
"""
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sampler as OS          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "posterior_c2.json")
LINE = 1215.67
BURN, THIN = 20000, 10       # the chains start at the truth, 5 sigma from the bulk of the amplitude_L posterior: long burn-in


def nbv(comps):
    """``comps [T, n, 4]`` (amplitude, centroid, L_fwhm, G_fwhm) in Hz -> ``[T, n, 3]`` (N, b, v), components ordered
    by centroid.  Uses the reference's physics.py formulas (restated in vamp_b200/physics.py, pinned by
    tests/golden/physics.npz)."""
    from vamp_b200 import physics
    comps = np.asarray(comps, dtype=np.float64)
    order = np.argsort(comps[:, :, 1], axis=1)
    comps = np.take_along_axis(comps, order[:, :, None], axis=1)
    sigma = comps[:, :, 3] / (2.0 * np.sqrt(2.0 * np.log(2.0)))               # VPfit.GaussianWidth
    N = physics.ColumnDensity(comps[:, :, 0], sigma)
    b = physics.DopplerParameter(sigma, LINE)
    v = physics.constants["c"]["value"] * (physics.Freq2wave(comps[:, :, 1]) - LINE) / LINE * 1e-3
    return np.stack([N, b, v], axis=2)


def _chain(job):
    seed, iters = job
    from vamp_b200 import synth
    reg = synth.config_c2()
    model = OS.CachedModel(reg["nu"], reg["flux"], 3, True)
    t = reg["truth"]
    start = np.concatenate([t.reshape(-1), [1.0 / reg["snr"]]])
    psd = np.concatenate([np.tile([0.05, 0.05 * synth.DNU, 0.5 * synth.DNU, 0.3 * synth.DNU], 3) *
                          np.array([t[c // 4, 0] if c % 4 == 0 else 1.0 for c in range(12)]), [0.001]])
    rng = np.random.default_rng([seed, 4242])
    trace, _, _ = OS.metropolis_sample(model, start, iters, BURN, THIN, rng, tune_interval=1000, proposal_sd=psd)
    return nbv(trace[:, :12].reshape(len(trace), 3, 4)), trace[:, 12]


def main():
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    t0 = time.time()
    with Pool(min(chains, os.cpu_count() or 4)) as pool:
        res = pool.map(_chain, [(s, iters) for s in range(chains)])
    q = np.stack([r[0] for r in res])                    # [chains, T, 3 comps, 3]
    sd = np.stack([r[1] for r in res])
    pooled = q.reshape(-1, 3, 3)
    from vamp_b200 import synth
    reg = synth.config_c2()
    truth = nbv(reg["truth"][None])[0]
    out = {"generator": "oracle/make_posterior_golden.py", "chains": chains, "iterations": iters, "burn": BURN,
           "thin": THIN, "samples_per_chain": int(q.shape[1]), "quantities": ["N", "b", "v"], "line": LINE,
           "truth": truth.tolist(),
           "mean": pooled.mean(axis=0).tolist(), "std": pooled.std(axis=0).tolist(),
           "q16": np.percentile(pooled, 16, axis=0).tolist(), "q50": np.percentile(pooled, 50, axis=0).tolist(),
           "q84": np.percentile(pooled, 84, axis=0).tolist(),
           "chain_means": q.mean(axis=1).tolist(),
           "mc_error": (q.mean(axis=1).std(axis=0, ddof=1) / np.sqrt(chains)).tolist(),
           "sd_mean": float(sd.mean()), "sd_std": float(sd.std()), "wall_seconds": time.time() - t0}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", OUT, "in %.0f s" % (time.time() - t0))
    for c in range(3):
        print("comp", c, "truth", truth[c], "mean", np.array(out["mean"][c]), "std", np.array(out["std"][c]),
              "mc", np.array(out["mc_error"][c]))


if __name__ == "__main__":
    main()
