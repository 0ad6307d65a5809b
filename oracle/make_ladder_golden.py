"""Generate ``tests/golden/ladder_oracle.json``: the component-count ladder of the RESTATED reference fit.

TEST INFRASTRUCTURE ONLY.  Runs, on the host cores of the build container, what the reference runs per region -
``VPregion.region_fit`` (vamp_1.0/vpregion.py:42-91) over ``VPfit.find_bic`` (vamp_1.0/vpfits.py:398-429): three
restarts of (PyMC 2 scalar Metropolis, 1 chain, ``tune_interval`` 1000 -> Nelder-Mead MAP from the LAST state -> BIC),
``n += 1`` until the mean BIC stops decreasing - with the oracle's restatements (``oracle/sampler.py``,
``oracle/region.py``; PyMC 2 itself is not installable here, so this is the restated sampler, not PyMC: parity
unpinned, see oracle/__init__.py).  The GPU tests compare the product's chosen n / minimum BIC with these outcomes
(tests/test_ladder_gpu.py) and ``bench.py`` quotes the oracle's ``n == truth`` rate beside the GPU's.

Regions: the first ``N_C4`` regions of BASELINE config C4 (``vamp_b200.synth.config_c4_region``, Voigt, 3000
iterations) and the seven detection regions of the reference's simba fixtures (tests/golden/simba.npz: CII 1036
fitted with Gaussians and 5000 iterations as in simba_spec_demo.ipynb cell 15, HI 1215 with Voigt profiles and the
default 3000 as in cell 25).

    python -m oracle.make_ladder_golden [n_c4] [processes]
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

from oracle import region as OR          # noqa: E402
from oracle import sampler as OS         # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ladder_oracle.json")
N_CAP = 12          # candidates above n0 (the product's ``nmax``)


def _ladder(job):
    name, x, flux, noise, voigt, iterations, seed, truth_n = job
    t0 = time.time()
    n0 = OR.estimate_n(flux)
    freedom = OR.freedom(len(flux), n0)
    rng = np.random.default_rng([int(seed), 20261018])
    tables = {}

    def fb(n):
        b, r, _ = OS.find_bic(x, flux, n, noise, freedom, voigt=voigt, iterations=iterations, rng=rng)
        tables[n] = (b, r)
        return b, r

    # a region of P pixels cannot carry more than (P - 1) // 2 components (the product's cap): stop the ladder there
    cap = min(n0 + N_CAP - 1, max(n0, (len(flux) - 1) // 2))
    kept, exit_n, hist = OR.region_fit_ladder(n0, fb, chi_limit=1.5, n_cap=cap)
    return {"name": name, "P": len(flux), "voigt": bool(voigt), "iterations": int(iterations), "n0": int(n0),
            "freedom": int(freedom), "kept_n": int(kept), "exit_n": int(exit_n), "truth_n": truth_n,
            "history": [[int(a), float(b), float(c)] for a, b, c in hist],
            "bic": {str(n): [float(v) for v in t[0]] for n, t in tables.items()},
            "red_chi": {str(n): [float(v) for v in t[1]] for n, t in tables.items()},
            "seconds": time.time() - t0}


def jobs(n_c4):
    from vamp_b200 import synth
    out = []
    z = np.load(os.path.join(ROOT, "tests", "golden", "simba.npz"))
    for spec, voigt, its in (("simba_CII1036", False, 5000), ("simba_H1215", True, 3000)):
        for r in range(int(z[spec + "/nregions"])):
            k = "%s/r%d/" % (spec, r)
            out.append(("%s/r%d" % (spec, r), z[k + "nu"], z[k + "flux"], z[k + "noise"], voigt, its, 1000 + len(out), None))
    for uid in range(n_c4):
        reg = synth.config_c4_region(uid)
        out.append(("c4/%d" % uid, reg["nu"], reg["flux"], reg["noise"], True, 3000, uid, int(len(reg["truth"]))))
    return out


def main():
    n_c4 = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
    js = jobs(n_c4)
    t0 = time.time()
    res = []
    with Pool(procs) as pool:
        for r in pool.imap_unordered(_ladder, js):
            res.append(r)
            print("%-18s P=%3d n0=%d kept=%d truth=%s  %.0f s" % (r["name"], r["P"], r["n0"], r["kept_n"], r["truth_n"],
                                                                 r["seconds"]), flush=True)
    res.sort(key=lambda r: [j[0] for j in js].index(r["name"]))
    with open(OUT, "w") as fh:
        json.dump({"generator": "oracle/make_ladder_golden.py", "n_c4": n_c4, "regions": res,
                   "wall_seconds": time.time() - t0}, fh, indent=1)
    print("wrote", OUT, "in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
