"""GPU: OUTCOME of the component-count ladder against the restated reference fit.

``tests/golden/ladder_oracle.json`` (generator: oracle/make_ladder_golden.py, run in the build container) holds, for
the first 32 regions of BASELINE config C4 and for the seven detection regions of the reference's simba fixtures,
what the restated reference does: ``VPregion.region_fit`` (vamp_1.0/vpregion.py:42-91) over ``VPfit.find_bic``
(vamp_1.0/vpfits.py:398-429) - 3 restarts of (1 chain of PyMC 2 scalar Metropolis, tune interval 1000, Nelder-Mead
from the last state).  The notebook anchors (vamp_1.0/simba_spec_demo.ipynb cells 15 and 25: CII 1036 -> n = 5, 4,
3, 1 with BIC ladder -34.29 -> -56.00 -> ...; HI 1215 -> n = 2, 1, 1) are the reference's own recorded outcomes.

The GPU fit deliberately samples better than the reference (many chains per restart, polish from the best state,
DESIGN.md section 6), so per candidate its BIC is at or below the oracle's; the ladder rule then is the same rule
on better tables.  What is asserted: agreement of the chosen n with the oracle (rates), minimum BIC not above the
oracle's, the notebook anchors within +-1, the n = 1 BIC values of the notebook reproduced, and - in the
reference-default mode (1 chain, tune 1000, polish from the last state) - the same agreement like for like.
"""
import json
import os

import numpy as np
import pytest

from vamp_b200 import synth
from vamp_b200.fit_regions import REFERENCE_DEFAULTS, fit_regions

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ladder_oracle.json")))
SIMBA = np.load(os.path.join(os.path.dirname(__file__), "golden", "simba.npz"))
C4_KW = dict(voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=8, window=3, nmax=8)
# vamp_1.0/simba_spec_demo.ipynb cell 15 (CII 1036, Gaussian, 5000 iterations) and cell 25 (HI 1215, Voigt):
NOTEBOOK_N = {"simba_CII1036": [5, 4, 3, 1], "simba_H1215": [2, 1, 1]}
NOTEBOOK_BIC_N1 = {"simba_CII1036/r0": -34.29, "simba_CII1036/r1": -43.64, "simba_CII1036/r2": -13.30,
                   "simba_H1215/r0": -0.79}
NOTEBOOK_LADDER_R0 = [-34.29, -56.00, -104.15, -156.39, -156.81]          # CII region 0, n = 1 .. 5


def _gold(prefix):
    return [r for r in GOLD["regions"] if r["name"].startswith(prefix)]


def _min_bic(tables):
    return min(float(np.min(v)) for v in tables.values())


def _report(name, payload):
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "ladder_%s.json" % name), "w") as fh:
            json.dump(payload, fh, indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("mode", ["default", "reference"])
def test_c4_ladder_outcome_vs_oracle(ctx, mode):
    gold = _gold("c4/")
    regs = [synth.config_c4_region(int(g["name"].split("/")[1])) for g in gold]
    host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
    kw = dict(C4_KW)
    if mode == "reference":
        kw.update(REFERENCE_DEFAULTS)
    res = fit_regions(host, uids=[int(g["name"].split("/")[1]) for g in gold], seed=11, distributed=False, **kw)
    n_gpu = np.array([r.n for r in res])
    n_orc = np.array([g["kept_n"] for g in gold])
    truth = np.array([g["truth_n"] for g in gold])
    assert [r.n0 for r in res] == [g["n0"] for g in gold]                     # estimate_n: identical by construction
    exact, within1 = float(np.mean(n_gpu == n_orc)), float(np.mean(np.abs(n_gpu - n_orc) <= 1))
    bic_gpu = np.array([_min_bic(r.bic) for r in res])
    bic_orc = np.array([_min_bic(g["bic"]) for g in gold])
    payload = {"mode": mode, "regions": len(gold), "n_gpu": n_gpu.tolist(), "n_oracle": n_orc.tolist(), "truth": truth.tolist(),
               "agree_exact": exact, "agree_within_1": within1,
               "gpu_n_equals_truth": float(np.mean(n_gpu == truth)), "oracle_n_equals_truth": float(np.mean(n_orc == truth)),
               "gpu_n_within_1_of_truth": float(np.mean(np.abs(n_gpu - truth) <= 1)),
               "oracle_n_within_1_of_truth": float(np.mean(np.abs(n_orc - truth) <= 1)),
               "min_bic_gpu_minus_oracle": (bic_gpu - bic_orc).tolist()}
    _report("c4_" + mode, payload)
    print(json.dumps({k: v for k, v in payload.items() if not isinstance(v, list)}))
    # two stochastic fits of the same ladder: most regions end on the same count, nearly all within one
    assert exact >= 0.4 and within1 >= 0.7, payload
    # against the truth the GPU ladder is not worse than the restated reference's
    assert payload["gpu_n_within_1_of_truth"] >= payload["oracle_n_within_1_of_truth"] - 0.1
    if mode == "default":
        # more chains and polish from the best state: the best BIC found is at or below the reference-style one
        # (a few BIC units of slack: the ladders may stop at different n)
        assert np.mean(bic_gpu <= bic_orc + 3.0) >= 0.85, payload["min_bic_gpu_minus_oracle"]


@pytest.mark.parametrize("mode", ["reference", "default"])
@pytest.mark.parametrize("spec,voigt,iters", [("simba_CII1036", False, 5000), ("simba_H1215", True, 3000)])
def test_simba_fixtures_reach_the_notebook_outcome(ctx, spec, voigt, iters, mode):
    """The reference's own fixtures and recorded outcomes.  ``reference`` mode (1 chain per restart, PyMC 2's tune
    interval, polish from the last state - what the notebook ran): final n within one of the notebook's.  ``default``
    mode (32 chains per restart, polish from the best state): per candidate the BIC is at or below the restated
    reference's, and - because the ladder's rule then sees lower BICs for the larger models (HI region 0, n = 3: -35
    against the reference-style +6) - the chosen n is the notebook's or larger."""
    gold = _gold(spec)
    host, n0 = [], []
    for j in range(int(SIMBA[spec + "/nregions"])):
        k = "%s/r%d/" % (spec, j)
        host.append((SIMBA[k + "nu"], SIMBA[k + "flux"], SIMBA[k + "noise"]))
        n0.append(int(SIMBA[k + "n0"]))
    kw = dict(voigt=voigt, iterations=iters, burn=300, thin=15, restarts=3, chains_per_restart=32, window=3, nmax=8)
    if mode == "reference":
        kw.update(REFERENCE_DEFAULTS)
    res = fit_regions(host, seed=5, distributed=False, **kw)
    n_gpu = [r.n for r in res]
    want = NOTEBOOK_N[spec]
    _report("%s_%s" % (spec, mode), {"n_gpu": n_gpu, "notebook": want, "oracle": [g["kept_n"] for g in gold],
                                     "bic": [{str(n): [float(v) for v in b] for n, b in r.bic.items()} for r in res]})
    assert [r.n0 for r in res] == n0
    if mode == "reference":
        assert all(abs(a - b) <= 1 for a, b in zip(n_gpu, want)), (n_gpu, want)
        return
    assert all(a >= b - 1 for a, b in zip(n_gpu, want)), (n_gpu, want)
    pairs = []
    for j, (r, g) in enumerate(zip(res, gold)):
        name = "%s/r%d" % (spec, j)
        # BIC of the 1-component fit, an (almost) deterministic function of the data: the GPU fit (K2 + K3) and the
        # restated reference fit (numpy / scipy sampler + fmin) land on the SAME optimum ...
        assert float(np.min(r.bic[1])) == pytest.approx(min(g["bic"]["1"]), abs=0.02), (name, r.bic[1], g["bic"]["1"])
        if name in NOTEBOOK_BIC_N1:
            # ... which is the value the reference's notebook printed, to within the 1.5 (Gauss) / 3 (Voigt) BIC units by
            # which that recorded run (unknown PyMC / numpy / astropy versions) differs from BOTH
            assert abs(float(np.min(r.bic[1])) - NOTEBOOK_BIC_N1[name]) < (3.0 if voigt else 1.5), (name, r.bic[1])
        # every candidate both fits tried: the GPU's best restart against the reference-style MEAN over restarts (the
        # quantity the ladder compares) and against its best restart
        for n_s, b_ref in g["bic"].items():
            if int(n_s) in r.bic:
                pairs.append((name, int(n_s), float(np.min(r.bic[int(n_s)])), float(np.mean(b_ref)), float(np.min(b_ref))))
    # many chains + polish from the best state: never worse than the reference-style mean over restarts, and at or
    # below even its best restart for most candidates (a single chain's last state occasionally falls into a deeper,
    # narrower optimum than the mode the pooled chains sit on: CII region 2, n = 3: -83.2 against -80.7)
    assert all(gpu <= ref_mean + 1.0 for _, _, gpu, ref_mean, _ in pairs), pairs
    assert np.mean([gpu <= ref_min + 1.0 for _, _, gpu, _, ref_min in pairs]) >= 0.7, pairs
    if spec == "simba_CII1036":
        # the reference's recorded ladder of region 0 (n = 1 .. 5): the GPU's per-candidate best BIC is at or below it
        for n, ref in zip(range(1, 6), NOTEBOOK_LADDER_R0):
            if n in res[0].bic:
                assert float(np.min(res[0].bic[n])) <= ref + 1.0, (n, res[0].bic[n], ref)


def test_records_do_not_depend_on_world_size_or_batching(ctx):
    """The records of a W-rank run (``emulate=(rank, W)``: exactly what rank ``rank`` would compute on its LPT shard),
    put together, are BIT-identical to the single-process records, for W = 2, 4, and for a run cut into small device
    batches: Philox streams are keyed by region uid and every schedule is chosen from global quantities."""
    uids = list(range(300, 324))
    regs = [synth.config_c4_region(u) for u in uids]
    host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
    kw = dict(C4_KW)
    kw.update(iterations=900, burn=150)
    one = fit_regions(host, uids=uids, seed=3, distributed=False, **kw)
    for world in (2, 4):
        parts = [fit_regions(host, uids=uids, seed=3, emulate=(r, world), **kw).records for r in range(world)]
        assert sum(len(p) for p in parts) == len(uids) and all(len(p) for p in parts)
        allr = np.concatenate(parts)
        allr = allr[np.argsort(allr[:, 0].astype(np.int64), kind="stable")]
        np.testing.assert_array_equal(allr, one.records)
    many = fit_regions(host, uids=uids, seed=3, distributed=False, max_batch_regions=5, **kw)
    assert many.checksum() == one.checksum()
    # candidates fitted ahead of the rule in thin late rounds (never / always): left out of the records
    for spec in (0, 10 ** 9):
        assert fit_regions(host, uids=uids, seed=3, distributed=False, spec_below=spec, **kw).checksum() == one.checksum()
    assert fit_regions(host, uids=uids, seed=4, distributed=False, **kw).checksum() != one.checksum()
