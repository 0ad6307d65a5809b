"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Outputs (committed, small):

* ``tests/golden/profiles.npz``   - seeded inputs and the outputs of the reference's
  ``VPfit.GaussFunction`` / ``VoigtFunction`` (through the wofz ``Voigt1D`` stub of
  ``oracle/load_reference.py``) / ``GaussianWidth`` / ``Chisquared`` /
  ``ReducedChisquared``, ``physics.Tau2flux(sum(...))`` and the vamp_2.0
  ``fit.py`` likelihood, on the BASELINE.json config shapes and edge cases.
* ``tests/golden/physics.npz``    - outputs of every ``physics.py`` function.
* ``tests/golden/wofz_mp.npz``    - w(z) from mpmath at 40 digits (independent of
  scipy) on a z-plane sample covering every branch of the device Faddeeva.
* ``tests/golden/doublet.npz``    - doublet regions (tied second line per component): sums of the reference's
  own profile-function calls, one per line.
* ``tests/golden/simba.npz``      - the two simba fixtures' regions (pixel bounds
  from simba_spec_demo.ipynb) with the reference
  ``VPregion.estimate_n`` / ``set_freedom`` results.
"""
import os

import numpy as np

from . import load_reference as LR
from vamp_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _cases():
    """(name, nu, flux, noise, voigt, comps[n,4], sd) on config shapes + edge cases."""
    cases = []
    rng = np.random.default_rng(20261018)

    def add(name, reg, voigt, comps, sd):
        cases.append((name, reg["nu"], reg["flux"], reg["noise"], voigt, np.asarray(comps, float), float(sd)))

    c1 = synth.config_c1()
    add("c1_truth_voigt", c1, True, c1["truth"], 0.02)
    for n in range(1, 7):
        c2 = synth.config_c2(region_id=100 + n, n_true=n)
        add("c2_truth_n%d" % n, c2, True, c2["truth"], 0.02)
        add("c2_prior_n%d" % n, c2, True, synth.prior_draw_theta(c2["nu"], n, True, rng), rng.uniform(0.01, 1))
        g = synth.make_region(256, n, 50.0, region_id=200 + n, voigt=False)
        add("c2_gauss_truth_n%d" % n, g, False, g["truth"], 0.02)
        add("c2_gauss_prior_n%d" % n, g, False, synth.prior_draw_theta(g["nu"], n, False, rng), rng.uniform(0.01, 1))
    for P in (1, 2, 17, 31, 32, 33, 50, 255, 257, 1000, 4096):
        n = 1 + P % 5
        r = synth.make_region(P, 1, 30.0, region_id=300 + P, voigt=True) if P >= 8 else \
            {"nu": synth.frequency_grid(max(P, 2))[:P] if P > 1 else synth.frequency_grid(2)[:1],
             "flux": np.full(P, 0.7), "noise": np.full(P, 0.05)}
        if P >= 8:
            th = synth.prior_draw_theta(r["nu"], n, True, rng)
        else:
            th = np.array([[1.0, r["nu"][0], 3 * synth.DNU, 5 * synth.DNU]] * n)
        add("ragged_P%d_voigt" % P, r, True, th, 0.05)
    # saturated line, tiny Lorentz width, tiny Gaussian width, edge centroid, huge widths
    base = synth.config_c2(region_id=400, n_true=2)
    nu = base["nu"]
    mid = nu[128]
    D = synth.DNU
    edge = [
        ("saturated", [[60.0, mid, 3 * D, 12 * D], [8.0, nu[60], 0.5 * D, 6 * D]]),
        ("tiny_L", [[1.5, mid, 1e-9 * D, 10 * D]]),
        ("tiny_G", [[1.5, mid, 6 * D, 1e-7 * D]]),
        ("pure_lorentz_like", [[2.0, nu[40], 30 * D, 0.02 * D]]),
        ("edge_centroid", [[1.0, nu[0], 2 * D, 8 * D], [0.7, nu[-1], 1 * D, 5 * D]]),
        ("max_widths", [[0.8, mid, 300.2 * D, 300.2 * D]]),
        ("sub_pixel", [[2.5, mid + 0.37 * D, 0.05 * D, 0.6 * D]]),
        ("zero_amplitude", [[0.0, mid, 2 * D, 8 * D]]),
    ]
    for name, th in edge:
        th = np.asarray(th, float)
        add("edge_" + name, base, True, th, 0.03)
    gedge = [
        ("g_saturated", [[40.0, mid, 9 * D, 0]]),
        ("g_narrow", [[1.0, mid + 0.2 * D, 0.3 * D, 0]]),
        ("g_wide", [[0.5, nu[10], 127 * D, 0]]),
    ]
    for name, th in gedge:
        add("edge_" + name, base, False, np.asarray(th, float), 0.03)
    return cases


def make_profiles(ref):
    out = {}
    names = []
    for name, nu, flux, noise, voigt, comps, sd in _cases():
        names.append(name)
        taus = []
        for a, c, w1, w2 in comps:
            if voigt:
                taus.append(ref.VPfit.VoigtFunction(nu, c, a, w1, w2))      # vpfits.py:58-76
            else:
                taus.append(ref.VPfit.GaussFunction(nu, a, c, w1))          # vpfits.py:44-54
        total = ref.physics.Tau2flux(sum(taus))                              # vpfits.py:336
        chi2 = ref.VPfit.Chisquared(flux, total, noise)                      # vpfits.py:118
        free = max(len(nu) - 3 * len(comps), 1)
        red = ref.VPfit.ReducedChisquared(flux, total, noise, free)          # vpfits.py:131
        fit = ref.fit2.DatasetFit(flux, noise, total, 0)                     # vamp_2.0 fit.py
        out[name + "/nu"] = nu
        out[name + "/flux"] = flux
        out[name + "/noise"] = noise
        out[name + "/voigt"] = np.array(voigt)
        out[name + "/comps"] = comps
        out[name + "/sd"] = np.array(sd)
        out[name + "/taus"] = np.array(taus)
        out[name + "/total"] = total
        out[name + "/chi2"] = np.array(float(chi2))
        out[name + "/red_chi2"] = np.array(float(red))
        out[name + "/freedom"] = np.array(free)
        out[name + "/lnL_noise_map"] = np.array(float(fit.likelihood))
    out["names"] = np.array(names)
    out["gaussian_width_in"] = np.array([0.1, 1.0, 2.3548200450309493, 1e12])
    out["gaussian_width_out"] = ref.VPfit.GaussianWidth(out["gaussian_width_in"])   # vpfits.py:80-88
    np.savez_compressed(os.path.join(OUT, "profiles.npz"), **out)
    return len(names)


def make_physics(ref):
    ph = ref.physics
    amp = np.array([0.3, 1.0, 2.7])
    sig = np.array([1e11, 3e11, 8e11])
    taus = np.linspace(0, 2, 50) ** 2
    edges = np.linspace(1215.0, 1216.0, 50)
    out = dict(
        amp=amp, sig=sig, taus=taus, edges=edges,
        c=np.array(ph.constants["c"]["value"]), sigma0=np.array(ph.constants["sigma0"]["value"]),
        ColumnDensity=ph.ColumnDensity(amp, sig),
        DopplerParameter=ph.DopplerParameter(sig, 1215.67),
        EquivalentWidthTau=np.array(ph.EquivalentWidthTau(taus, edges)),
        EquivalentWidthFlux=np.array(ph.EquivalentWidthFlux(np.exp(-taus), edges)),
        ErrorB=ph.ErrorB(sig * 0.1, 1215.67),
        ErrorN=ph.ErrorN(amp, sig, amp * 0.05, sig * 0.1, np.zeros(3)),
        Errorl=ph.Errorl(sig * 0.01),
        Tau2flux=ph.Tau2flux(taus), Flux2tau=ph.Flux2tau(np.exp(-taus)),
        Freq2wave=ph.Freq2wave(np.array([2.45e15, 2.46e15])),
        Wave2freq=ph.Wave2freq(edges), Wave2red=ph.Wave2red(edges * 3.0, 1215.67),
    )
    np.savez_compressed(os.path.join(OUT, "physics.npz"), **out)


def make_wofz_mp():
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(0, 14, 300), 10 ** rng.uniform(-4, 3, 150), [0.0, 0.0, 5.9999, 6.0001, 12.0, 1e4]])
    ys = np.concatenate([10 ** rng.uniform(-6, 1.3, 300), 10 ** rng.uniform(-6, 3, 150), [1e-8, 6.0, 0.1, 0.1, 0.01, 1e4]])
    re = np.empty(xs.size)
    im = np.empty(xs.size)
    for i, (x, y) in enumerate(zip(xs, ys)):
        z = mp.mpc(x, y)
        w = mp.exp(-z * z) * mp.erfc(-1j * z)
        re[i], im[i] = float(w.real), float(w.imag)
    np.savez_compressed(os.path.join(OUT, "wofz_mp.npz"), x=xs, y=ys, re=re, im=im)


def make_simba(ref):
    out = {}
    bounds = {"simba_CII1036.h5": [[677, 714], [720, 765], [774, 791], [932, 955]],
              "simba_H1215.h5": [[672, 716], [711, 762], [767, 796]]}
    line = {"simba_CII1036.h5": 1036.0, "simba_H1215.h5": 1215.0}
    for fname, regs in bounds.items():
        fx = LR.read_simba_fixture(fname)
        key = fname.split(".")[0]
        nu_all = ref.physics.Wave2freq(fx["wavelength"])       # vpspectrum.py:64
        for j, (s, e) in enumerate(regs):
            # vpspectrum.py:274-279: flipped so that frequency ascends
            flux = np.flip(fx["flux"][s:e], 0)
            noise = np.flip(fx["noise"][s:e], 0)
            nu = np.flip(nu_all[s:e], 0)
            reg = ref.VPregion(nu, flux, noise, voigt=False, chi_limit=1.5)
            p = "%s/r%d/" % (key, j)
            out[p + "nu"], out[p + "flux"], out[p + "noise"] = nu, flux, noise
            out[p + "n0"] = np.array(reg.n)
            out[p + "freedom"] = np.array(reg.freedom)
        out[key + "/nregions"] = np.array(len(regs))
        out[key + "/line"] = np.array(line[fname])
    np.savez_compressed(os.path.join(OUT, "simba.npz"), **out)


def make_detect(ref):
    """Detection regions (vpspectrum.py:67-174) from the reference's own ``compute_detection_regions`` on
    the two simba sightlines and on seeded synthetic sightlines (odd and even lengths, several
    thresholds / widths, ``extend`` on and off)."""
    import contextlib
    import io
    VPspectrum = LR.load_spectrum_class()
    out = {}
    cases = []
    for fname in ("simba_CII1036.h5", "simba_H1215.h5"):
        fx = LR.read_simba_fixture(fname)
        cases.append((fname.split(".")[0], fx["wavelength"], fx["flux"], fx["noise"]))
    rng = np.random.default_rng(20260101)
    for k, P in enumerate((257, 600, 1001, 90)):
        wave = 1215.0 + 0.0203 * np.arange(P)
        tau = np.zeros(P)
        for _ in range(max(1, P // 120)):
            c, s, a = rng.uniform(0.1, 0.9) * P, rng.uniform(2.0, 9.0), rng.uniform(0.3, 3.0)
            tau += a * np.exp(-0.5 * ((np.arange(P) - c) / s) ** 2)
        snr = (20.0, 50.0, 100.0, 30.0)[k]
        noise = np.full(P, 1.0 / snr) * rng.uniform(0.8, 1.2, P)
        flux = np.exp(-tau) + rng.normal(0.0, 1.0, P) * noise
        cases.append(("synth%d" % k, wave, flux, noise))
    settings = [dict(min_region_width=2, N_sigma=4.0, extend=False), dict(min_region_width=2, N_sigma=4.0, extend=True),
                dict(min_region_width=5, N_sigma=6.0, extend=False), dict(min_region_width=2, N_sigma=3.0, extend=False, std_min=3, std_max=8)]
    for name, wave, flux, noise in cases:
        out[name + "/wavelength"], out[name + "/flux"], out[name + "/noise"] = wave, flux, noise
        for j, kw in enumerate(settings):
            sp = VPspectrum(1215.0)
            sp.wavelength_array, sp.flux_array, sp.noise_array = wave.copy(), flux.copy(), noise.copy()
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                sp.compute_detection_regions(**kw)
            out["%s/s%d/region_pixels" % (name, j)] = np.array(sp.region_pixels, dtype=np.int64).reshape(-1, 2)
    out["names"] = np.array([c[0] for c in cases])
    out["settings"] = np.array([repr(sorted(kw.items())) for kw in settings])
    np.savez_compressed(os.path.join(OUT, "detect.npz"), **out)
    return {n: [out["%s/s%d/region_pixels" % (n, j)].tolist() for j in range(len(settings))] for n in out["names"]}


def make_doublets(ref):
    """Doublet regions (BASELINE config C3; a tied second line per component - this repo's extension): the optical
    depths are sums of the REFERENCE'S OWN ``VoigtFunction`` / ``GaussFunction`` calls (vpfits.py:44-76), one per
    line, with the second line's arguments formed as include/vamp_b200.h ``vamp_regions_set_multiplet`` states."""
    out, names = {}, []
    lsf = synth.gaussian_lsf(2.0, 7)
    cases = [("civ_voigt_n2", synth.make_doublet_region(256, 2, 50.0, 501, "CIV", True)),
             ("civ_voigt_n3_lsf", synth.make_doublet_region(256, 3, 80.0, 502, "CIV", True, lsf=lsf)),
             ("mgii_voigt_n1", synth.make_doublet_region(250, 1, 30.0, 503, "MgII", True)),
             ("civ_gauss_n2", synth.make_doublet_region(200, 2, 50.0, 504, "CIV", False)),
             ("mgii_gauss_n2_lsf", synth.make_doublet_region(256, 2, 40.0, 505, "MgII", False, lsf=lsf))]
    rng = np.random.default_rng(77)
    for name, reg in cases:
        voigt, (fr, ar, lr) = reg["voigt"], reg["multiplet"]
        comps = reg["truth"] * (1.0 + rng.uniform(-0.03, 0.03, reg["truth"].shape))
        comps[:, 1] = reg["truth"][:, 1] + rng.uniform(-1, 1, len(comps)) * reg["dnu"]
        nu = reg["nu"]
        taus = []
        for a, c, w1, w2 in comps:
            if voigt:
                t = ref.VPfit.VoigtFunction(nu, c, a, w1, w2) + ref.VPfit.VoigtFunction(nu, fr * c, ar * a, lr * w1, fr * w2)
            else:
                t = ref.VPfit.GaussFunction(nu, a, c, w1) + ref.VPfit.GaussFunction(nu, ar * a, fr * c, fr * w1)
            taus.append(t)
        total = ref.physics.Tau2flux(sum(taus))
        names.append(name)
        for k, v in (("nu", nu), ("flux", reg["flux"]), ("noise", reg["noise"]), ("voigt", np.array(voigt)),
                     ("comps", comps), ("multiplet", np.array([fr, ar, lr])), ("taus", np.array(taus)), ("total", total),
                     ("lsf", lsf if name.endswith("lsf") else np.array([1.0])),
                     ("chi2", np.array(float(ref.VPfit.Chisquared(reg["flux"], total, reg["noise"]))))):
            out[name + "/" + k] = v
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "doublet.npz"), **out)
    return len(names)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = LR.load()
    if "--doublets-only" in __import__("sys").argv:
        print("wrote %d doublet cases" % make_doublets(ref))
        return
    make_doublets(ref)
    n = make_profiles(ref)
    make_physics(ref)
    make_wofz_mp()
    make_simba(ref)
    det = make_detect(ref)
    print("detection regions:", {k: v[0] for k, v in det.items()})
    print("wrote %d profile cases + physics + wofz_mp + simba to %s" % (n, OUT))


if __name__ == "__main__":
    main()
