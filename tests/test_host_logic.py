"""CPU: host-side guards of the product (no GPU needed: every check fires before the device is touched)."""
import types
import warnings

import numpy as np
import pytest

from vamp_b200 import _lib, engine
from vamp_b200 import vpspectrum as VS


def _fake_ctx():
    return types.SimpleNamespace(lib=None, handle=None, torch_device="cpu")


def test_regions_reject_bad_noise_and_flux_before_upload():
    """a zero / NaN / inf noise pixel (weight inf or NaN) used to freeze chains silently and turn BIC / chi^2
    into NaN; it is refused at upload, by ``engine.Regions`` and again by ``vamp_regions_create`` (code -4)."""
    nu = 1e15 + 4e10 * np.arange(40)
    good = (nu, np.ones(40), np.full(40, 0.02))
    for k, bad in ((2, 0.0), (2, np.nan), (2, np.inf), (1, np.nan), (1, np.inf)):
        arrs = [a.copy() for a in good]
        arrs[k][17] = bad
        with pytest.raises(ValueError, match="region 1.*pixel 17"):
            engine.Regions(_fake_ctx(), [good, tuple(arrs)])


def test_sampler_refuses_regions_it_cannot_hold_with_guidance():
    """Regions above 2048 pixels (above 256 for multiplet region sets and the adaptive step method) are refused before
    any device state is allocated, with the way out."""
    regs = types.SimpleNamespace(ctx=_fake_ctx(), npix=np.array([200, 2100], dtype=np.int32), multiplet=None)
    with pytest.raises(_lib.VampError, match="2100 pixels.*split_long_regions"):
        engine.Sampler(regs, np.array([0, 1], np.int32), np.array([1, 1], np.int32), np.arange(2, dtype=np.uint64))
    regs = types.SimpleNamespace(ctx=_fake_ctx(), npix=np.array([200, 300], dtype=np.int32), multiplet=[(1.0, 0.5, 1.0)] * 2)
    with pytest.raises(_lib.VampError, match="300 pixels.*split_long_regions"):
        engine.Sampler(regs, np.array([0, 1], np.int32), np.array([1, 1], np.int32), np.arange(2, dtype=np.uint64))
    regs = types.SimpleNamespace(ctx=_fake_ctx(), npix=np.array([200, 300], dtype=np.int32), multiplet=None)
    with pytest.raises(_lib.VampError, match="300 pixels.*split_long_regions"):
        engine.Sampler(regs, np.array([0, 1], np.int32), np.array([1, 1], np.int32), np.arange(2, dtype=np.uint64),
                       step_method="adaptive")


def test_long_class_of_regions():
    np.testing.assert_array_equal(engine.long_class([1, 256, 257, 512, 513, 1024, 1025, 2048]), [0, 0, 1, 1, 2, 2, 3, 3])


def test_split_long_regions_flags_cuts_through_absorption():
    P = 900
    wave = 1215.0 + 0.02 * np.arange(P)
    flux = np.ones(P)
    flux[100:330] = 0.4                       # region A: 300 px, continuum only at its far end
    flux[450:800] = 0.4                       # region B: saturated trough all the way - no continuum to cut at
    flux[620] = 1.0                           # ... except one pixel in its middle half
    sp = VS.VPspectrum(1215.0).set_spectrum(wave, flux, np.full(P, 0.01))
    sp.region_pixels = [[60, 360], [440, 810]]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sp.split_long_regions()
    assert all(e - s <= 256 for s, e in sp.region_pixels)
    assert [s for s, _ in sp.region_pixels][0] == 60 and sp.region_pixels[-1][1] == 810
    assert 620 in [s for s, _ in sp.region_pixels]            # the continuum pixel is where region B is cut
    assert len(sp.cuts_through_absorption) == len(w) >= 1     # region A had to be cut through absorption
    assert all(flux[c] < 1.0 - 4.0 * 0.01 for c in sp.cuts_through_absorption)
    assert 620 not in sp.cuts_through_absorption


def test_vpfit_std_deviation_is_the_normal_precision():
    """vpfits.py:39: ``std_deviation`` holds 1 / sd^2 (sd free) or 1 / noise^2; vpfits_intro.ipynb cell 37 passes it to
    ``VPfit.Chisquared`` as a number."""
    from vamp_b200.vpfits import VPfit
    f = VPfit(noise=np.full(5, 0.1))
    np.testing.assert_allclose(f.std_deviation.value, 100.0)
    f = VPfit()
    assert np.isnan(f.std_deviation.value)            # no fit yet
    f._sd = 0.05
    assert f.std_deviation.value == pytest.approx(400.0)
    assert VPfit.Chisquared(np.ones(3), np.zeros(3), f.std_deviation) == pytest.approx(3.0 / 400.0 ** 2)


def test_ladder_examined_matches_the_replay():
    """``batch.ladder_examined`` = the last candidate ``ladder_replay`` looks at, whatever lies beyond it in the table."""
    from vamp_b200 import batch
    rng = np.random.default_rng(0)
    for trial in range(300):
        n0, k = int(rng.integers(1, 5)), int(rng.integers(1, 7))
        tables = {n0 + j: (rng.normal(-50 - 3 * j * rng.integers(-1, 3), 4.0, 3), rng.uniform(0.8, 3.0, 3)) for j in range(k)}
        kept, exit_n, decided = batch.ladder_replay(n0, tables, 1.5)
        last = batch.ladder_examined(n0, tables, 1.5)
        assert n0 <= last <= n0 + k - 1
        # cutting the table behind `last` changes nothing; cutting `last` itself leaves the rule undecided (or at n0)
        cut = {n: v for n, v in tables.items() if n <= last}
        assert batch.ladder_replay(n0, cut, 1.5) == (kept, exit_n, decided)
        if decided and last > n0:
            assert batch.ladder_replay(n0, {n: v for n, v in tables.items() if n < last}, 1.5)[2] is False
