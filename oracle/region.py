"""Oracle: the component-count ladder of ``VPregion`` (vamp_1.0/vpregion.py).

TEST INFRASTRUCTURE ONLY.  Restates the control flow of
``VPregion.estimate_n`` (:21-35), ``set_freedom`` (:37-39) and
``region_fit`` (:42-91) over an abstract ``find_bic(n)`` so that the product's
batched ladder replay can be checked against it on the same BIC tables.
"""
import numpy as np
from scipy.ndimage import gaussian_filter
from scipy.signal import argrelextrema


def estimate_n(flux_array):
    """vpregion.py:21-35: number of local minima of the sigma=3 smoothed flux,
    forced to 1 when fewer than 4."""
    n = argrelextrema(gaussian_filter(np.asarray(flux_array, float), 3), np.less)[0].shape[0]
    return 1 if n < 4 else int(n)


def freedom(num_pixels, n):
    """vpregion.py:37-39 (``3 n`` even for Voigt fits)."""
    return num_pixels - 3 * n


def region_fit_ladder(n0, find_bic, chi_limit=1.5, n_cap=64):
    """vpregion.py:54-91.  ``find_bic(n) -> (bic_array, red_chi_array)``.

    Returns ``(n_of_kept_fit, self_n_at_exit, history)`` where ``history`` is the
    list of ``(n, bic_used, mean_red_chi)`` in call order.  Quirks preserved:
    the first pass compares against ``bic_array[-1]`` (not the mean, :64); when
    BIC stops decreasing the kept fit is the previous n but ``self.n`` is only
    decremented under ``verbose`` (:84-87, default verbose=True)."""
    n = n0
    first = True
    history = []
    bic_old = None
    kept_n = None
    while True:
        bic_array, red_chi_array = find_bic(n)
        if first:
            first = False
            bic_old = bic_array[-1]
            kept_n = n
            history.append((n, bic_old, float(np.average(red_chi_array))))
            n += 1
        else:
            cur = float(np.average(bic_array))
            history.append((n, cur, float(np.average(red_chi_array))))
            if bic_old > cur:
                bic_old = cur
                kept_n = n
                if np.average(red_chi_array) < chi_limit:
                    break
                n += 1
            else:
                n -= 1  # verbose=True branch of the reference
                break
        if n > n_cap:
            break
    return kept_n, n, history
