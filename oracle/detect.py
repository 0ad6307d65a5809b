"""Oracle: detection regions of a sightline (TEST INFRASTRUCTURE ONLY).

Restates ``VPspectrum.compute_detection_regions`` (vamp_1.0/vpspectrum.py:67-174): equivalent
widths of the flux decrement and of the noise per pixel, matched filter with Gaussians of
std = std_min .. std_max-1 pixels, detection ratio = best flux_func / sqrt(noise_func), then the
reference's sequential region selection.  Pinned against the reference's own method on the simba
sightlines and synthetic spectra (tests/golden/detect.npz, oracle/make_golden.py:make_detect).
"""
import numpy as np


def equivalent_widths(wavelength, flux, noise):
    """vpspectrum.py:92-103: interior pixels only; a decrement below the noise counts as zero."""
    P = len(wavelength)
    few, newz = np.zeros(P), np.zeros(P)
    dl = 0.5 * np.abs(wavelength[:-2] - wavelength[2:])
    dec = 1.0 - flux[1:-1]
    dec = np.where(dec < noise[1:-1], 0.0, dec)
    few[1:-1] = dl * dec
    newz[1:-1] = dl * noise[1:-1]
    return few, newz


def detection_ratio(wavelength, flux, noise, std_min=2, std_max=11):
    """vpspectrum.py:105-122.  ``np.convolve(a, g, 'same')`` with a kernel as long as the signal:
    out[i] = sum_j a[j] g[i - j + (P-1)//2], g[m] = exp(-(m - (P-1)/2)^2 / 2 std^2) - for even P the
    Gaussian sits half a pixel off centre (a quirk of the reference, kept)."""
    wavelength, flux, noise = (np.asarray(a, dtype=np.float64) for a in (wavelength, flux, noise))
    P = len(wavelength)
    few, newz = equivalent_widths(wavelength, flux, noise)
    xarr = np.arange(P) - (P - 1) / 2.0
    det = np.full(P, -np.inf)
    for std in range(std_min, std_max):
        g = np.exp(-0.5 * (xarr / std) ** 2)            # VPfit.GaussFunction(xarr, 1, 0, std), vpfits.py:44-54
        ff = np.convolve(few, g, "same")
        nf = np.convolve(newz ** 2, g ** 2, "same")
        with np.errstate(divide="ignore", invalid="ignore"):
            r = ff[1:-1] / np.sqrt(nf[1:-1])
        better = r > det[1:-1]                            # NaN compares false, as in the reference
        det[1:-1] = np.where(better, r, det[1:-1])
    return det


def select_regions(det_ratio, wavelength, flux, noise, min_region_width=2, N_sigma=4.0, extend=False):
    """vpspectrum.py:124-174 (sequential by construction)."""
    P = len(flux)
    start, endpoints = 0, []
    for i in range(P):
        if start == 0 and det_ratio[i] > N_sigma and flux[i] < 1.0:
            start = i
        elif start != 0 and (det_ratio[i] < N_sigma or flux[i] > 1.0):
            if (i - start) > min_region_width:
                endpoints.append([start, i])
            start = 0
    if extend:
        expanded = []
        for s, e in endpoints:
            i = s
            while i > 0 and flux[i] < 1.0:
                i -= 1
            j = e
            while j < P - 1 and flux[j] < 1.0:
                j += 1
            expanded.append([i, j])
    else:
        expanded = endpoints
    pixels, buffer = [], 3
    for i, (s, e) in enumerate(expanded):
        if i < len(expanded) - 1 and e > expanded[i + 1][0]:
            e = expanded[i + 1][1]
        for j in range(s, e):
            if 1.0 - flux[j] > abs(noise[j]) * N_sigma:
                if s >= buffer:
                    s -= buffer
                if e < P - buffer:
                    e += buffer
                pixels.append([s, e])
                break
    return pixels


def detection_regions(wavelength, flux, noise, min_region_width=2, N_sigma=4.0, extend=False, std_min=2, std_max=11):
    det = detection_ratio(wavelength, flux, noise, std_min, std_max)
    return select_regions(det, wavelength, flux, noise, min_region_width, N_sigma, extend), det
