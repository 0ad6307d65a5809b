"""Host-side physics helpers with the reference's names and quirks (vamp_1.0/physics.py).

Cheap O(lines) numpy arithmetic that runs once per fitted region; kept on the host on purpose
(SURVEY.md section 2 row 2).  Quirks preserved verbatim because downstream catalogues depend
on them: c = 2.98e8 m/s and sigma0 = 0.0263 (physics.py:3-4), b uses 2.355/sqrt(2)
(physics.py:27), ErrorN ignores the covariance term (physics.py:85-87).
The only function on the hot path, ``Tau2flux`` (physics.py:98-105, called at vpfits.py:336),
is fused into the CUDA kernels; the numpy version here serves post-processing and plots.
"""
import numpy as np

constants = {"c": {"value": 2.98e8, "units": "m/s", "def": "Speed of light in a vacuum"},
             "sigma0": {"value": 0.0263, "units": "cm**2 / s", "def": "Cross section for absorption"}}


def ColumnDensity(amplitude, sigma):
    """Column density in cm**-2 (physics.py:6-15)."""
    return amplitude * sigma * np.sqrt(2 * np.pi) / constants["sigma0"]["value"]


def DopplerParameter(sigma, line):
    """Doppler b parameter in km/s; ``line`` in Angstrom (physics.py:17-27)."""
    return (line * 1.e-10 * sigma * 2.355 / np.sqrt(2)) * 1.e-3


def EquivalentWidthTau(taus, edges):
    """physics.py:29-42: sum of the flux decrement times the bin spacing."""
    return np.sum(1 - np.exp(-1 * np.asarray(taus))) * (np.abs(edges[-1] - edges[0]) / (len(edges) - 1))


def EquivalentWidthFlux(fluxes, edges):
    """physics.py:45-58."""
    delta = (np.abs(edges[-1] - edges[0])) / (len(edges) - 1)
    return np.sum((1 - np.asarray(fluxes)) * delta)


def ErrorB(std_s, line):
    """physics.py:61-69."""
    return DopplerParameter(std_s, line)


def ErrorN(amplitude, sigma, std_a, std_s, cov_as):
    """physics.py:71-87 (covariance deliberately ignored, as in the reference)."""
    prefactor = np.sqrt(2. * np.pi) / constants["sigma0"]["value"]
    return prefactor * np.sqrt(sigma ** 2 * std_a ** 2 + amplitude ** 2 * std_s ** 2)


def Errorl(std_f):
    """physics.py:90-96."""
    return constants["c"]["value"] * std_f / 1.e-10


def Tau2flux(tau):
    """physics.py:98-105."""
    return np.exp(-np.asarray(tau))


def Flux2tau(flux):
    """physics.py:107-114."""
    return -1 * np.log(flux)


def Freq2wave(frequency):
    """physics.py:116-120."""
    return (constants["c"]["value"] / frequency) / 1.e-10


def Wave2freq(wavelength):
    """physics.py:122-126."""
    return constants["c"]["value"] / (wavelength * 1.e-10)


def Wave2red(wave, rest_wave):
    """physics.py:128-134."""
    return (wave - rest_wave) / rest_wave


# Doublets of BASELINE config C3 (rest wavelength in Angstrom, oscillator strength f, damping Gamma in 1/s; Morton
# 2003 values).  NOT IN THE REFERENCE, which fits single lines only.
DOUBLETS = {
    "CIV": ((1548.204, 0.1899, 2.643e8), (1550.781, 0.09475, 2.628e8)),
    "MgII": ((2796.354, 0.6155, 2.625e8), (2803.531, 0.3058, 2.595e8)),
}


def doublet_ratios(line1, line2=None):
    """``(freq_ratio, amp_ratio, lorentz_ratio)`` of a doublet for ``Regions(multiplet=...)`` /
    ``fit_regions(multiplet=...)``: the second line sits at ``nu_2/nu_1`` times the first line's frequency (same
    velocity), its Doppler width scales likewise (same b), its optical depth by ``(f_2 lambda_2)/(f_1 lambda_1)``
    (same column density) and its Lorentzian width by ``Gamma_2/Gamma_1``.  ``line1`` / ``line2`` are
    ``(wavelength, f, Gamma)`` tuples, or ``line1`` names an entry of :data:`DOUBLETS`."""
    if line2 is None:
        line1, line2 = DOUBLETS[line1]
    (l1, f1, g1), (l2, f2, g2) = line1, line2
    return l1 / l2, (f2 * l2) / (f1 * l1), g2 / g1
