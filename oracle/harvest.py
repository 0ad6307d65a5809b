"""Oracle: the result harvest of ``VPspectrum.fit_spectrum`` (vamp_1.0/vpspectrum.py:351-426).

TEST INFRASTRUCTURE ONLY.  Given the per-region fitted parameters (what the reference reads off ``region.fit``:
``estimated_variables[i][...].value``, ``estimated_profiles[k].value``, ``chain_covariance`` / ``mcmc.stats()``) it
restates how the reference turns them into the ``params`` and ``flux_model`` dictionaries, line by line, with the
oracle density (``oracle/density.py``) for the profiles and the reference's physics formulas restated inline
(vamp_1.0/physics.py; the product's copies are pinned to the reference's own outputs by tests/golden/physics.npz).
``tests/test_spectrum_gpu.py`` feeds it the MAP parameters of a GPU fit and compares key by key with what
``vamp_b200.vpspectrum.VPspectrum`` harvested.
"""
import numpy as np

from . import density as D

C = 2.98e8            # vamp_1.0/physics.py:3  (the reference's own value)
SIGMA0 = 0.0263       # vamp_1.0/physics.py:4


def freq2wave(frequency):
    """physics.py:116-120."""
    return (C / np.asarray(frequency, dtype=np.float64)) / 1.e-10


def column_density(amplitude, sigma):
    """physics.py:6-15."""
    return amplitude * sigma * np.sqrt(2 * np.pi) / SIGMA0


def doppler_parameter(sigma, line):
    """physics.py:17-27."""
    return (line * 1.e-10 * sigma * 2.355 / np.sqrt(2)) * 1.e-3


def equivalent_width_tau(taus, edges):
    """physics.py:29-42."""
    return np.sum(1 - np.exp(-1 * np.asarray(taus))) * (np.abs(edges[-1] - edges[0]) / (len(edges) - 1))


def equivalent_width_flux(fluxes, edges):
    """physics.py:45-58."""
    delta = (np.abs(edges[-1] - edges[0])) / (len(edges) - 1)
    return np.sum((1 - np.asarray(fluxes)) * delta)


def error_n(amplitude, sigma, std_a, std_s, cov_as):
    """physics.py:71-87 (the covariance argument is accepted and ignored, as in the reference)."""
    prefactor = np.sqrt(2. * np.pi) / SIGMA0
    return prefactor * np.sqrt(sigma ** 2 * std_a ** 2 + amplitude ** 2 * std_s ** 2)


def harvest(wavelength_array, frequency_array, flux_array, region_pixels, fits, line, voigt, best_chi_squared,
            mcmc_cov=False, get_mcmc_err=False, difficult_fit=False):
    """vpspectrum.py:260-266 (the dictionaries) and :351-426 (the per-region harvest).

    ``fits[j]``: dict with ``theta [n, 4]`` = (amplitude, centroid, sigma | L_fwhm, - | G_fwhm) in Hz - the MAP values
    the reference reads from ``estimated_variables`` - and, for ``mcmc_cov``, ``cov [n, 3, 3]`` = ``chain_covariance``
    (vpfits.py:432-456: covariance of (amplitude, sigma, centroid) samples), for ``get_mcmc_err`` ``std_a`` / ``std_s``
    = ``mcmc.stats()[...]['standard deviation']``."""
    R = len(region_pixels)
    params = {'b': np.array([]), 'b_std': np.array([]), 'N': np.array([]), 'N_std': np.array([]),
              'EW': np.array([]), 'centers': np.array([]), 'region_numbers': np.array([])}
    flux_model = {'total': np.ones(len(flux_array)), 'chi_squared': np.zeros(R), 'region_pixels': region_pixels,
                  'amplitude': np.array([]), 'sigmas': np.array([]), 'centers': np.array([]),
                  'region_numbers': np.array([]), 'EW': np.zeros(R), 'std_a': np.array([]), 'std_s': np.array([]),
                  'std_c': np.array([]), 'cov_as': np.array([]), 'difficult_fit': difficult_fit}
    for j, (start, end) in enumerate(region_pixels):
        waves = np.flip(wavelength_array[start:end], 0)                    # :274-277
        nu = np.flip(frequency_array[start:end], 0)
        th = np.asarray(fits[j]["theta"], dtype=np.float64)
        n = th.shape[0]
        taus = D.component_taus(nu, th, voigt)                             # estimated_profiles[k].value
        total = D.tau2flux(taus.sum(axis=0))                               # fit.total.value (vpfits.py:334-336)
        flux_model['chi_squared'][j] = best_chi_squared[j]                 # :353
        flux_model['total'][start:end] = np.flip(total, 0)                 # :358
        flux_model['region_' + str(j) + '_wave'] = np.flip(waves, 0)       # :360
        flux_model['region_' + str(j) + '_flux'] = np.ones((n, len(nu)))
        for k in range(n):
            flux_model['region_' + str(j) + '_flux'][k] = np.flip(D.tau2flux(taus[k]), 0)          # :363
        flux_model['EW'][j] = equivalent_width_flux(edges=flux_model['region_' + str(j) + '_wave'],
                                                    fluxes=flux_model['region_' + str(j) + '_flux'])   # :365-366
        heights = th[:, 0].copy()                                          # :368
        centers = freq2wave(th[:, 1])                                      # :369
        region_numbers = np.arange(n)
        sigmas = D.gaussian_width(th[:, 3]) if voigt else th[:, 2].copy()  # :372-377
        for key, val in (('amplitude', heights), ('centers', centers), ('region_numbers', region_numbers),
                         ('sigmas', sigmas)):
            flux_model[key] = np.append(flux_model[key], val)              # :379-382
        if mcmc_cov:                                                       # :384-397
            cov = np.asarray(fits[j]["cov"])
            std_a = np.sqrt([cov[i][0][0] for i in range(n)])
            std_s = np.sqrt([cov[i][1][1] for i in range(n)])
            std_c = np.sqrt([cov[i][2][2] for i in range(n)])
            cov_as = np.array([cov[i][0][1] for i in range(n)])
            for key, val in (('std_a', std_a), ('std_s', std_s), ('std_c', std_c), ('cov_as', cov_as)):
                flux_model[key] = np.append(flux_model[key], val)
            params['N_std'] = np.append(params['N_std'], error_n(heights, sigmas, std_a, std_s, cov_as))
        elif get_mcmc_err:                                                 # :399-418
            std_s, std_a = np.asarray(fits[j]["std_s"]), np.asarray(fits[j]["std_a"])
            flux_model['std_s'] = np.append(flux_model['std_s'], std_s)
            flux_model['std_a'] = np.append(flux_model['std_a'], std_a)
            params['b_std'] = np.append(params['b_std'], doppler_parameter(std_s, line))             # ErrorB
            params['N_std'] = np.append(params['N_std'], error_n(heights, sigmas, std_a, std_s, [0] * len(std_a)))
        params['b'] = np.append(params['b'], doppler_parameter(sigmas, line))                        # :420-426
        params['N'] = np.append(params['N'], column_density(heights, sigmas))
        params['centers'] = np.append(params['centers'], centers)
        params['region_numbers'] = np.append(params['region_numbers'], region_numbers)
        for k in range(n):
            params['EW'] = np.append(params['EW'], equivalent_width_tau(taus[k], [waves[0], waves[-1]]))
    return params, flux_model
