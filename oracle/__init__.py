"""CPU oracle for the VAMP likelihood / MCMC / BIC hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  ``vamp_b200`` never imports it and has no CPU
fallback: the product path fails loudly when ``libvamp_b200.so`` is missing.

What the oracle is
------------------
A numpy / scipy restatement of the arithmetic the reference (sarahappleby/VAMP,
``vamp_1.0``) performs on the hot path, each function citing the reference
``file:line`` it follows:

* ``density``  - tau profiles, ``exp(-sum tau)``, Gaussian log-likelihood, priors,
  chi-squared, BIC/AIC (``vamp_1.0/vpfits.py``, ``vamp_1.0/physics.py``,
  ``vamp_2.0/vamp_src/fit/fit.py``).
* ``sampler``  - restatement of the PyMC 2.3 scalar Metropolis sampler + MAP/BIC
  the reference delegates to (third party, absent from ``/root/reference``).
* ``region``   - the candidate-count ladder of ``vamp_1.0/vpregion.py:42-91``.
* ``chain_replay`` - a float64 replay of the *device* sampler (same Philox
  stream) used for step-by-step chain parity of the CUDA kernels.
* ``load_reference`` / ``make_golden`` - load the reference's own modules from
  ``/root/reference`` (build container only) and write ``tests/golden/*.npz``.

Parity pinning status
-----------------------------------------------
The reference ships no tests, golden vectors or fixtures for this path.
* PINNED by executing the reference's own lines in the build container
  (``make_golden.py``): ``VPfit.GaussFunction``, ``VPfit.GaussianWidth``,
  ``VPfit.Chisquared``/``ReducedChisquared``, ``physics.Tau2flux`` and friends,
  and the per-pixel-noise likelihood of ``vamp_2.0/vamp_src/fit/fit.py``.
* PINNED to ``scipy.special.wofz`` (what north_star names as the reference
  Voigt) and cross-checked against ``scipy.special.voigt_profile`` and mpmath:
  ``VPfit.VoigtFunction`` -> astropy ``Voigt1D`` (astropy itself is absent).
* PINNED by notebook arithmetic: ``BIC = k ln N - 2 lnL`` with ``sd`` counted in
  ``k`` (``vamp_1.0/vpfits_intro.ipynb`` cell 20).
* PARITY UNPINNED: the PyMC 2 sampler / MAP (third party, version-defined;
  statistical parity only) and the LSF convolution (absent from the reference).
"""
