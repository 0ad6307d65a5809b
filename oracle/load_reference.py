"""Import the reference's own modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so this
module is used solely by ``oracle/make_golden.py`` (which writes committed
fixtures) and by CPU tests that skip when the tree is absent.

The reference cannot be imported as-is: ``vpfits.py`` needs
matplotlib, pymc 2 and astropy; ``vpregion.py`` is a SyntaxError on Python 3
(py2 ``print`` at :128,:137).  We inject inert stub modules so that the
reference's *own lines* for the static profile / chi-squared functions execute,
and load ``vpregion.py`` truncated before its dead py2-only function.
Nothing here copies reference source into the repo: files are read and exec'd
from where they lie.
"""
import importlib.util
import os
import sys
import types

import numpy as np
from scipy.special import wofz

REFERENCE_ROOT = os.environ.get("VAMP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vamp_1.0", "vpfits.py"))


class _Voigt1DStub:
    """Stand-in for ``astropy.modeling.models.Voigt1D`` (wofz method)."""

    def __init__(self, x_0, amplitude_L, fwhm_L, fwhm_G):
        self.p = (x_0, amplitude_L, fwhm_L, fwhm_G)

    def __call__(self, x):
        x_0, amplitude_L, fwhm_L, fwhm_G = self.p
        z = (2.0 * (np.asarray(x, dtype=np.float64) - x_0) + 1j * fwhm_L) * np.sqrt(np.log(2.0)) / fwhm_G
        return wofz(z).real * np.sqrt(np.pi * np.log(2.0)) / fwhm_G * fwhm_L * amplitude_L


class _Anything(types.ModuleType):
    """Module whose every attribute is an inert callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _inert(*a, **k):
            return _Anything(name)

        return _inert

    def __call__(self, *a, **k):
        return _Anything("call")

    def __pow__(self, other):
        return self

    def __rtruediv__(self, other):
        return self


def _install_stubs():
    stubs = {}
    for name in ("matplotlib", "matplotlib.pyplot", "pymc", "h5py", "pandas_stub"):
        stubs[name] = _Anything(name)
    astropy = types.ModuleType("astropy")
    modeling = types.ModuleType("astropy.modeling")
    models = types.ModuleType("astropy.modeling.models")
    models.Voigt1D = _Voigt1DStub
    astropy.modeling = modeling
    modeling.models = models
    stubs.update({"astropy": astropy, "astropy.modeling": modeling, "astropy.modeling.models": models})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load():
    """Returns a namespace with ``physics``, ``vpfits`` (module), ``VPfit``, ``fit2``
    (vamp_2.0 fit.py) and ``VPregion`` (class from the truncated vpregion.py)."""
    if _CACHE:
        return types.SimpleNamespace(**_CACHE)
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    v1 = os.path.join(REFERENCE_ROOT, "vamp_1.0")
    saved = _install_stubs()
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("physics", "vpfits", "vpregion")}
    try:
        sys.path.insert(0, v1)
        physics = _load("physics", os.path.join(v1, "physics.py"))
        sys.modules["physics"] = physics
        vpfits = _load("vpfits", os.path.join(v1, "vpfits.py"))
        sys.modules["vpfits"] = vpfits
        fit2 = _load("_vamp2_fit", os.path.join(REFERENCE_ROOT, "vamp_2.0", "vamp_src", "fit", "fit.py"))
        # vpregion.py: exec only the VPregion class proper (lines 1-92).
        with open(os.path.join(v1, "vpregion.py")) as fh:
            head = "".join(fh.readlines()[:92])
        head = head.replace("from scipy.ndimage.filters import", "from scipy.ndimage import")
        ns = {"__name__": "vpregion"}
        exec(compile(head, os.path.join(v1, "vpregion.py"), "exec"), ns)
    finally:
        sys.path[:] = saved_path
        _restore(saved)
        _restore(saved_mods)
    _CACHE.update(physics=physics, vpfits=vpfits, VPfit=vpfits.VPfit, fit2=fit2, VPregion=ns["VPregion"])
    return types.SimpleNamespace(**_CACHE)


def load_spectrum_class():
    """The reference's ``VPspectrum`` class (vamp_1.0/vpspectrum.py), exec'd from where it lies with
    inert stubs for pymc / matplotlib / h5py and the truncated ``vpregion`` module (the file itself
    is a SyntaxError on Python 3).  ``compute_detection_regions`` and ``split_difficult_region``
    then run the reference's own lines."""
    if "VPspectrum" in _CACHE:
        return _CACHE["VPspectrum"]
    ref = load()
    v1 = os.path.join(REFERENCE_ROOT, "vamp_1.0")
    saved = _install_stubs()
    fake_region = types.ModuleType("vpregion")
    fake_region.VPregion = ref.VPregion
    saved_mods = {k: sys.modules.get(k) for k in ("physics", "vpfits", "vpregion")}
    sys.modules.update({"physics": ref.physics, "vpfits": ref.vpfits, "vpregion": fake_region})
    try:
        mod = _load("_ref_vpspectrum", os.path.join(v1, "vpspectrum.py"))
    finally:
        _restore(saved)
        _restore(saved_mods)
    _CACHE["VPspectrum"] = mod.VPspectrum
    return mod.VPspectrum


def read_simba_fixture(name):
    """Raw-offset reader for vamp_1.0/data/simba_*.h5 (no h5py here):
    contiguous little-endian float64[1000] datasets."""
    offsets = {"velocity": 2048, "flux": 10048, "wavelength": 18048, "tau": 28096,
               "noise": 36096, "density": 44096, "temp": 52096}
    path = os.path.join(REFERENCE_ROOT, "vamp_1.0", "data", name)
    raw = open(path, "rb").read()
    return {k: np.frombuffer(raw, dtype="<f8", count=1000, offset=o).copy() for k, o in offsets.items()}
