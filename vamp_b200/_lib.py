"""ctypes binding of libvamp_b200.so (the C ABI declared in include/vamp_b200.h).

This is the stub a reference maintainer would add next to ``vamp_1.0/vpfits.py`` (see
INTEGRATION.md).  There is NO CPU fallback: if the shared library is missing or no B200 is
visible, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAMP_B200_LIB") or os.path.join(_HERE, "libvamp_b200.so")   # env override: kernel experiments

VAMP_F32, VAMP_F64, VAMP_MIXED = 0, 1, 2
VAMP_GAUSS, VAMP_VOIGT = 0, 1
VAMP_STEP_METROPOLIS, VAMP_STEP_ADAPTIVE = 0, 1
VAMP_AM_STRIDE = 40

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_void_pp = C.POINTER(C.c_void_p)


class McmcArgs(C.Structure):
    """``vamp_mcmc_args`` (include/vamp_b200.h)."""
    _fields_ = [
        ("n_chains", C.c_int), ("cmax", C.c_int), ("max_ncomp", C.c_int), ("sd_free", C.c_int),
        ("it0", C.c_int), ("it1", C.c_int), ("burn", C.c_int), ("thin", C.c_int),
        ("tune_interval", C.c_int), ("tune_throughout", C.c_int), ("seed", C.c_ulonglong),
        ("item_region", C.c_void_p), ("item_ncomp", C.c_void_p), ("chain_uid", C.c_void_p),
        ("theta", C.c_void_p), ("asf", C.c_void_p), ("psd", C.c_void_p), ("acc", C.c_void_p),
        ("rej", C.c_void_p), ("acc_total", C.c_void_p), ("logp", C.c_void_p), ("lnL", C.c_void_p),
        ("best_theta", C.c_void_p), ("best_logp", C.c_void_p), ("mean", C.c_void_p), ("m2", C.c_void_p),
        ("comom", C.c_void_p), ("ntally", C.c_void_p), ("sum_lnL", C.c_void_p), ("sum_lnL2", C.c_void_p),
        ("trace", C.c_void_p), ("trace_cap", C.c_int), ("work", C.c_void_p),
        ("step_kind", C.c_int), ("am_delay", C.c_int), ("am_interval", C.c_int), ("pair_tail", C.c_int),
        ("max_npix", C.c_int), ("am_state", C.c_void_p),
    ]


class MapArgs(C.Structure):
    """``vamp_map_args`` (include/vamp_b200.h)."""
    _fields_ = [
        ("n_starts", C.c_int), ("cmax", C.c_int), ("max_ncomp", C.c_int), ("sd_free", C.c_int),
        ("maxiter", C.c_int), ("maxfun", C.c_int), ("xtol", C.c_double), ("ftol", C.c_double),
        ("item_region", C.c_void_p), ("item_ncomp", C.c_void_p), ("theta", C.c_void_p),
        ("out_logp", C.c_void_p), ("out_lnL", C.c_void_p), ("out_chi2", C.c_void_p), ("out_bic", C.c_void_p),
        ("out_iters", C.c_void_p), ("out_nfev", C.c_void_p),
    ]


# name -> (restype, argtypes): every symbol include/vamp_b200.h declares
SIGNATURES = {
    "vamp_last_error": (C.c_char_p, []),
    "vamp_abi_version": (C.c_int, []),
    "vamp_ctx_create": (C.c_int, [C.c_int, c_void_pp]),
    "vamp_ctx_destroy": (C.c_int, [C.c_void_p]),
    "vamp_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vamp_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "vamp_regions_create": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int, c_void_pp]),
    "vamp_regions_destroy": (C.c_int, [C.c_void_p]),
    "vamp_regions_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p, C.POINTER(C.c_longlong)]),
    "vamp_regions_scale": (C.c_int, [C.c_void_p, c_dbl_p, c_dbl_p]),
    "vamp_regions_set_multiplet": (C.c_int, [C.c_void_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "vamp_detect_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_void_p]),
    "vamp_detect_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_int, C.c_int, c_dbl_p]),
    "vamp_loglike_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]),
    "vamp_loglike_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, C.c_int,
                                    c_dbl_p, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, C.c_longlong]),
    "vamp_mcmc_init_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(McmcArgs), C.c_int]),
    "vamp_mcmc_run_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(McmcArgs)]),
    "vamp_map_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(MapArgs)]),
    "vamp_profile_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dbl_p, C.c_double, C.c_double, C.c_double,
                                    C.c_double, c_dbl_p]),
    "vamp_faddeeva_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p]),
    "vamp_philox_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
}

_lib = None


class VampError(RuntimeError):
    pass


def load():
    """Load libvamp_b200.so (once).  Raises if it has not been built - no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VampError("%s not found: build it with `python -m vamp_b200.build` "
                        "(vamp_b200 has no CPU fallback)" % LIB_PATH)
    try:
        import torch  # noqa: F401  (brings libcudart.so.12 into the process)
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.vamp_abi_version() != 5:
        raise VampError("libvamp_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VampError("libvamp_b200: %s (code %d)" % (load().vamp_last_error().decode(), rc))


def as_dbl_p(a):
    return a.ctypes.data_as(c_dbl_p)


def as_int_p(a):
    return a.ctypes.data_as(c_int_p)
