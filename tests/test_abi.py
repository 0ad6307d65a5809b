"""CPU: the C-ABI library builds, loads and exports every symbol include/vamp_b200.h declares;
without a GPU every entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vamp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vamp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    from vamp_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/vamp_b200.h but not exported" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.vamp_abi_version() == 5


def test_struct_layouts_match_header():
    from vamp_b200 import _lib
    src = open(os.path.join(ROOT, "include", "vamp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, cls in (("vamp_mcmc_args", _lib.McmcArgs), ("vamp_map_args", _lib.MapArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = [re.sub(r"[\s\*]", "", d).split("[")[0] for d in re.sub(r"^[\w\s]+?(?=[\*\w]+(,|$))", "", decl).split(",")]
            # simpler: the last identifier(s) of the declaration
            ids = re.findall(r"\*?\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:,|$)", decl)
            fields.extend(ids)
        assert [f[0] for f in cls._fields_] == fields, cname


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vamp_b200 import _lib, engine
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.vamp_ctx_create(0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.vamp_last_error()
    with pytest.raises(_lib.VampError):
        engine.default_context()
    from vamp_b200.vpfits import VPfit
    import numpy as np
    with pytest.raises(_lib.VampError):
        VPfit.VoigtFunction(np.linspace(0, 1, 8), 0.5, 1.0, 0.1, 0.2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vamp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt, f
