"""CPU checks of the drop-in boundary: libwsocean.so loads and exports every symbol include/wsocean.h
declares, and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from watersurfacerendering_b200 import _lib as L
from watersurfacerendering_b200 import build as B

HEADER = os.path.join(ROOT, "include", "wsocean.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"WSO_API\s+[\w\s\*]+?\b(wso_\w+)\s*\(", src)))


def test_library_builds_and_loads():
    B.build()
    assert os.path.exists(L.LIB_PATH)
    lib = L.load()
    assert b"sm_100a" in lib.wso_version()


def test_every_declared_symbol_is_exported_and_bound():
    lib = C.CDLL(B.build())
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in wsocean.h but not exported"
        assert n in L.SYMBOLS, f"{n} declared in wsocean.h but not bound in _lib.py"
    assert sorted(L.SYMBOLS) == names


def test_only_wso_symbols_are_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    ours = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert ours and all(s.startswith("wso_") for s in ours), ours


def test_default_params_are_the_reference_defaults():
    lib = L.load()
    p = L.WsoParams()
    assert lib.wso_default_params(C.byref(p)) == L.WSO_OK
    # reference: WSTessendorf.h:36-43, 181
    assert (p.tile_size, p.tile_length, p.wind_dir_x, p.wind_dir_y) == (512, 1000.0, 1.0, 1.0)
    assert p.wind_speed == 30.0 and p.anim_period == 200.0 and abs(p.phillips_const - 3e-7) < 1e-12
    assert abs(p.damping - 0.1) < 1e-7 and p.lambda_ == -1.0
    assert lib.wso_default_params(None) == L.WSO_ERR_INVALID_ARG


def test_create_validates_arguments_before_touching_the_gpu():
    lib = L.load()
    p = L.WsoParams()
    lib.wso_default_params(C.byref(p))
    h = C.c_void_p()
    for bad in (0, 3, 100, 8, 16384):   # not pow2 / below 16 / above the single-CTA limit
        p.tile_size = bad
        assert lib.wso_create(C.byref(p), 0, 1, 1, C.byref(h)) == L.WSO_ERR_BAD_TILE_SIZE
        assert not h.value
    p.tile_size = 512
    assert lib.wso_create(C.byref(p), 0, 0, 1, C.byref(h)) == L.WSO_ERR_INVALID_ARG
    p.tile_length = -1.0
    assert lib.wso_create(C.byref(p), 0, 1, 1, C.byref(h)) == L.WSO_ERR_INVALID_ARG


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    p = L.WsoParams()
    lib.wso_default_params(C.byref(p))
    h = C.c_void_p()
    assert lib.wso_create(C.byref(p), 0, 1, 1, C.byref(h)) == L.WSO_ERR_CUDA
    assert b"no CPU fallback" in lib.wso_last_error(None)
    from watersurfacerendering_b200 import WSTessendorf
    with pytest.raises(L.WsoError):
        WSTessendorf(64, 100.0)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    pkg = os.path.join(ROOT, "watersurfacerendering_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
