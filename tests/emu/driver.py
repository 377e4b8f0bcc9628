"""TEST INFRASTRUCTURE: builds and drives tests/emu/emu.cpp — the CUDA kernel bodies compiled for the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "watersurfacerendering_b200", "csrc")
LIB = os.path.join(HERE, "libwsoemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [os.path.join(HERE, "emu.cpp"), os.path.join(CSRC, "wso_kernels.cuh"),
                os.path.join(CSRC, "wso_device.cuh")]
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-std=c++17", "-O2", "-w", "-ffp-contract=off", "-fPIC", "-shared",
                                   "-I", CSRC, "-o", LIB, os.path.join(HERE, "emu.cpp")])
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.wso_emu_compute.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp,
                                      vp]
        L.wso_emu_compute.restype = C.c_int
        _lib = L
    return _lib


def compute(n, tile_length, lam, h0_re, h0_im, omega, t, variant=0, want_w=False, anim_period=200.0):
    """Run emulated K1+K2+K3 for one tile-frame. h0_* are (N,N) row-major [m][n] (reference layout).
    anim_period=None forces the direct sincosf path; otherwise the sincos table is used when omega allows."""
    logn = int(np.log2(n))
    amp_t = np.ascontiguousarray(np.stack([h0_re.T, h0_im.T], axis=-1).astype(np.float32))
    om_t = np.ascontiguousarray(omega.T.astype(np.float32))
    idx = np.arange(n, dtype=np.float32)
    kv = (np.pi * (np.float32(2) * idx - np.float32(n)).astype(np.float64)
          / np.float64(np.float32(tile_length))).astype(np.float32)
    disp = np.zeros((n, n, 4), np.float32)
    norm = np.zeros((n, n, 4), np.float32)
    mm = np.zeros(2, np.float32)
    a = np.zeros(1, np.float32)
    w = np.zeros((n // 2, 4, n), np.complex64) if want_w else None
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    omega0 = 0.0 if anim_period is None else float(
        np.float32(np.float64(np.float32(2.0)) * np.pi / np.float64(np.float32(anim_period))))
    rc = lib().wso_emu_compute(logn, variant, p(amp_t), p(om_t), p(kv), omega0, float(lam), float(t), p(disp),
                               p(norm), p(mm), p(a), p(w))
    if rc != 0:
        raise ValueError(f"no emulated configuration for logn={logn} variant={variant}")
    return a[0], disp, norm, mm[0], mm[1], w
