"""ORACLE / TEST INFRASTRUCTURE — not product code.

Two CPU restatements of the reference hot path (kentril0/WaterSurfaceRendering,
src/scene/WSTessendorf.{h,cpp}):

* ``PortOracle``  — ctypes driver for ``oracle/libwsoracle.so`` (oracle/ws_oracle.cpp, plain C++).
* ``numpy_*``     — an independent NumPy restatement (fp32 element-wise math, float64 ``ifft2``).

Both are pinned against the reference's own code (oracle/_ref) in tests/test_oracle.py and against
the fixtures under tests/golden/.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB_PATH = os.path.join(_HERE, "libwsoracle.so")

H0_DTYPE = np.dtype(
    [("re", "<f4"), ("im", "<f4"), ("re_c", "<f4"), ("im_c", "<f4"), ("omega", "<f4")]
)


@dataclass
class OceanParams:
    """Defaults = the reference's (reference: WSTessendorf.h:36-43, 181)."""
    tile_size: int = 512
    tile_length: float = 1000.0
    wind_x: float = 1.0
    wind_y: float = 1.0
    wind_speed: float = 30.0
    phillips_const: float = 3e-7
    damping: float = 0.1
    anim_period: float = 200.0
    lam: float = -1.0


def build_port(force: bool = False) -> str:
    src = os.path.join(_HERE, "ws_oracle.cpp")
    if force or not os.path.exists(PORT_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(PORT_LIB_PATH)
    ):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_port()
        L = C.CDLL(PORT_LIB_PATH)
        vp, f32, i32 = C.c_void_p, C.c_float, C.c_int
        L.wso_oracle_wave_vectors.argtypes = [i32, f32, vp]
        L.wso_oracle_gauss_array.argtypes = [i32, C.c_uint, vp]
        L.wso_oracle_base_wave_heights.argtypes = [i32, f32, f32, f32, f32, f32, f32, f32, vp, vp]
        L.wso_oracle_spectra.argtypes = [i32, f32, vp, f32, vp]
        L.wso_oracle_compute_waves.argtypes = [i32, f32, f32, vp, f32, vp, vp, vp, i32]
        L.wso_oracle_compute_waves.restype = f32
        L.wso_oracle_set_threads.argtypes = [i32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class PortOracle:
    """C++ restatement, same call sequence as the reference model."""

    def __init__(self, params: OceanParams):
        self.p = params
        self.h0 = None
        self.min_height = np.float32(-1.0)   # reference: WSTessendorf.h:226-227
        self.max_height = np.float32(1.0)

    def wave_vectors(self) -> np.ndarray:
        n = self.p.tile_size
        out = np.zeros((n, n, 4), np.float32)
        lib().wso_oracle_wave_vectors(n, self.p.tile_length, _p(out))
        return out

    def gauss_array(self, seed: int) -> np.ndarray:
        n = self.p.tile_size
        xi = np.zeros((n, n), np.complex64)
        lib().wso_oracle_gauss_array(n, seed, _p(xi))
        return xi

    def prepare(self, xi: np.ndarray) -> np.ndarray:
        p, n = self.p, self.p.tile_size
        xi = np.ascontiguousarray(xi, np.complex64).reshape(n, n)
        h0 = np.zeros((n, n), H0_DTYPE)
        lib().wso_oracle_base_wave_heights(
            n, p.tile_length, p.wind_x, p.wind_y, p.wind_speed, p.phillips_const, p.damping,
            p.anim_period, _p(xi), _p(h0))
        self.h0 = h0
        return h0

    def import_h0(self, h0: np.ndarray):
        n = self.p.tile_size
        self.h0 = np.ascontiguousarray(h0, H0_DTYPE).reshape(n, n)

    def spectra(self, t: float) -> np.ndarray:
        n = self.p.tile_size
        out = np.zeros((7, n, n), np.complex64)
        lib().wso_oracle_spectra(n, self.p.tile_length, _p(self.h0), float(t), _p(out))
        return out

    def compute_waves(self, t: float, fft_mode: int = 0):
        """-> (A, disp[N,N,4], norm[N,N,4]); also sets min_height / max_height."""
        n = self.p.tile_size
        disp = np.zeros((n, n, 4), np.float32)
        norm = np.zeros((n, n, 4), np.float32)
        mm = np.zeros(2, np.float32)
        a = lib().wso_oracle_compute_waves(
            n, self.p.tile_length, self.p.lam, _p(self.h0), float(t), _p(disp), _p(norm), _p(mm),
            fft_mode)
        self.min_height, self.max_height = mm[0], mm[1]
        return np.float32(a), disp, norm


    def jacobian(self, t: float, parts: bool = False):
        """displacement.w under the reference's COMPUTE_JACOBIAN switch -> float32 [N, N].

        PARITY UNPINNED: that switch is dead code in the reference (WSTessendorf.cpp:158-175, 330-335, 368-377, 421-428;
        it does not compile there: m_dzDisplacementX / m_dxDisplacementZ and their plans are never declared in the
        header), so there is no reference output to check this restatement against.  It follows the intent of those
        lines: two more spectra dzDx = (0 + i*kz) * DisplacementX, dxDz = (0 + i*kx) * DisplacementZ (fp32 complex
        products, cpp:331-334), the same backward transform and sign, and
        J = (1 + l*s*dxDx)(1 + l*s*dzDz) - (l*s*dxDz)(l*s*dzDx) in fp32 (cpp:422-426)."""
        n = self.p.tile_size
        spec = self.spectra(t)
        k = numpy_wave_numbers(n, self.p.tile_length)
        ikz = (1j * k[:, None].astype(np.complex64)).astype(np.complex64)
        ikx = (1j * k[None, :].astype(np.complex64)).astype(np.complex64)
        dz_dx = (ikz * spec[3]).astype(np.complex64)
        dx_dz = (ikx * spec[4]).astype(np.complex64)

        def back(x):  # FFTW_BACKWARD, unnormalised, float64 rounded once to fp32 (as oracle/cpu_fft.h does)
            return (np.fft.ifft2(x.astype(np.complex128)).real * (float(n) * n)).astype(np.float32)

        idx = np.arange(n)
        sign = np.where(((idx[:, None] + idx[None, :]) & 1) == 1, np.float32(-1), np.float32(1)).astype(np.float32)
        lam = np.float32(self.p.lam)
        one = np.float32(1)
        dxdx, dzdz = back(spec[5]), back(spec[6])
        a, b = back(dx_dz), back(dz_dx)
        j = ((one + lam * sign * dxdx) * (one + lam * sign * dzdz)
             - (lam * sign * a) * (lam * sign * b)).astype(np.float32)
        if parts:  # the four signed derivative fields as well (tests)
            return j, sign * dxdx, sign * dzdz, sign * a, sign * b
        return j


# ------------------------------------------------------------------------------------------------
# Independent NumPy restatement
# ------------------------------------------------------------------------------------------------
F = np.float32


def numpy_wave_numbers(n: int, tile_length: float) -> np.ndarray:
    """reference: WSTessendorf.cpp:75-80 — fp32 index arithmetic, float64 product, narrowed once."""
    idx = np.arange(n, dtype=np.float32)
    centred = F(2.0) * idx - F(n)
    return (np.pi * centred.astype(np.float64) / np.float64(F(tile_length))).astype(np.float32)


def numpy_unit_vectors(kx: np.ndarray, kz: np.ndarray):
    """reference: WSTessendorf.h:133-136. kx, kz broadcastable fp32 -> (ux, uz, |k|)."""
    d = kx * kx + kz * kz
    ln = np.sqrt(d)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = F(1.0) / np.sqrt(d)
        ux = np.where(ln > F(1e-5), kx * inv, F(0))
        uz = np.where(ln > F(1e-5), kz * inv, F(0))
    return ux.astype(np.float32), uz.astype(np.float32), ln.astype(np.float32)


def numpy_base_wave_heights(p: OceanParams, xi: np.ndarray) -> np.ndarray:
    """reference: WSTessendorf.cpp:105-148 + WSTessendorf.h:237-263, 284-297."""
    n = p.tile_size
    kv = numpy_wave_numbers(n, p.tile_length)
    kx, kz = kv[None, :], kv[:, None]
    ux, uz, k = numpy_unit_vectors(kx, kz)
    winv = F(1.0) / np.sqrt(F(p.wind_x) * F(p.wind_x) + F(p.wind_y) * F(p.wind_y))
    wx, wy = F(p.wind_x) * winv, F(p.wind_y) * winv
    speed = max(F(0.0001), F(p.wind_speed))
    g = F(9.81)
    base_freq = F(np.float64(F(2.0)) * np.pi / np.float64(F(p.anim_period)))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore", under="ignore"):
        k2 = k * k
        k4 = k2 * k2
        cf = ux * wx + uz * wy
        cf = cf * cf
        L = speed * speed / g
        L2 = L * L
        ph = F(p.phillips_const) * np.exp(F(-1.0) / (k2 * L2)) / k4 * cf * np.exp(
            -k2 * F(p.damping) * F(p.damping))
        s = np.sqrt(ph.astype(np.float32))
        c = F(1.0) / np.sqrt(F(2.0))
        xi = np.asarray(xi, np.complex64).reshape(n, n)
        re = c * xi.real * s
        im = c * xi.imag * s
        omega = np.floor(np.sqrt(g * k) / base_freq) * base_freq
    ok = k > F(1e-5)
    h0 = np.zeros((n, n), H0_DTYPE)
    h0["re"] = np.where(ok, re, F(0))
    h0["im"] = np.where(ok, im, F(0))
    h0["re_c"] = h0["re"]
    h0["im_c"] = -h0["im"]
    h0["omega"] = np.where(ok, omega, F(0))
    return h0


def numpy_spectra(p: OceanParams, h0: np.ndarray, t: float) -> np.ndarray:
    """reference: WSTessendorf.cpp:294-336 — (7,N,N) complex64, the reference's buffer order."""
    n = p.tile_size
    kv = numpy_wave_numbers(n, p.tile_length)
    kx, kz = kv[None, :], kv[:, None]
    ux, uz, _ = numpy_unit_vectors(kx, kz)
    ph = (h0["omega"] * F(t)).astype(np.float32)
    pc = np.cos(ph.astype(np.float64)).astype(np.float32)
    ps = np.sin(ph.astype(np.float64)).astype(np.float32)

    def cmul(ar, ai, br, bi):
        return ar * br - ai * bi, ar * bi + ai * br

    r1, i1 = cmul(h0["re"], h0["im"], pc, ps)
    r2, i2 = cmul(h0["re_c"], h0["im_c"], pc, -ps)
    hr, hi = r1 + r2, i1 + i2
    z = np.zeros_like(hr)
    out = np.zeros((7, n, n), np.complex64)

    def put(i, c):
        out[i] = c[0] + 1j * c[1]

    dx = cmul(z, -ux + z, hr, hi)
    dz = cmul(z, -uz + z, hr, hi)
    put(0, (hr, hi))
    put(1, cmul(z, kx + z, hr, hi))
    put(2, cmul(z, kz + z, hr, hi))
    put(3, dx)
    put(4, dz)
    put(5, cmul(z, kx + z, *dx))
    put(6, cmul(z, kz + z, *dz))
    return out


def numpy_compute_waves(p: OceanParams, h0: np.ndarray, t: float):
    """reference: WSTessendorf.cpp:284-455 -> (A, disp, norm, min, max); FFT in float64."""
    n = p.tile_size
    spec = numpy_spectra(p, h0, t).astype(np.complex128)
    mm, nn = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    sign = np.where((mm + nn) & 1, -1.0, 1.0)
    fld = [(sign * np.real(np.fft.ifft2(spec[i]) * (n * n))).astype(np.float32) for i in range(7)]
    h = fld[0]
    hmax = max(np.float32(np.finfo(np.float32).tiny), h.max())
    hmin = min(np.float32(np.finfo(np.float32).max), h.min())
    a = max(abs(hmin), abs(hmax))
    inv = F(1.0) / F(a)
    lam = F(p.lam)
    disp = np.stack([lam * fld[3], h * inv, lam * fld[4], np.ones_like(h)], axis=-1).astype(np.float32)
    norm = np.stack([fld[1], fld[2], fld[5], fld[6]], axis=-1).astype(np.float32)
    return F(a), disp, norm, F(hmin), F(hmax)
