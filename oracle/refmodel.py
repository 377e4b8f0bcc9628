"""ORACLE / TEST INFRASTRUCTURE — not product code.

ctypes driver for ``oracle/_ref/libwsref.so``: the reference's own, unmodified
``src/scene/WSTessendorf.cpp`` (compiled from /root/reference by ``oracle/Makefile``) behind the
FFTW-API shim of ``oracle/ref_harness.cpp``.  Mirrors the reference surface-model API
(reference: src/scene/WSTessendorf.h:58-122).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libwsref.so")

# reference: WSTessendorf.h:142-147  (sizeof == 20, no padding)
H0_DTYPE = np.dtype(
    [("re", "<f4"), ("im", "<f4"), ("re_c", "<f4"), ("im_c", "<f4"), ("omega", "<f4")]
)

FFT_FLOAT64 = 0   # accuracy mode (parity)
FFT_FLOAT32 = 1   # timing mode ("reference code + shim FFT, not FFTW")
FFT_NOOP = 2      # leave the pre-FFT spectra in the work arrays
FFT_REAL_FFTW = 3  # dlopen("libfftw3f.so.3") when the box has one

_lib = None


def _cpu_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2" in f.read()
    except OSError:
        return False


def available() -> bool:
    """The shim FFT inside libwsref.so is built with -mavx2 (oracle/Makefile)."""
    return os.path.exists(REF_LIB_PATH) and _cpu_has_avx2()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(
                f"{REF_LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists"
            )
        L = C.CDLL(REF_LIB_PATH)
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        L.wsref_create.restype = vp
        L.wsref_create.argtypes = [u32, f32]
        L.wsref_destroy.argtypes = [vp]
        for name in ("tile_length", "wind_speed", "animation_period", "phillips_const", "lambda",
                     "damping"):
            getattr(L, f"wsref_set_{name}").argtypes = [vp, f32]
            getattr(L, f"wsref_get_{name}").argtypes = [vp]
            getattr(L, f"wsref_get_{name}").restype = f32
        L.wsref_set_tile_size.argtypes = [vp, u32]
        L.wsref_get_tile_size.argtypes = [vp]
        L.wsref_get_tile_size.restype = u32
        L.wsref_set_wind_direction.argtypes = [vp, f32, f32]
        L.wsref_get_wind_dir.argtypes = [vp, vp]
        for name in ("base_freq", "min_height", "max_height"):
            getattr(L, f"wsref_get_{name}").argtypes = [vp]
            getattr(L, f"wsref_get_{name}").restype = f32
        L.wsref_prepare.argtypes = [vp, C.c_uint]
        L.wsref_gauss_array.argtypes = [vp, C.c_uint, vp]
        L.wsref_prepare_with_gauss.argtypes = [vp, vp]
        L.wsref_h0_stride.restype = C.c_size_t
        L.wsref_export_h0.argtypes = [vp, vp]
        L.wsref_import_h0.argtypes = [vp, vp]
        L.wsref_export_wave_vectors.argtypes = [vp, vp]
        L.wsref_compute_waves.argtypes = [vp, f32]
        L.wsref_compute_waves.restype = f32
        L.wsref_get_displacements.argtypes = [vp, vp]
        L.wsref_get_normals.argtypes = [vp, vp]
        L.wsref_export_work_arrays.argtypes = [vp, vp]
        L.wsref_set_fft_mode.argtypes = [C.c_int]
        L.wsref_set_fft_mode.restype = C.c_int
        L.wsref_set_threads.argtypes = [C.c_int]
        assert L.wsref_h0_stride() == H0_DTYPE.itemsize == 20
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefWSTessendorf:
    """The reference class, driven from Python.  Method names follow the reference's."""

    def __init__(self, tile_size: int = 512, tile_length: float = 1000.0, fft_mode: int = FFT_FLOAT64):
        self._L = lib()
        if self._L.wsref_set_fft_mode(fft_mode) != 0:
            raise RuntimeError("requested FFT mode unavailable (no libfftw3f.so.3)")
        self._h = self._L.wsref_create(tile_size, tile_length)
        self._prepared = False

    def close(self):
        if self._h:
            self._L.wsref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # setters / getters (reference: WSTessendorf.cpp:459-505, WSTessendorf.h:82-91)
    def SetTileSize(self, n): self._L.wsref_set_tile_size(self._h, n)
    def SetTileLength(self, l): self._L.wsref_set_tile_length(self._h, l)
    def SetWindDirection(self, x, y): self._L.wsref_set_wind_direction(self._h, x, y)
    def SetWindSpeed(self, v): self._L.wsref_set_wind_speed(self._h, v)
    def SetAnimationPeriod(self, T): self._L.wsref_set_animation_period(self._h, T)
    def SetPhillipsConst(self, A): self._L.wsref_set_phillips_const(self._h, A)
    def SetLambda(self, l): self._L.wsref_set_lambda(self._h, l)
    def SetDamping(self, d): self._L.wsref_set_damping(self._h, d)

    def GetTileSize(self): return int(self._L.wsref_get_tile_size(self._h))
    def GetTileLength(self): return float(self._L.wsref_get_tile_length(self._h))
    def GetWindDir(self):
        xy = np.zeros(2, np.float32)
        self._L.wsref_get_wind_dir(self._h, _p(xy))
        return xy
    def GetWindSpeed(self): return float(self._L.wsref_get_wind_speed(self._h))
    def GetAnimationPeriod(self): return float(self._L.wsref_get_animation_period(self._h))
    def GetBaseFreq(self): return np.float32(self._L.wsref_get_base_freq(self._h))
    def GetPhillipsConst(self): return float(self._L.wsref_get_phillips_const(self._h))
    def GetDamping(self): return float(self._L.wsref_get_damping(self._h))
    def GetDisplacementLambda(self): return float(self._L.wsref_get_lambda(self._h))
    def GetMinHeight(self): return np.float32(self._L.wsref_get_min_height(self._h))
    def GetMaxHeight(self): return np.float32(self._L.wsref_get_max_height(self._h))

    def Prepare(self, seed: int = 1234):
        self._L.wsref_prepare(self._h, seed)
        self._prepared = True

    def PrepareWithGauss(self, xi: np.ndarray):
        n = self.GetTileSize()
        xi = np.ascontiguousarray(xi, dtype=np.complex64).reshape(n, n)
        self._L.wsref_prepare_with_gauss(self._h, _p(xi))
        self._prepared = True

    def GaussArray(self, seed: int) -> np.ndarray:
        n = self.GetTileSize()
        xi = np.zeros((n, n), np.complex64)
        self._L.wsref_gauss_array(self._h, seed, _p(xi))
        return xi

    def ExportH0(self) -> np.ndarray:
        assert self._prepared
        n = self.GetTileSize()
        h0 = np.zeros((n, n), H0_DTYPE)
        self._L.wsref_export_h0(self._h, _p(h0))
        return h0

    def ImportH0(self, h0: np.ndarray):
        assert self._prepared
        n = self.GetTileSize()
        h0 = np.ascontiguousarray(h0, dtype=H0_DTYPE).reshape(n, n)
        self._L.wsref_import_h0(self._h, _p(h0))

    def ExportWaveVectors(self) -> np.ndarray:
        assert self._prepared
        n = self.GetTileSize()
        wv = np.zeros((n, n, 4), np.float32)
        self._L.wsref_export_wave_vectors(self._h, _p(wv))
        return wv

    def ComputeWaves(self, t: float) -> np.float32:
        assert self._prepared
        return np.float32(self._L.wsref_compute_waves(self._h, float(t)))

    def GetDisplacements(self) -> np.ndarray:
        n = self.GetTileSize()
        d = np.zeros((n, n, 4), np.float32)
        self._L.wsref_get_displacements(self._h, _p(d))
        return d

    def GetNormals(self) -> np.ndarray:
        n = self.GetTileSize()
        d = np.zeros((n, n, 4), np.float32)
        self._L.wsref_get_normals(self._h, _p(d))
        return d

    def ExportWorkArrays(self) -> np.ndarray:
        """(7, N, N) complex64: Height, SlopeX, SlopeZ, Dx, Dz, dxDx, dzDz (post- or pre-FFT by mode)."""
        n = self.GetTileSize()
        w = np.zeros((7, n, n), np.complex64)
        self._L.wsref_export_work_arrays(self._h, _p(w))
        return w


def set_threads(n: int):
    lib().wsref_set_threads(int(n))


def max_threads() -> int:
    return int(lib().wsref_max_threads())
