"""CPU emulation of the warp-per-line kernel bodies (wso_kernels2.cuh) versus the oracle.

tests/emu/emu2.cpp compiles the exact device code of K1 / K2h / K2 against a fiber-based SIMT context
(tests/emu/fiber_simt.h): every CUDA thread is a fiber, warp shuffles / __syncwarp / named barriers / __syncthreads
block until all participants arrive, 1-D bulk copies land at issue.  This pins the lane and register index logic of the
radix-32 x radix-L transforms, the mirror exchanges by shuffle, the second-stage thread assignment, the special rows
0 and N/2 and the persistent double-buffered line pipeline without a GPU; the -m gpu tests check the real thing.
"""
import functools

import numpy as np
import pytest

import packed_model as M
from conftest import SCALAR_REL_TOL, assert_maps_close, rel_l2
from emu import driver as E
from oracle import port as P

TIMES = (37.125, 3.5)


@functools.lru_cache(maxsize=2)
def _case(n):
    rng = np.random.default_rng(n)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    p = P.OceanParams(tile_size=n, tile_length=1000.0 * n / 512)
    o = P.PortOracle(p)
    h0 = o.prepare(xi)
    refs = [o.compute_waves(t) for t in TIMES]
    return p, h0, refs


# (n, variant): K1 tiling CP x NF and K2 line groups per CTA / CTAs, see CFG() in tests/emu/emu2.cpp
@pytest.mark.parametrize("n,variant,frames", [(512, 0, 2), (512, 1, 1), (512, 2, 2), (1024, 0, 1), (1024, 1, 2),
                                              (1024, 2, 1), (1024, 3, 1), (2048, 0, 1), (2048, 1, 1), (2048, 2, 1)])
def test_warp_core_vs_oracle(n, variant, frames):
    p, h0, refs = _case(n)
    times = TIMES[:frames]
    a, disp, norm, mn, mx, _ = E.compute2(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], times,
                                          variant=variant)
    for k in range(frames):
        a_ref, d_ref, n_ref = refs[k]
        assert_maps_close(disp[k], norm[k], d_ref, n_ref, f"N={n} v{variant} frame {k}")
        assert abs(a[k] - a_ref) <= SCALAR_REL_TOL * a_ref
        assert a[k] == max(abs(mn[k]), abs(mx[k]))


@pytest.mark.parametrize("n,variant", [(512, 0), (1024, 0)])
def test_warp_core_intermediate_layout(n, variant):
    """K1's Hermitian-packed intermediate W[m'][f][j][half] bin for bin against the float64 model: the layout the
    CTA-per-line kernels read and write as well (the two kernel sets can be mixed)."""
    p, h0, _ = _case(n)
    t = TIMES[0]
    _, _, _, _, _, w = E.compute2(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], [t], variant=variant,
                                  want_w=True)
    idx = np.arange(n, dtype=np.float32)
    kv = (np.pi * (np.float32(2) * idx - np.float32(n)).astype(np.float64)
          / np.float64(np.float32(p.tile_length))).astype(np.float32)
    Wm = M.pass1(M.evolve_Z(n, kv, h0["re"], h0["im"], h0["omega"], t))
    for f in range(4):
        assert rel_l2(w[0][:, f, :], Wm[:, f, :]) < 2e-6, f"field {f}"


def test_warp_core_one_hot_layout():
    """One-hot spectra at the special wave vectors (index 0 = Nyquist line, N/2 = DC line, N-1, interior) must light up
    exactly the oracle's texels: index / Hermitian layout of the shuffle exchanges and the special-row paths."""
    n = 512
    p = P.OceanParams(tile_size=n, tile_length=1000.0)
    o = P.PortOracle(p)
    w0 = np.float32(np.float64(np.float32(2.0)) * np.pi / np.float64(np.float32(200.0)))  # (float)(2.0f * M_PI / T)
    for (m, c) in [(0, 0), (0, 5), (5, 0), (n // 2, 3), (3, n // 2), (n - 1, n - 1), (1, n - 1), (n // 2, n // 2),
                   (7, 9), (0, n // 2), (16, 32), (n - 16, 17), (255, 257), (32, 480)]:
        h0 = np.zeros((n, n), P.H0_DTYPE)
        h0["omega"] = np.float32(10) * w0  # omega(k) == omega(-k) everywhere: the pair-summed records apply
        h0[m, c] = (0.7, -0.3, 0.7, 0.3, np.float32(10) * w0)
        o.import_h0(h0)
        a_ref, d_ref, n_ref = o.compute_waves(3.0)
        a, disp, norm, mn, mx, _ = E.compute2(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], [3.0],
                                              variant=2)
        if a_ref > 1e-30:
            scale = max(np.abs(d_ref[..., :3]).max(), 1e-30)
            assert np.abs(disp[0][..., :3] - d_ref[..., :3]).max() <= 2e-6 * scale, (m, c)
            scale = max(np.abs(n_ref).max(), 1e-30)
            assert np.abs(norm[0] - n_ref).max() <= 2e-6 * scale, (m, c)
