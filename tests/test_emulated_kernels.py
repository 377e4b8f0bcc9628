"""CPU emulation of the sm_100a kernel bodies (tests/emu) versus the reference fixtures and the oracle.

The exact code of wso_kernels.cuh / wso_device.cuh (evolve + Hermitian packing, Stockham stages, shared-memory
layout, split/pack index logic) is compiled by g++ and stepped thread by thread.  This checks layout and index
logic without a GPU; the GPU parity tests (-m gpu) check the real thing.
"""
import numpy as np
import pytest

import packed_model as M
from conftest import SCALAR_REL_TOL, assert_maps_close, h0_struct, load_golden, rel_l2
from emu import driver as E
from oracle import port as P


@pytest.mark.parametrize("name,variants", [("n16_default", (0, 1)), ("n64_default", (0, 1, 2)),
                                           ("n64_wind", (0, 1)), ("n256_default", (0, 1))])
def test_emulated_kernels_vs_reference_fixture(name, variants):
    g, params = load_golden(name)
    h0 = g["h0"]
    for v in variants:
        for i, t in enumerate(g["t"]):
            a, disp, norm, mn, mx, _ = E.compute(params["tile_size"], params["tile_length"], params["lam"],
                                                 h0[..., 0], h0[..., 1], h0[..., 4], float(t), variant=v,
                                                 anim_period=params["anim_period"])
            # the per-frame sincos table must be bit-identical to per-point sincosf
            a2, disp2, norm2, _, _, _ = E.compute(params["tile_size"], params["tile_length"], params["lam"],
                                                  h0[..., 0], h0[..., 1], h0[..., 4], float(t), variant=v,
                                                  anim_period=None)
            assert a == a2 and disp.tobytes() == disp2.tobytes() and norm.tobytes() == norm2.tobytes()
            assert_maps_close(disp, norm, g["disp"][i], g["norm"][i], f"{name} v{v} t={t}")
            assert abs(a - g["A"][i]) <= SCALAR_REL_TOL * g["A"][i]
            assert abs(mn - g["minh"][i]) <= SCALAR_REL_TOL * g["A"][i]
            assert abs(mx - g["maxh"][i]) <= SCALAR_REL_TOL * g["A"][i]


def test_emulated_intermediate_layout_matches_model():
    """K1's Hermitian-packed intermediate W[m'][f][slot] bin-for-bin against the float64 model."""
    g, params = load_golden("n64_wind")
    h0 = g["h0"]
    n = params["tile_size"]
    t = float(g["t"][1])
    _, _, _, _, _, w = E.compute(n, params["tile_length"], params["lam"], h0[..., 0], h0[..., 1], h0[..., 4],
                                 t, variant=1, want_w=True)
    idx = np.arange(n, dtype=np.float32)
    kv = (np.pi * (np.float32(2) * idx - np.float32(n)).astype(np.float64)
          / np.float64(np.float32(params["tile_length"]))).astype(np.float32)
    Wm = M.pass1(M.evolve_Z(n, kv, h0[..., 0], h0[..., 1], h0[..., 4], t))
    for f in range(4):
        assert rel_l2(w[:, f, :], Wm[:, f, :]) < 2e-6, f"field {f}"


# (512, 0), (1024, 0), (2048, 0): tilings with NF * R0 == 8 - K1 runs its fused front end (evolve + first stage in
# registers); (512, 1): the unfused path
@pytest.mark.parametrize("n,variant", [(512, 0), (512, 1), (1024, 0), (2048, 0)])
def test_emulated_kernels_vs_oracle_large(n, variant):
    rng = np.random.default_rng(n)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    p = P.OceanParams(tile_size=n, tile_length=1000.0 * n / 512)
    o = P.PortOracle(p)
    h0 = o.prepare(xi)
    t = 37.125
    a_ref, d_ref, n_ref = o.compute_waves(t)
    a, disp, norm, mn, mx, _ = E.compute(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], t,
                                         variant=variant)
    assert_maps_close(disp, norm, d_ref, n_ref, f"N={n}")
    assert abs(a - a_ref) <= SCALAR_REL_TOL * a_ref


# variants 200+: the persistent forms (Pass1::run_persistent / Pass2::run_persistent) - a few emulated CTAs walk the work
# items with the inputs of the next item prefetched into the thread states.  Same arithmetic, different walk: the maps
# must equal the one-CTA-per-item kernels' BIT FOR BIT (and therefore pass the oracle gate).
@pytest.mark.parametrize("n,variant,base", [(512, 200, 0), (512, 201, 0), (1024, 200, 0), (1024, 201, 0), (2048, 200, None)])
def test_emulated_persistent_kernels(n, variant, base):
    rng = np.random.default_rng(n + 1)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    p = P.OceanParams(tile_size=n, tile_length=1000.0 * n / 512)
    o = P.PortOracle(p)
    h0 = o.prepare(xi)
    t = 11.375
    a, disp, norm, mn, mx, _ = E.compute(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], t, variant=variant)
    if base is not None:
        a0, d0, n0, mn0, mx0, _ = E.compute(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], t, variant=base)
        assert (a, mn, mx) == (a0, mn0, mx0)
        assert disp.tobytes() == d0.tobytes() and norm.tobytes() == n0.tobytes()
    if base is None or n == 512:
        a_ref, d_ref, n_ref = o.compute_waves(t)
        assert_maps_close(disp, norm, d_ref, n_ref, f"N={n} persistent")
        assert abs(a - a_ref) <= SCALAR_REL_TOL * a_ref


def test_one_hot_layout():
    """One-hot h0 at special wave vectors (index 0 = Nyquist line, N/2 = DC line, N-1) must light up
    exactly the oracle's pattern (index / Hermitian layout check)."""
    n = 16
    p = P.OceanParams(tile_size=n, tile_length=40.0)
    o = P.PortOracle(p)
    for (m, c) in [(0, 0), (0, 5), (5, 0), (n // 2, 3), (3, n // 2), (n - 1, n - 1), (1, n - 1), (n // 2, n // 2),
                   (7, 9), (0, n // 2)]:
        h0 = np.zeros((n, n), P.H0_DTYPE)
        h0[m, c] = (0.7, -0.3, 0.7, 0.3, 0.31415927)
        o.import_h0(h0)
        a_ref, d_ref, n_ref = o.compute_waves(3.0)
        a, disp, norm, mn, mx, _ = E.compute(n, p.tile_length, p.lam, h0["re"], h0["im"], h0["omega"], 3.0)
        if a_ref > 1e-30:
            scale = max(np.abs(d_ref[..., :3]).max(), 1e-30)
            assert np.abs(disp[..., :3] - d_ref[..., :3]).max() <= 2e-6 * scale, (m, c)
            scale = max(np.abs(n_ref).max(), 1e-30)
            assert np.abs(norm - n_ref).max() <= 2e-6 * scale, (m, c)


@pytest.mark.parametrize("name,variant,world", [("n64_default", 0, 1), ("n64_default", 0, 2), ("n64_wind", 0, 4),
                                                 ("n64_wind", 1, 2), ("n64_default", 1, 8), ("n256_default", 0, 4),
                                                 ("n256_default", 1, 2)])
def test_emulated_slab_decomposition_matches_single_device(name, variant, world):
    """Slab path (BASELINE config 5) on `world` emulated devices: K1 stores = the transpose into the row owners'
    buffers, min/max reduced over ranks, K2 on local rows (variant 1: the two-CTA cluster-pair K2).  Must agree
    with the single-device kernels BIT FOR BIT (same arithmetic, different placement) and with the fixture."""
    g, params = load_golden(name)
    h0 = g["h0"]
    n = params["tile_size"]
    i = len(g["t"]) - 1
    t = float(g["t"][i])
    a1, d1, n1, mn1, mx1, _ = E.compute(n, params["tile_length"], params["lam"], h0[..., 0], h0[..., 1], h0[..., 4], t,
                                        anim_period=params["anim_period"])
    a, disp, norm, mn, mx = E.compute_slab(n, params["tile_length"], params["lam"], h0[..., 0], h0[..., 1], h0[..., 4],
                                           t, world, variant=variant, anim_period=params["anim_period"])
    assert (a, mn, mx) == (a1, mn1, mx1)
    assert disp.tobytes() == d1.tobytes() and norm.tobytes() == n1.tobytes()
    assert_maps_close(disp, norm, g["disp"][i], g["norm"][i], f"{name} slab x{world}")


@pytest.mark.parametrize("name,variant", [("n16_default", 100), ("n64_wind", 101), ("n256_default", 100)])
def test_emulated_jacobian_channel(name, variant):
    """with the Jacobian switch on, displacement.w = J (oracle restatement of the reference's dead
    COMPUTE_JACOBIAN lines - parity unpinned, see PortOracle.jacobian) and every other channel is unchanged."""
    g, params = load_golden(name)
    h0 = g["h0"]
    n = params["tile_size"]
    o = P.PortOracle(P.OceanParams(tile_size=n, tile_length=params["tile_length"], lam=params["lam"]))
    o.import_h0(h0_struct(h0))
    for i, t in enumerate(g["t"][:2]):
        args = (n, params["tile_length"], params["lam"], h0[..., 0], h0[..., 1], h0[..., 4], float(t))
        a, disp, norm, mn, mx, _ = E.compute(*args, variant=variant, anim_period=params["anim_period"])
        a0, disp0, norm0, mn0, mx0, _ = E.compute(*args, variant=variant - 100, anim_period=params["anim_period"])
        assert (a, mn, mx) == (a0, mn0, mx0)
        assert_maps_close(disp, norm, g["disp"][i], g["norm"][i], f"{name} jacobian t={t}", skip_w=True)
        # packed fields 0, 2, 3 are untouched: bit-identical; Dz shares its complex transform with the new dzDx
        # field, so disp.z moves in the last bits only (covered by the gate above)
        assert np.array_equal(disp[..., :2], disp0[..., :2]) and np.array_equal(norm, norm0)
        j_ref = o.jacobian(float(t)).astype(np.float64)
        j = disp[..., 3].astype(np.float64)
        assert rel_l2(j, j_ref) <= 1e-5
        assert np.abs(j - j_ref).max() <= 1e-4 * max(j_ref.max() - j_ref.min(), 1e-30)
