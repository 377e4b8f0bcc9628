"""GPU tests (-m gpu) of Prepare() on the device: wso_prepare_gauss_device / wso_prepare_counter /
wso_slab_prepare_counter_device against the oracle's Prepare (reference: WSTessendorf.cpp:60-148, WSTessendorf.h:237-297)
on the same Gaussian array.

Gate: wave-vector layout, quantised dispersion and heightAmp_conj == conj(heightAmp) bit-exact; amplitudes within 2 ulp of
the oracle's (the device evaluates the two exp() of the Phillips spectrum in float64 and rounds once, the reference calls
expf - they may round differently in rare wave vectors) with at most 2 % of the wave vectors differing at all; the maps
computed from the device-built spectrum pass the usual parity gate against the oracle's maps.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import SCALAR_REL_TOL, assert_maps_close, h0_struct, load_golden
from oracle import port as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wso():
    import watersurfacerendering_b200 as W
    return W


def _ulp_diff(a, b):
    """Distance in units in the last place between two fp32 arrays (same sign assumed where both are non-zero)."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2 ** 31) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2 ** 31) - ib, ib)
    return np.abs(ia - ib)


def _assert_h0_close(got, ref, what):
    assert got["omega"].tobytes() == ref["omega"].tobytes(), f"{what}: dispersion must be bit-exact"
    assert np.array_equal(got["re_c"], got["re"]) and np.array_equal(got["im_c"], -got["im"]), f"{what}: conj"
    for f in ("re", "im"):
        d = _ulp_diff(got[f], ref[f])
        assert d.max() <= 2, f"{what}: {f} differs by {d.max()} ulp"
        assert (d != 0).mean() <= 0.02, f"{what}: {100 * (d != 0).mean():.2f} % of {f} differ"
    # zeros (the DC wave vector) are exact zeros
    assert np.array_equal(got["re"] == 0, ref["re"] == 0)


def _oracle(n, seed=0, **kw):
    p = P.OceanParams(tile_size=n, tile_length=1000.0 * n / 512, **kw)
    o = P.PortOracle(p)
    rng = np.random.default_rng(seed + n)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    o.prepare(xi)
    return p, o, xi


def _apply(ws, p):
    ws.SetTileLength(p.tile_length)
    ws.SetWindDirection((p.wind_x, p.wind_y))
    ws.SetWindSpeed(p.wind_speed)
    ws.SetPhillipsConst(p.phillips_const)
    ws.SetDamping(p.damping)
    ws.SetAnimationPeriod(p.anim_period)
    ws.SetLambda(p.lam)


@pytest.mark.parametrize("n", [16, 64, 256, 512, 1024])
def test_device_prepare_matches_oracle_prepare(wso, n):
    p, o, xi = _oracle(n)
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGaussOnDevice(xi)
        _assert_h0_close(ws.ExportH0(), o.h0, f"N={n}")
        for t in (0.0, 7.75):
            a = ws.ComputeWaves(t)
            a_ref, d_ref, n_ref = o.compute_waves(t)
            assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), d_ref, n_ref, f"device prepare N={n} t={t}")
            assert abs(a - a_ref) <= SCALAR_REL_TOL * a_ref


@pytest.mark.parametrize("name", ["n64_wind", "n256_default"])
def test_device_prepare_against_reference_fixture(wso, name):
    """Gaussian array and h0 exported from the reference itself (tests/golden): non-default wind, damping, period."""
    g, params = load_golden(name)
    p = P.OceanParams(**params)
    with wso.WSTessendorf(params["tile_size"], params["tile_length"]) as ws:
        _apply(ws, p)
        ws.PrepareWithGaussOnDevice(g["xi"])
        _assert_h0_close(ws.ExportH0(), h0_struct(g["h0"]), name)
        for i, t in enumerate(g["t"]):
            a = ws.ComputeWaves(float(t))
            assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), g["disp"][i], g["norm"][i], f"{name} t={t}")
            assert abs(a - g["A"][i]) <= SCALAR_REL_TOL * g["A"][i]


def test_device_and_host_prepare_give_the_same_device_arrays(wso):
    """Same xi through wso_prepare_gauss (host build + upload) and wso_prepare_gauss_device: identical maps up to the
    1-ulp amplitude differences, and the SAME launch configuration (sincos table, pair records) afterwards."""
    n = 256
    p, o, xi = _oracle(n, seed=3)
    with wso.WSTessendorf(n, p.tile_length) as a, wso.WSTessendorf(n, p.tile_length) as b:
        a.PrepareWithGauss(xi)
        b.PrepareWithGaussOnDevice(xi)
        for t in (0.0, 123.5):
            a.ComputeWaves(t)
            b.ComputeWaves(t)
            assert_maps_close(b.GetDisplacements(), b.GetNormals(), a.GetDisplacements(), a.GetNormals(), f"t={t}")


@pytest.mark.parametrize("n", [64, 512])
def test_counter_prepare_on_device_matches_host_generator(wso, n):
    """wso_prepare_counter (Gaussian array drawn in the kernel) against wso_counter_h0 (host build of the same generator)."""
    from watersurfacerendering_b200 import _lib as L
    from watersurfacerendering_b200.slab import counter_h0
    seed = 77
    with wso.WSTessendorf(n, 1000.0 * n / 512) as ws:
        ws.SetWindDirection((0.3, -1.0))
        ws.PrepareCounterOnDevice(seed)
        got = ws.ExportH0()
        # wso_counter_h0 takes parameters as wso_create does (it normalises the wind direction itself): hand it the
        # direction as given to the setter, not the stored, already normalised one
        raw = ws._raw()
        raw.wind_dir_x, raw.wind_dir_y = 0.3, -1.0
        ref = counter_h0(raw, seed, 0, n).reshape(n, n)
        assert got["omega"].tobytes() == ref["omega"].tobytes()
        for f in ("re", "im"):
            # the Box-Muller draw is float64 on both sides; its rounding to fp32 and the expf may each differ by 1 ulp
            d = _ulp_diff(got[f], ref[f])
            assert d.max() <= 3 and (d != 0).mean() <= 0.02
        # feed the host-generated spectrum to a second context: maps agree within the gate
        with wso.WSTessendorf(n, 1000.0 * n / 512) as ws2:
            ws2.SetWindDirection((0.3, -1.0))
            ws2.ImportH0(ref)
            ws.ComputeWaves(3.25)
            ws2.ComputeWaves(3.25)
            assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), ws2.GetDisplacements(), ws2.GetNormals(), "counter")


def test_device_prepare_is_repeatable_and_honours_new_parameters(wso):
    n = 128
    p, o, xi = _oracle(n, seed=5)
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGaussOnDevice(xi)
        h_a = ws.ExportH0().copy()
        ws.PrepareWithGaussOnDevice(xi)
        assert ws.ExportH0().tobytes() == h_a.tobytes()
        ws.SetWindSpeed(12.0)
        ws.SetAnimationPeriod(120.0)
        ws.PrepareWithGaussOnDevice(xi)
        p2 = P.OceanParams(tile_size=n, tile_length=p.tile_length, wind_speed=12.0, anim_period=120.0)
        o2 = P.PortOracle(p2)
        o2.prepare(xi)
        _assert_h0_close(ws.ExportH0(), o2.h0, "after parameter change")


def test_slab_device_prepare_matches_host_prepare(wso):
    """One-rank slab context at 2048^2: the spectrum built by the device kernel against the host-built one, through the
    maps (both contexts run the same slab kernels)."""
    from watersurfacerendering_b200.slab import SlabBackend
    n, L_, seed, t = 2048, 4000.0, 7, 10.0
    outs = []
    for device_side in (False, True):
        b = SlabBackend(n, L_, 0, 1, 0)
        try:
            if device_side:
                b.prepare_counter_device(seed)
            else:
                b.prepare_counter(seed)
            b.pass1(t)
            b.heights()
            b.pass2()
            b.sync()
            outs.append((b.read_heights(), b.local_rows(0).copy(), b.local_rows(1).copy()))
        finally:
            b.close()
    (h_a, d_a, n_a), (h_b, d_b, n_b) = outs
    assert abs(h_a[0] - h_b[0]) <= SCALAR_REL_TOL * h_a[0]
    assert_maps_close(d_b, n_b, d_a, n_a, "slab device prepare")
