"""GPU parity tests (-m gpu): the sm_100a path, called through the C ABI (ctypes -> libwsocean.so), against
the oracle on the same h0.

Gate (BASELINE.json north_star): per map channel rel-L2 <= 1e-5 and max-abs <= 1e-4 x channel range in
fp32; amplitude/min/max rel <= 1e-6; disp.w == 1.0 exactly; index/Hermitian layout exact
(one-hot tests).  /root/reference is never read here: the oracle is oracle/libwsoracle.so (restatement,
pinned in tests/test_oracle.py) and the committed fixtures under tests/golden/.
"""
import numpy as np
import pytest

from conftest import SCALAR_REL_TOL, assert_maps_close, h0_struct, load_golden, rel_l2
from oracle import port as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wso():
    import watersurfacerendering_b200 as W
    return W


def _oracle_for(n, L=None, seed=0, **kw):
    L = 1000.0 * n / 512 if L is None else L
    p = P.OceanParams(tile_size=n, tile_length=L, **kw)
    o = P.PortOracle(p)
    rng = np.random.default_rng(seed + n)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    o.prepare(xi)
    return p, o, xi


def _apply_params(ws, p: P.OceanParams, tile=0):
    ws.SetTileLength(p.tile_length, tile)
    ws.SetWindDirection((p.wind_x, p.wind_y), tile)
    ws.SetWindSpeed(p.wind_speed, tile)
    ws.SetPhillipsConst(p.phillips_const, tile)
    ws.SetDamping(p.damping, tile)
    ws.SetAnimationPeriod(p.anim_period, tile)
    ws.SetLambda(p.lam, tile)


def _check_frame(ws, o, t, what):
    a = ws.ComputeWaves(t)
    a_ref, d_ref, n_ref = o.compute_waves(t)
    assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), d_ref, n_ref, what)
    assert abs(a - a_ref) <= SCALAR_REL_TOL * a_ref, (what, a, a_ref)
    assert abs(ws.GetMinHeight() - o.min_height) <= SCALAR_REL_TOL * a_ref
    assert abs(ws.GetMaxHeight() - o.max_height) <= SCALAR_REL_TOL * a_ref


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["n16_default", "n64_default", "n64_wind", "n256_default"])
def test_reference_fixtures(wso, name):
    """Outputs of the reference's own code (tests/golden) for the exported h0."""
    g, params = load_golden(name)
    with wso.WSTessendorf(params["tile_size"], params["tile_length"]) as ws:
        ws.SetLambda(params["lam"])
        ws.ImportH0(h0_struct(g["h0"]))
        for i, t in enumerate(g["t"]):
            a = ws.ComputeWaves(float(t))
            assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), g["disp"][i], g["norm"][i], f"{name} t={t}")
            assert abs(a - g["A"][i]) <= SCALAR_REL_TOL * g["A"][i]
            assert abs(ws.GetMinHeight() - g["minh"][i]) <= SCALAR_REL_TOL * g["A"][i]
            assert abs(ws.GetMaxHeight() - g["maxh"][i]) <= SCALAR_REL_TOL * g["A"][i]


@pytest.mark.parametrize("name", ["n16_default", "n64_default", "n64_wind", "n256_default"])
def test_prepare_reproduces_reference_h0_bit_exact(wso, name):
    """Prepare() with srand(seed): same rand() consumption, same fp32 Phillips arithmetic as the reference."""
    g, params = load_golden(name)
    with wso.WSTessendorf(params["tile_size"], params["tile_length"]) as ws:
        _apply_params(ws, P.OceanParams(**params))
        ws.Prepare(seed=int(g["seed"]))
        assert ws.ExportH0().tobytes() == h0_struct(g["h0"]).tobytes()
        ws.PrepareWithGauss(g["xi"])
        assert ws.ExportH0().tobytes() == h0_struct(g["h0"]).tobytes()


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048])
def test_all_tile_sizes_vs_oracle(wso, n):
    p, o, xi = _oracle_for(n)
    times = [0.0, 1.5, 4711.25] if n <= 1024 else [36.6]
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGauss(xi)
        assert ws.ExportH0().tobytes() == o.h0.tobytes()
        for t in times:
            _check_frame(ws, o, t, f"N={n} t={t}")


def test_4096_vs_oracle(wso):
    n = 4096
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGauss(xi)
        _check_frame(ws, o, 12.5, "N=4096")


@pytest.mark.parametrize("kw", [
    dict(wind_x=-0.3, wind_y=1.0, wind_speed=8.0, lam=-0.5),
    dict(wind_x=1.0, wind_y=0.0, wind_speed=55.0, damping=0.4, lam=1.25),
    dict(phillips_const=9e-7, anim_period=60.0, lam=0.0),
    dict(wind_speed=1e-9),   # clamped to 1e-4 like SetWindSpeed: an (almost) flat ocean
])
def test_parameter_variants(wso, kw):
    n = 128
    p, o, xi = _oracle_for(n, L=77.0, seed=5, **kw)
    with wso.WSTessendorf(n, 77.0) as ws:
        _apply_params(ws, p)
        ws.PrepareWithGauss(xi)
        assert ws.ExportH0().tobytes() == o.h0.tobytes()
        a_ref, d_ref, n_ref = o.compute_waves(9.75)
        if a_ref < 1e-30:
            # Flat ocean (every h0 underflows to 0): the one input where the reference's start values decide the result.
            # masterMax starts at FLT_MIN - the smallest POSITIVE float, not the lowest (WSTessendorf.cpp:289) - so with
            # all heights 0 the reference returns A = max(|0|, |FLT_MIN|) = FLT_MIN, min 0, max FLT_MIN, and
            # NormalizeHeights (cpp:443-455) multiplies the zero heights by 1/FLT_MIN: still 0, no NaN.
            flt_min = np.float32(1.17549435e-38)
            assert np.float32(a_ref) == flt_min, "oracle: A must be FLT_MIN on a flat ocean"
            a = ws.ComputeWaves(9.75)
            assert np.float32(a) == flt_min
            assert np.float32(ws.GetMinHeight()) == np.float32(0.0) and np.float32(ws.GetMaxHeight()) == flt_min
            d, nm = ws.GetDisplacements(), ws.GetNormals()
            assert np.all(d[..., :3] == 0.0) and np.all(d[..., 3] == 1.0) and np.all(nm == 0.0)
            assert np.all(d_ref[..., :3] == 0.0) and np.all(n_ref == 0.0)
            return
        _check_frame(ws, o, 9.75, str(kw))


def test_lambda_takes_effect_without_prepare(wso):
    """reference: SetLambda needs no Prepare() (WaterSurfaceMesh.cpp:902)."""
    n = 64
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGauss(xi)
        for lam in (-1.0, 0.35, 2.0):
            ws.SetLambda(lam)
            o.p.lam = lam
            assert ws.GetDisplacementLambda() == np.float32(lam)
            _check_frame(ws, o, 3.25, f"lambda={lam}")


def test_one_hot_layout(wso):
    """Index / Hermitian layout: a single non-zero wave vector at the special indices (0 = Nyquist line,
    N/2 = DC line, N-1, and generic) must produce exactly the oracle's pattern."""
    n = 32
    p = P.OceanParams(tile_size=n, tile_length=40.0)
    o = P.PortOracle(p)
    with wso.WSTessendorf(n, 40.0) as ws:
        for (m, c) in [(0, 0), (0, 5), (5, 0), (n // 2, 3), (3, n // 2), (n - 1, n - 1), (1, n - 1),
                       (7, 9), (0, n // 2), (n // 2, 0), (n - 1, 0), (17, 30)]:
            h0 = np.zeros((n, n), P.H0_DTYPE)
            h0[m, c] = (0.7, -0.3, 0.7, 0.3, 0.31415927)
            o.import_h0(h0)
            ws.ImportH0(h0)
            a_ref, d_ref, n_ref = o.compute_waves(3.0)
            a = ws.ComputeWaves(3.0)
            assert abs(a - a_ref) <= 2e-6 * a_ref
            d, nm = ws.GetDisplacements(), ws.GetNormals()
            assert np.abs(d[..., :3] - d_ref[..., :3]).max() <= 2e-6 * np.abs(d_ref[..., :3]).max(), (m, c)
            assert np.abs(nm - n_ref).max() <= 2e-6 * max(np.abs(n_ref).max(), 1e-30), (m, c)
            assert np.all(d[..., 3] == 1.0)


def test_batch_matches_single_frames(wso):
    """Animation frames are independent (BASELINE config 2/3): a batch == the same frames one by one."""
    n = 256
    p, o, xi = _oracle_for(n)
    times = np.array([0.05 * i for i in range(37)], np.float32)   # 37: not a multiple of the chunk
    with wso.WSTessendorf(n, p.tile_length, max_slots=40) as ws:
        ws.PrepareWithGauss(xi)
        ws.compute_batch(times, first_slot=2)
        a, mn, mx = ws.read_heights(2, times.size)
        for i in (0, 1, 17, 36):
            d = ws.copy_map(0, 2 + i)
            nm = ws.copy_map(1, 2 + i)
            a1 = ws.ComputeWaves(float(times[i]))
            assert a1 == a[i] and ws.GetMinHeight() == mn[i] and ws.GetMaxHeight() == mx[i]
            assert d.tobytes() == ws.GetDisplacements().tobytes()
            assert nm.tobytes() == ws.GetNormals().tobytes()
        a_ref, d_ref, n_ref = o.compute_waves(float(times[36]))
        assert_maps_close(ws.copy_map(0, 38), ws.copy_map(1, 38), d_ref, n_ref, "batch frame 36")


@pytest.mark.parametrize("n", [64, 512, 2048])
def test_frame_graph_matches_plain_launches(wso, n):
    """One ComputeWaves(t) per frame goes out as one CUDA-graph launch (wso_set_frame_graph, wsocean.h): the maps are bit-
    identical to the three plain launches, the graph is captured once per shape and re-used with new parameters (time,
    lambda, a re-prepared spectrum), and a change of the tile size captures again."""
    p, o, xi = _oracle_for(n)
    times = [0.0, 0.35, 7.125, 12.5]
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGauss(xi)
        ws.set_frame_graph(False)
        plain = []
        for t in times:
            a = ws.ComputeWaves(t)
            plain.append((a, ws.GetDisplacements().copy(), ws.GetNormals().copy()))
        assert ws.stats()["frame_graph_launches"] == 0
        ws.set_frame_graph(True, every_size=True)
        for rep in range(2):
            for t, (a0, d0, n0) in zip(times, plain):
                a = ws.ComputeWaves(t)
                assert a == a0
                assert ws.GetDisplacements().tobytes() == d0.tobytes(), (n, t)
                assert ws.GetNormals().tobytes() == n0.tobytes(), (n, t)
        st = ws.stats()
        assert st["frame_graph_captures"] == 1 and st["frame_graph_launches"] >= 2 * len(times) - 1, st
        _check_frame(ws, o, 3.75, f"graph n={n}")
        # parameters that live in the kernel arguments: lambda, and a new spectrum in the same buffers
        ws.SetLambda(-0.5)
        d_before = ws.GetDisplacements().copy()
        ws.ComputeWaves(3.75)
        d_after = ws.GetDisplacements()
        assert np.array_equal(d_after[..., 1], d_before[..., 1]) and not np.array_equal(d_after[..., 0], d_before[..., 0])
        ws.SetLambda(p.lam)
        p2, o2, xi2 = _oracle_for(n, seed=5)
        ws.PrepareWithGauss(xi2)
        _check_frame(ws, o2, 1.25, f"graph n={n}, second spectrum")
        assert ws.stats()["frame_graph_captures"] == 1
        if n == 64:
            ws.SetTileSize(128)
            p3, o3, xi3 = _oracle_for(128)
            ws.SetTileLength(p3.tile_length)
            ws.PrepareWithGauss(xi3)
            for t in (0.5, 2.0, 9.0):
                _check_frame(ws, o3, t, "graph after a tile size change")
            assert ws.stats()["frame_graph_captures"] == 2
        if n == 2048:   # default policy: graph launches up to 512^2 only
            ws.set_frame_graph(True)
            before = ws.stats()["frame_graph_launches"]
            _check_frame(ws, o2, 2.5, "plain launches at 2048")
            assert ws.stats()["frame_graph_launches"] == before


def test_imported_spectrum_without_table_and_graph_recapture(wso):
    """An imported spectrum whose dispersion values are not multiples of one base frequency (nothing Prepare() builds, but
    wso_import_h0 accepts it) runs on the general K1 body: direct sincosf instead of the per-frame table.  Checked against the
    oracle fed the same records; the frame graph notices the other kernel variant and captures again, both ways."""
    n = 128
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.set_frame_graph(True, every_size=True)
        ws.PrepareWithGauss(xi)
        for t in (0.25, 1.0, 4.5):
            _check_frame(ws, o, t, "table spectrum")
        assert ws.stats()["frame_graph_captures"] == 1
        h0 = o.h0.copy()
        o.h0["omega"] = (o.h0["omega"] * np.float32(1.37)).astype(np.float32)
        ws.ImportH0(o.h0)
        for t in (0.25, 1.0, 4.5):
            _check_frame(ws, o, t, "spectrum without the table")
        assert ws.stats()["frame_graph_captures"] == 2
        o.h0[...] = h0
        ws.ImportH0(o.h0)
        for t in (0.5, 2.0):
            _check_frame(ws, o, t, "table spectrum again")
        assert ws.stats()["frame_graph_captures"] == 3


def test_frame_inside_a_callers_graph_capture(wso):
    """A caller that captures its own CUDA graph around the per-frame call (stream bound with wso_set_stream) gets the three
    kernels recorded as plain launches - the library's own frame graph stands aside - and the replay writes the same maps."""
    import torch
    n = 256
    p, o, xi = _oracle_for(n)
    t = np.array([1.5], np.float32)
    with wso.WSTessendorf(n, p.tile_length, max_slots=2) as ws:
        ws.PrepareWithGauss(xi)
        s = torch.cuda.Stream()
        ws.set_stream(s.cuda_stream)
        for _ in range(3):   # plain, capture of the library's own frame graph, graph launch
            ws.compute_batch(t, first_slot=0)
        ws.sync()
        assert ws.stats()["frame_graph_launches"] == 2
        d_ref, n_ref = ws.copy_map(0, 0), ws.copy_map(1, 0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            ws.compute_batch(t, first_slot=1)
        assert ws.stats()["frame_graph_launches"] == 2
        g.replay()
        torch.cuda.synchronize()
        assert ws.copy_map(0, 1).tobytes() == d_ref.tobytes() and ws.copy_map(1, 1).tobytes() == n_ref.tobytes()
        a_ref, dd, nn = o.compute_waves(1.5)
        assert_maps_close(ws.copy_map(0, 1), ws.copy_map(1, 1), dd, nn, "frame recorded in a caller's graph")
        ws.set_stream(None)


@pytest.mark.parametrize("n", [512, 1024, 2048])
def test_bulk_tilings_vs_oracle(wso, n):
    """The batched (Bulk) CTA tilings of the benchmarked sizes - K1's fused front end at 512^2 (NF4, radix-2 first
    stage) and 1024^2 (NF2, radix 4), the 1024-thread K1 at 2048^2, the paired W layout - against the oracle.  A launch
    of a few tile-frames takes the latency tilings instead, so single-frame tests do not reach these kernels at
    512^2 / 1024^2."""
    p, o, xi = _oracle_for(n)
    nframes = 9 if n <= 1024 else 3
    times = np.array([1.25 + 0.05 * i for i in range(nframes)], np.float32)
    with wso.WSTessendorf(n, p.tile_length, max_slots=nframes) as ws:
        ws.PrepareWithGauss(xi)
        ws.compute_batch(times)
        a, mn, mx = ws.read_heights(0, nframes)
        for i in (0, nframes // 2, nframes - 1):
            a_ref, d_ref, n_ref = o.compute_waves(float(times[i]))
            assert_maps_close(ws.copy_map(0, i), ws.copy_map(1, i), d_ref, n_ref, f"N={n} bulk frame {i}")
            assert abs(a[i] - a_ref) <= SCALAR_REL_TOL * a_ref
        # the single-frame (latency-tiling) path computes the same frame within the gate as well
        d_b, n_b = ws.copy_map(0, nframes - 1), ws.copy_map(1, nframes - 1)
        a1 = ws.ComputeWaves(float(times[-1]))
        assert abs(a1 - a[-1]) <= SCALAR_REL_TOL * a1
        assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), d_b, n_b, f"N={n} latency vs bulk tiling")


_KERNEL_SET_REFS = {}


@pytest.mark.parametrize("n,mask", [(512, 1), (512, 2), (512, 4), (512, 7), (1024, 0), (1024, 1), (1024, 2), (1024, 4),
                                    (1024, 7), (2048, 3), (2048, 4), (2048, 7),
                                    # bits 4 / 6: K1 / K2 of the CTA-per-line set in their persistent form
                                    (512, 0x10), (512, 0x50), (1024, 0x10), (1024, 0x40), (1024, 0x52), (2048, 0x40),
                                    (2048, 0)])
def test_both_kernel_sets_vs_oracle(wso, n, mask):
    """The warp-per-line kernels (wso_kernels2.cu: radix-32 register stages, shuffle exchanges, bulk-copy line pipeline)
    against the oracle, alone and mixed kernel by kernel with the CTA-per-line set (mask bit k = kernel k on the
    warp-per-line set; the two sets share the intermediate W layout).  Several frames per launch so that the persistent
    K2 / K2h walk more than one row item per group and more than one item per launch."""
    from watersurfacerendering_b200 import _lib
    nframes = 7 if n <= 1024 else 3
    times = np.array([0.75 + 0.35 * i for i in range(nframes)], np.float32)
    if n not in _KERNEL_SET_REFS:  # the oracle frames are the slow part: once per size
        p, o, xi = _oracle_for(n)
        _KERNEL_SET_REFS[n] = (p, xi, {i: o.compute_waves(float(times[i])) for i in (0, nframes // 2, nframes - 1)})
    p, xi, refs = _KERNEL_SET_REFS[n]
    L = _lib.load()
    assert L.wso_select_kernels(mask) == 0
    try:
        with wso.WSTessendorf(n, p.tile_length, max_slots=nframes) as ws:
            ws.PrepareWithGauss(xi)
            ws.compute_batch(times)
            a, mn, mx = ws.read_heights(0, nframes)
            for i in (0, nframes // 2, nframes - 1):
                a_ref, d_ref, n_ref = refs[i]
                assert_maps_close(ws.copy_map(0, i), ws.copy_map(1, i), d_ref, n_ref, f"N={n} mask {mask} frame {i}")
                assert abs(a[i] - a_ref) <= SCALAR_REL_TOL * a_ref
                assert a[i] == max(abs(mn[i]), abs(mx[i]))
    finally:
        L.wso_select_kernels(-1)


def test_maps_hold_reference_defaults_before_first_compute(wso):
    """reference: Prepare() resizes the maps with displacement (0,0,0,0) and normal (0,1,0,0) (WSTessendorf.cpp:48-54);
    a consumer that reads them before the first ComputeWaves() must see exactly that."""
    n = 64
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length, max_slots=3) as ws:
        ws.PrepareWithGauss(xi)
        for slot in range(3):
            d, nm = ws.copy_map(0, slot), ws.copy_map(1, slot)
            assert np.all(d == 0.0)
            assert np.all(nm[..., 1] == 1.0) and np.all(nm[..., [0, 2, 3]] == 0.0)
        d, nm = ws.GetDisplacements(), ws.GetNormals()
        assert np.all(d == 0.0) and np.all(nm[..., 1] == 1.0) and np.all(nm[..., [0, 2, 3]] == 0.0)


def test_compact_h0_and_async_compute(wso):
    """the 12-byte spectrum record (amp, dispersion; heightAmp_conj is redundant, WSTessendorf.cpp:132-135)
    round-trips bit for bit against the 20-byte one, and the asynchronous frame (event instead of a blocking call with a
    host copy) produces the same device maps as ComputeWaves."""
    n = 128
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length, max_slots=2) as ws:
        ws.PrepareWithGauss(xi)
        full = ws.ExportH0()
        comp = ws.ExportH0Compact()
        assert np.array_equal(comp[..., 0], full["re"]) and np.array_equal(comp[..., 1], full["im"])
        assert np.array_equal(comp[..., 2], full["omega"])
        a1 = ws.ComputeWaves(5.5)
        d1, n1 = ws.GetDisplacements().copy(), ws.GetNormals().copy()
        with wso.WSTessendorf(n, p.tile_length, max_slots=2) as ws2:
            ws2.ImportH0Compact(comp)
            assert ws2.ExportH0().tobytes() == full.tobytes()
            ev = ws2.ComputeWavesAsync(5.5)
            ws2.WaitEvent(ev)
            a2, mn2, mx2 = ws2.read_heights(0, 1)
            assert a2[0] == a1
            assert ws2.copy_map(0, 0).tobytes() == d1.tobytes() and ws2.copy_map(1, 0).tobytes() == n1.tobytes()


def test_independent_tiles_in_one_batch(wso):
    """BASELINE config 4 in miniature: tiles with different wind / seed, batched in one launch."""
    n, ntiles = 128, 5
    oracles = []
    with wso.WSTessendorf(n, 250.0, max_tiles=ntiles, max_slots=ntiles) as ws:
        for j in range(ntiles):
            ang = 2 * np.pi * j / ntiles
            p, o, xi = _oracle_for(n, L=250.0, seed=1000 + j, wind_x=float(np.cos(ang)),
                                   wind_y=float(np.sin(ang)), wind_speed=5.0 + 4.0 * j, lam=-1.0 + 0.3 * j)
            _apply_params(ws, p, tile=j)
            ws.PrepareWithGauss(xi, tile=j)
            assert ws.ExportH0(tile=j).tobytes() == o.h0.tobytes()
            oracles.append(o)
        ws.compute_batch(np.full(ntiles, 10.0, np.float32), tiles=np.arange(ntiles))
        a, mn, mx = ws.read_heights(0, ntiles)
        for j, o in enumerate(oracles):
            a_ref, d_ref, n_ref = o.compute_waves(10.0)
            assert_maps_close(ws.copy_map(0, j), ws.copy_map(1, j), d_ref, n_ref, f"tile {j}")
            assert abs(a[j] - a_ref) <= SCALAR_REL_TOL * a_ref


def test_compute_to_host_streams_every_frame(wso):
    n = 256
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length, max_slots=128) as ws:
        ws.PrepareWithGauss(xi)
        chunk = ws.stats()["chunk"]
        nfr = 2 * chunk + 3 if chunk < 40 else chunk + 3
        times = np.linspace(0.0, 50.0, nfr).astype(np.float32)
        disp = wso.PinnedBuffer((nfr, n, n, 4))
        norm = wso.PinnedBuffer((nfr, n, n, 4))
        a, mn, mx = ws.compute_to_host(times, disp.array, norm.array)
        for i in (0, chunk - 1, min(chunk, nfr - 1), nfr - 1):
            a1 = ws.ComputeWaves(float(times[i]))
            assert a1 == a[i]
            assert disp.array[i].tobytes() == ws.GetDisplacements().tobytes()
            assert norm.array[i].tobytes() == ws.GetNormals().tobytes()
        a_ref, d_ref, n_ref = o.compute_waves(float(times[-1]))
        assert_maps_close(disp.array[-1], norm.array[-1], d_ref, n_ref, "streamed last frame")
        disp.close()
        norm.close()


def test_tile_size_change_and_ignored_sizes(wso):
    """reference: SetTileSize ignores non-powers-of-two (WSTessendorf.cpp:459-468); a new size applies at Prepare."""
    with wso.WSTessendorf(64, 100.0) as ws:
        assert ws.SetTileSize(100) is False and ws.GetTileSize() == 64
        assert ws.SetTileSize(128) is True
        p, o, xi = _oracle_for(128, L=100.0)
        ws.PrepareWithGauss(xi)
        _check_frame(ws, o, 2.0, "resized to 128")


def test_error_codes(wso):
    from watersurfacerendering_b200 import _lib as L
    with wso.WSTessendorf(32, 50.0) as ws:
        with pytest.raises(L.WsoError) as e:
            ws.ComputeWaves(0.0)
        assert e.value.status == L.WSO_ERR_NOT_PREPARED
        h0 = np.zeros((32, 32), P.H0_DTYPE)
        h0[3, 4] = (0.5, 0.25, 0.5, 0.25, 1.0)   # conj amplitude is NOT the conjugate
        with pytest.raises(L.WsoError) as e:
            ws.ImportH0(h0)
        assert e.value.status == L.WSO_ERR_H0_NOT_CONJUGATE
        with pytest.raises(L.WsoError) as e:
            ws.compute_batch([0.0, 1.0])          # only one slot
        assert e.value.status in (L.WSO_ERR_INVALID_ARG, L.WSO_ERR_NOT_PREPARED)


def test_full_size_properties_2048(wso):
    """Size-independent properties at a BASELINE size: point symmetry, Parseval on the
    height field, disp.w == 1, |disp.y| <= 1 with the bound attained, and 24 texels against the direct
    float64 Fourier sum."""
    n = 2048
    p, o, xi = _oracle_for(n, seed=2)
    t = 0.05 * 41
    with wso.WSTessendorf(n, p.tile_length) as ws:
        ws.PrepareWithGauss(xi)
        a = float(ws.ComputeWaves(t))
        d = ws.GetDisplacements().copy()
        nm = ws.GetNormals().copy()
    h0 = o.h0
    ph = (h0["omega"] * np.float32(t)).astype(np.float32).astype(np.float64)
    H = 2.0 * (h0["re"].astype(np.float64) * np.cos(ph) - h0["im"].astype(np.float64) * np.sin(ph))

    def refl(x):
        return np.roll(x[::-1, ::-1], (1, 1), axis=(0, 1))

    assert np.all(d[..., 3] == 1.0)
    assert abs(np.abs(d[..., 1]).max() - 1.0) < 1e-6
    height = d[..., 1].astype(np.float64) * a
    assert np.abs(height - refl(height)).max() < 2e-5 * a
    for ch in (d[..., 0], d[..., 2], nm[..., 0], nm[..., 1]):
        ch = ch.astype(np.float64)
        assert np.abs(ch + refl(ch)).max() < 2e-5 * np.abs(ch).max()
    # Parseval: sum_x h^2 = N^2 sum_k |Y|^2 with Y the Hermitian part of the (real) spectrum
    Y = 0.5 * (H + refl(H))
    assert abs(np.sum(height ** 2) / (n * n * np.sum(Y ** 2)) - 1.0) < 1e-5
    # direct Fourier sum at random texels
    rng = np.random.default_rng(0)
    mm = np.arange(n)
    for _ in range(24):
        r, c = int(rng.integers(n)), int(rng.integers(n))
        er = np.exp(2j * np.pi * (mm * r % n) / n)
        ec = np.exp(2j * np.pi * (mm * c % n) / n)
        val = np.real(er @ H @ ec) * (-1.0 if (r + c) & 1 else 1.0)
        assert abs(height[r, c] - val) < 2e-5 * a, (r, c)


def test_maximum_size_properties_8192(wso):
    """Largest single-device grid (kMaxLogN = 13; first radix 2: paired W layout with four butterfly pairs per K2
    thread, 69 KB shared-memory lines).  The oracle would need minutes here, so the check is through size-independent
    properties on a spectrum built by the device-side Prepare(): disp.w == 1, |disp.y| <= 1 with the bound attained,
    point symmetry of every channel, Parseval on the height field and texels against the direct float64 Fourier sum."""
    n = 8192
    t = 7.75
    with wso.WSTessendorf(n, 1000.0 * n / 512) as ws:
        ws.PrepareCounterOnDevice(5)
        h0 = ws.ExportH0()
        a = float(ws.ComputeWaves(t))
        d = ws.GetDisplacements().copy()
        nm = ws.GetNormals().copy()
    assert np.isfinite(a) and a > 0
    ph = (h0["omega"] * np.float32(t)).astype(np.float32).astype(np.float64)
    H = 2.0 * (h0["re"].astype(np.float64) * np.cos(ph) - h0["im"].astype(np.float64) * np.sin(ph))
    del ph

    def refl(x):
        return np.roll(x[::-1, ::-1], (1, 1), axis=(0, 1))

    assert np.all(d[..., 3] == 1.0)
    assert abs(np.abs(d[..., 1]).max() - 1.0) < 1e-6
    height = d[..., 1].astype(np.float64) * a
    assert np.abs(height - refl(height)).max() < 4e-5 * a
    for ch in (d[..., 0], d[..., 2], nm[..., 0], nm[..., 1]):
        ch = ch.astype(np.float64)
        assert np.abs(ch + refl(ch)).max() < 4e-5 * np.abs(ch).max()
    for ch in (nm[..., 2], nm[..., 3]):
        ch = ch.astype(np.float64)
        assert np.abs(ch - refl(ch)).max() < 4e-5 * np.abs(ch).max()
    Y = 0.5 * (H + refl(H))
    assert abs(np.sum(height ** 2) / (float(n) * n * np.sum(Y ** 2)) - 1.0) < 1e-5
    del Y
    rng = np.random.default_rng(1)
    mm = np.arange(n)
    for _ in range(8):
        r, c = int(rng.integers(n)), int(rng.integers(n))
        er = np.exp(2j * np.pi * (mm * r % n) / n)
        ec = np.exp(2j * np.pi * (mm * c % n) / n)
        row = (er.real @ H) + 1j * (er.imag @ H)   # keeps H real: no 1 GB complex copy
        val = np.real(row @ ec) * (-1.0 if (r + c) & 1 else 1.0)
        assert abs(height[r, c] - val) < 4e-5 * a, (r, c)


@pytest.mark.parametrize("n", [64, 512, 1024])
def test_jacobian_channel(wso, n):
    """SURVEY row f-4 (the reference's COMPUTE_JACOBIAN switch, dead code there: parity unpinned, the oracle is the
    restatement PortOracle.jacobian): with the switch on displacement.w = J, every other channel stays within the gate
    of the reference maps; off again it is exactly 1.0f.  Single frames and a batch."""
    p, o, xi = _oracle_for(n)
    with wso.WSTessendorf(n, p.tile_length, max_slots=3) as ws:
        ws.PrepareWithGauss(xi)
        assert ws.GetComputeJacobian() is False
        ws.SetComputeJacobian(True)
        assert ws.GetComputeJacobian() is True
        times = np.array([0.75, 2.5, 31.0], np.float32)
        ws.compute_batch(times)
        a, _, _ = ws.read_heights(0, 3)
        for i in (0, 2):
            t = float(times[i])
            a_ref, d_ref, n_ref = o.compute_waves(t)
            d, nm = ws.copy_map(0, i), ws.copy_map(1, i)
            assert_maps_close(d, nm, d_ref, n_ref, f"N={n} jacobian t={t}", skip_w=True)
            assert abs(a[i] - a_ref) <= SCALAR_REL_TOL * a_ref
            j_ref = o.jacobian(t).astype(np.float64)
            j = d[..., 3].astype(np.float64)
            assert rel_l2(j, j_ref) <= 1e-5, f"N={n} J rel-L2 {rel_l2(j, j_ref):.3e}"
            assert np.abs(j - j_ref).max() <= 1e-4 * (j_ref.max() - j_ref.min())
        a1 = ws.ComputeWaves(2.5)   # single-frame entry point, same kernels
        assert np.array_equal(ws.GetDisplacements(), ws.copy_map(0, 1)) and a1 == a[1]
        ws.SetComputeJacobian(False)
        ws.ComputeWaves(2.5)
        assert np.all(ws.GetDisplacements()[..., 3] == 1.0)


def test_jacobian_channel_size_limit(wso):
    from watersurfacerendering_b200 import _lib as L
    with wso.WSTessendorf(8192, 16000.0) as ws:
        with pytest.raises(L.WsoError) as e:
            ws.SetComputeJacobian(True)
        assert e.value.status == L.WSO_ERR_INVALID_ARG
        assert ws.GetComputeJacobian() is False
