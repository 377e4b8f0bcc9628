"""Slab-decomposed path (BASELINE configs[4]).

CPU (-m "not gpu"): the real SlabOcean orchestration (all-to-all transpose + min/max all-reduce) over world_size-2
gloo, with the emulated kernel bodies (tests/emu) as the per-rank backend, against the reference fixture.
GPU (-m gpu): the CUDA slab kernels through the C ABI - world 1 in-process (incl. the two-CTA cluster K2), and
2 ranks under torchrun when the box has >= 2 GPUs - against the oracle at N = 2048 and the fixtures at N = 64/256.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, SCALAR_REL_TOL, assert_maps_close, h0_struct, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _emu_worker(rank, world, port, name, variant, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import driver as E
    from watersurfacerendering_b200.slab import SlabOcean
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, params = load_golden(name)
    h0 = g["h0"]
    i = len(g["t"]) - 1
    b = E.EmuSlabBackend(params["tile_size"], params["tile_length"], params["lam"], h0[..., 0], h0[..., 1], h0[..., 4],
                         rank, world, variant=variant, anim_period=params["anim_period"])
    ocean = SlabOcean(b, fused=False)
    ocean.compute(float(g["t"][i]))
    disp, norm = ocean.gather_maps()
    a, mn, mx = b.read_heights()
    ok = True
    try:
        assert_maps_close(disp, norm, g["disp"][i], g["norm"][i], f"{name} rank {rank}")
        assert abs(a - g["A"][i]) <= SCALAR_REL_TOL * g["A"][i]
        assert abs(mn - g["minh"][i]) <= SCALAR_REL_TOL * g["A"][i]
        assert abs(mx - g["maxh"][i]) <= SCALAR_REL_TOL * g["A"][i]
        msg = ""
    except AssertionError as ex:  # report through the queue: a spawned process cannot fail the test by itself
        ok, msg = False, str(ex)
    q.put((rank, ok, msg, float(a)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,variant", [("n64_default", 0), ("n64_wind", 1)])
def test_two_rank_gloo_slab_exchange(name, variant):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emu_worker, args=(r, world, port, name, variant, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, msg, a in res:
        assert ok, f"rank {rank}: {msg}"
    assert res[0][3] == res[1][3]   # both ranks agree on the global amplitude


def test_counter_gauss_is_deterministic_and_normal():
    """The counter-based Gaussian source behind wso_slab_prepare_counter: same (seed, index) -> same draw on every rank,
    moments of a standard normal, and h0 conjugate symmetry as the kernels require."""
    import watersurfacerendering_b200 as W
    from watersurfacerendering_b200 import _lib as L
    from watersurfacerendering_b200.slab import counter_h0
    import ctypes as C
    p = L.WsoParams()
    L.load().wso_default_params(C.byref(p))
    p.tile_size, p.tile_length = 256, 500.0
    a = counter_h0(p, 7, 0, 256)
    b = counter_h0(p, 7, 64, 32)
    assert a[64:96].tobytes() == b.tobytes()
    c = counter_h0(p, 8, 64, 32)
    assert c.tobytes() != b.tobytes()
    assert np.array_equal(a["re"], a["re_c"]) and np.array_equal(a["im"], -a["im_c"])
    assert a["omega"][128, 128] == 0 and a["re"][128, 128] == 0   # DC bin (reference: k = 0 at index N/2)
    # xi recovered from a flat spectrum: unit variance, zero mean, uncorrelated parts
    p.phillips_const = 1.0
    amp = counter_h0(p, 7, 0, 256)
    z = amp["re"] + 1j * amp["im"]
    mask = np.abs(z) > 0
    ph = np.angle(z[mask])
    assert abs(np.mean(np.cos(ph))) < 0.02 and abs(np.mean(np.sin(ph))) < 0.02


# --------------------------------------------------------------------------------------------------- GPU
def _oracle(n, seed=0):
    from oracle import port as P
    p = P.OceanParams(tile_size=n, tile_length=1000.0 * n / 512)
    o = P.PortOracle(p)
    rng = np.random.default_rng(seed + n)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    o.prepare(xi)
    return p, o


@pytest.mark.gpu
@pytest.mark.parametrize("name,pair", [("n64_default", False), ("n64_wind", True), ("n256_default", False),
                                       ("n256_default", True)])
def test_slab_world1_fixtures(name, pair):
    from watersurfacerendering_b200.slab import SlabBackend, SlabOcean
    g, params = load_golden(name)
    b = SlabBackend(params["tile_size"], params["tile_length"], 0, 1, 0, lambda_=params["lam"],
                    anim_period=params["anim_period"], wind_dir_x=params["wind_x"], wind_dir_y=params["wind_y"],
                    wind_speed=params["wind_speed"], phillips_const=params["phillips_const"], damping=params["damping"])
    try:
        b.force_pair(pair)
        b.import_h0(h0_struct(g["h0"]))
        ocean = SlabOcean(b)
        for i, t in enumerate(g["t"]):
            ocean.compute(float(t))
            b.sync()
            disp, norm = ocean.gather_maps()
            assert_maps_close(disp, norm, g["disp"][i], g["norm"][i], f"{name} pair={pair} t={t}")
            a, mn, mx = b.read_heights()
            assert abs(a - g["A"][i]) <= SCALAR_REL_TOL * g["A"][i]
            assert abs(mn - g["minh"][i]) <= SCALAR_REL_TOL * g["A"][i]
            assert abs(mx - g["maxh"][i]) <= SCALAR_REL_TOL * g["A"][i]
    finally:
        b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pair", [False, True])
def test_slab_world1_2048_vs_oracle_and_batched_path(pair):
    """N = 2048 through the slab kernels on one device: vs the oracle (tolerance gate) and vs the regular batched
    path (bit-exact: same arithmetic, different data placement)."""
    import watersurfacerendering_b200 as W
    from watersurfacerendering_b200.slab import SlabBackend, SlabOcean
    n = 2048
    p, o = _oracle(n)
    t = 12.5
    a_ref, d_ref, n_ref = o.compute_waves(t)
    b = SlabBackend(n, p.tile_length, 0, 1, 0)
    try:
        b.force_pair(pair)
        b.import_h0(o.h0)
        ocean = SlabOcean(b)
        ocean.compute(t)
        b.sync()
        disp, norm = ocean.gather_maps()
        a, mn, mx = b.read_heights()
    finally:
        b.close()
    assert_maps_close(disp, norm, d_ref, n_ref, f"slab 2048 pair={pair}")
    assert abs(a - a_ref) <= SCALAR_REL_TOL * a_ref
    with W.WSTessendorf(n, p.tile_length) as ws:
        ws.ImportH0(o.h0)
        a2 = ws.ComputeWaves(t)
        assert a2 == a
        assert ws.GetDisplacements().tobytes() == disp.tobytes()
        assert ws.GetNormals().tobytes() == norm.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("fused", [0, 1, 3])
def test_slab_multi_gpu_vs_oracle(fused, world, tmp_path):
    """2 / 4 / 8 ranks under torchrun (as many as the box has): exchange kernel + NCCL all-to-all, and exchange kernel with
    direct peer stores over NVLink (fused 1; 3 = issued field by field on a second stream behind K1); N = 2048 vs the oracle,
    including three frames enqueued without host syncs."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    out = tmp_path / "res.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_worker.py"), "--size", "2048",
           "--fused", str(min(fused, 1)), "--pipeline", str(1 if fused == 3 else 0), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert out.read_text().strip() == "ok", out.read_text()


# --------------------------------------------------------------------------------------------------- 16384^2
def _huge_ok():
    if os.environ.get("WSO_SKIP_HUGE") == "1":
        return "WSO_SKIP_HUGE=1"
    try:
        import psutil
        if psutil.virtual_memory().available < 28 << 30:
            return "needs ~28 GB of host memory"
    except Exception:
        pass
    if torch.cuda.is_available() and torch.cuda.mem_get_info(0)[0] < 40 << 30:
        return "needs ~40 GB of device memory"
    return None


@pytest.mark.gpu
def test_slab_16384_sampled_texels_symmetry_parseval():
    """BASELINE configs[4] at FULL size (16384^2, the slab kernels incl. the two-CTA cluster K2 and the exchange kernel, one
    rank): SURVEY 8(d)'s size-independent checks -
      (i)   64 output texels (8 rows x 8 columns incl. 0, N/2, N-1) of height, Dx and slope-x against a direct float64
            evaluation of the Fourier sum  out[m'][n'] = (-1)^(m'+n') Re sum_{m,n} X[m][n] e^{+2 pi i (m m' + n n')/N}
            (reference: WSTessendorf.cpp:292-336 for X, :338-378 + :385-437 for the transform, sign and packing);
      (ii)  the point symmetry of SURVEY 4-7: height even, Dx / Dz / slopes odd under (m,n) -> (-m,-n) mod N, bit for bit;
      (iii) Parseval on the height field: sum out^2 = N^2 sum |even(X)|^2, even(X)[k] = (X[k] + conj X[-k]) / 2.
    The spectrum is the counter-based one (identical on host and device, wso_counter_h0)."""
    why = _huge_ok()
    if why:
        pytest.skip(why)
    from watersurfacerendering_b200 import _lib as L
    from watersurfacerendering_b200.slab import SlabBackend, SlabOcean, counter_h0
    n, seed, t = 16384, 7, np.float32(10.0)
    length = 32000.0
    b = SlabBackend(n, length, 0, 1, 0)
    try:
        b.prepare_counter_device(seed)
        ocean = SlabOcean(b)
        ocean.compute(float(t))
        b.sync()
        amp, mn, mx = b.read_heights()
        rows = b.row_index().astype(np.int64)
        inv_rows = np.empty(n, np.int64)
        inv_rows[rows] = np.arange(n)
        disp = b.local_rows(L.WSO_MAP_DISPLACEMENT)   # [local row][n][4], local row r = global row rows[r]
        norm = b.local_rows(L.WSO_MAP_NORMAL)
        lam = float(b.params.lambda_) if hasattr(b.params, "lambda_") else -1.0
    finally:
        b.close()
    assert np.isfinite(amp) and amp > 0 and amp == max(abs(mn), abs(mx))
    # (ii) point symmetry, bit for bit (row N-m of global row m; column N-n)
    gm = np.array([0, 1, 5, 4095, 8191, 8192, 8193, 12000, 16383])
    for m in gm:
        ra, rb = disp[inv_rows[m]], disp[inv_rows[(n - m) % n]]
        na, nb = norm[inv_rows[m]], norm[inv_rows[(n - m) % n]]
        mirror = (n - np.arange(n)) % n
        assert np.array_equal(ra[:, 1], rb[mirror, 1]), f"height not even at row {m}"
        assert np.array_equal(ra[:, 0], -rb[mirror, 0]) and np.array_equal(ra[:, 2], -rb[mirror, 2]), f"Dx/Dz not odd at row {m}"
        assert np.array_equal(na[:, 0], -nb[mirror, 0]) and np.array_equal(na[:, 1], -nb[mirror, 1]), f"slopes not odd at row {m}"
        assert np.array_equal(na[:, 2], nb[mirror, 2]) and np.array_equal(na[:, 3], nb[mirror, 3])
    assert np.all(disp[..., 3] == 1.0)
    # host side: X = h~(k, t) row chunk by row chunk in float64 (phase = the reference's single fp32 product omega*t)
    pr = L.WsoParams()
    import ctypes
    L.load().wso_default_params(ctypes.byref(pr))
    pr.tile_size, pr.tile_length = n, length
    idx = np.arange(n, dtype=np.float32)
    kv = (np.pi * (np.float32(2) * idx - np.float32(n)).astype(np.float64) / np.float64(np.float32(length))).astype(np.float32).astype(np.float64)
    cols = np.array([0, 1, 37, 4096, 8191, 8192, 12345, 16383])
    trow = np.array([0, 3, 64, 5000, 8192, 8193, 11111, 16383])
    En = np.exp(2j * np.pi * np.outer(np.arange(n), cols) / n)           # [n][8]
    acc = {k: np.zeros((len(trow), len(cols)), np.complex128) for k in ("h", "dx", "sx")}
    Xall = np.empty((n, n), np.complex64)
    chunk = 256
    for m0 in range(0, n, chunk):
        h0 = counter_h0(pr, seed, m0, chunk)
        ph = (h0["omega"] * t).astype(np.float64)                          # fl32(omega * t), WSTessendorf.h:267
        c, s = np.cos(ph), np.sin(ph)
        a = h0["re"].astype(np.float64) + 1j * h0["im"].astype(np.float64)
        ac = h0["re_c"].astype(np.float64) + 1j * h0["im_c"].astype(np.float64)
        X = a * (c + 1j * s) + ac * (c - 1j * s)                            # WaveHeightFT, WSTessendorf.h:265-275
        Xall[m0:m0 + chunk] = X.astype(np.complex64)
        kx = kv[None, :]
        kz = kv[m0:m0 + chunk, None]
        kl = np.sqrt(kx * kx + kz * kz)
        ux = np.where(kl > 1e-5, kx / np.where(kl > 1e-5, kl, 1.0), 0.0)
        Em = np.exp(2j * np.pi * np.outer(trow, np.arange(m0, m0 + chunk)) / n)   # [8][chunk]
        acc["h"] += Em @ (X @ En)
        acc["dx"] += Em @ ((-1j * ux * X) @ En)                           # cpp:316-323
        acc["sx"] += Em @ ((1j * kx * X) @ En)                            # cpp:304-307
    sign = np.where((trow[:, None] + cols[None, :]) % 2 == 0, 1.0, -1.0)
    ref_h, ref_dx, ref_sx = sign * acc["h"].real, sign * acc["dx"].real, sign * acc["sx"].real
    got = disp[inv_rows[trow]][:, cols, :]
    gnorm = norm[inv_rows[trow]][:, cols, :]
    # the parity gate's max-abs tolerance: 1e-4 x the channel's range over the whole map
    rng_h = float(mx) - float(mn)
    rng_dx = float(disp[..., 0].max()) - float(disp[..., 0].min())
    rng_sx = float(norm[..., 0].max()) - float(norm[..., 0].min())
    assert np.abs(got[..., 1].astype(np.float64) * float(amp) - ref_h).max() <= 1e-4 * rng_h
    assert np.abs(got[..., 0].astype(np.float64) - lam * ref_dx).max() <= 1e-4 * rng_dx
    assert np.abs(gnorm[..., 0].astype(np.float64) - ref_sx).max() <= 1e-4 * rng_sx
    # (iii) Parseval on the height field
    Xm = Xall[(n - np.arange(n)) % n][:, (n - np.arange(n)) % n]
    even_sq = 0.0
    for m0 in range(0, n, 1024):
        y = 0.5 * (Xall[m0:m0 + 1024].astype(np.complex128) + np.conj(Xm[m0:m0 + 1024].astype(np.complex128)))
        even_sq += float(np.sum(y.real ** 2 + y.imag ** 2))
    out_sq = float(np.sum((disp[..., 1].astype(np.float64) * float(amp)) ** 2))
    assert abs(out_sq - float(n) * n * even_sq) <= 2e-4 * out_sq, (out_sq, float(n) * n * even_sq)
