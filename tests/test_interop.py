"""External-memory interop of the map arrays: the library exports the memory K2 writes as a POSIX
file descriptor (what a Vulkan device imports with VK_KHR_external_memory_fd) or writes into memory imported from
another API.  Replaces the staging memcpy + upload of the reference (scene/WaterSurfaceMesh.cpp:701-755,
vulkan/Texture2D.cpp:175-226).  There is no Vulkan loader in the image, so the consumer side is played by the CUDA
driver API (cuda-python): it imports the exported fd as a foreign process would and reads the texels back."""
import os

import numpy as np
import pytest

import oracle.port as P
from conftest import SCALAR_REL_TOL, assert_maps_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wso():
    import watersurfacerendering_b200 as w
    return w


def _drv():
    from cuda.bindings import driver
    return driver


def _ok(res):
    err, *rest = res
    assert int(err) == 0, f"CUDA driver call failed: {err}"
    return rest[0] if len(rest) == 1 else rest


def _read_through_fd(fd: int, nbytes: int, device: int = 0) -> np.ndarray:
    """Import a shareable handle, map it and copy its bytes to the host (the 'other API' side of the hand-over)."""
    drv = _drv()
    _ok(drv.cuInit(0))
    ctx = _ok(drv.cuDevicePrimaryCtxRetain(device))
    _ok(drv.cuCtxPushCurrent(ctx))
    try:
        handle = _ok(drv.cuMemImportFromShareableHandle(fd, drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR))
        va = _ok(drv.cuMemAddressReserve(nbytes, 0, 0, 0))
        _ok(drv.cuMemMap(va, nbytes, 0, handle, 0))
        acc = drv.CUmemAccessDesc()
        acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = device
        acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READ
        _ok(drv.cuMemSetAccess(va, nbytes, [acc], 1))
        out = np.empty(nbytes, np.uint8)
        _ok(drv.cuMemcpyDtoH(out.ctypes.data, va, nbytes))
        _ok(drv.cuMemUnmap(va, nbytes))
        _ok(drv.cuMemAddressFree(va, nbytes))
        _ok(drv.cuMemRelease(handle))
        return out
    finally:
        drv.cuCtxPopCurrent()
        drv.cuDevicePrimaryCtxRelease(device)


def _oracle(n, seed=11):
    p = P.OceanParams(tile_size=n, tile_length=1000.0 * n / 512)
    o = P.PortOracle(p)
    rng = np.random.default_rng(seed + n)
    xi = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    o.prepare(xi)
    return p, o, xi


def test_export_requires_opt_in(wso):
    from watersurfacerendering_b200 import _lib as L
    with wso.WSTessendorf(64, 125.0) as ws:
        with pytest.raises(L.WsoError):
            ws.export_fd(0)


@pytest.mark.parametrize("n,slots", [(64, 1), (256, 3)])
def test_exported_fd_shows_the_maps_k2_wrote(wso, n, slots):
    p, o, xi = _oracle(n)
    with wso.WSTessendorf(n, p.tile_length, max_slots=slots) as ws:
        ws.PrepareWithGauss(xi)
        ws.set_exportable(True)          # after Prepare: the spectrum must survive the re-allocation of the maps
        times = [0.5 + 2.25 * i for i in range(slots)]
        ws.compute_batch(times)
        ws.sync()
        for which in (0, 1):
            fd, nbytes = ws.export_fd(which)
            assert fd >= 0 and nbytes >= slots * n * n * 16
            try:
                raw = _read_through_fd(fd, nbytes)
            finally:
                os.close(fd)
            seen = raw[: slots * n * n * 16].view(np.float32).reshape(slots, n, n, 4)
            for s in range(slots):
                # bit-identical to the library's own accessor: same memory
                assert seen[s].tobytes() == ws.copy_map(which, s).tobytes()
        # and the contents are the reference's maps
        for s, t in enumerate(times):
            a_ref, d_ref, n_ref = o.compute_waves(t)
            assert_maps_close(ws.copy_map(0, s), ws.copy_map(1, s), d_ref, n_ref, f"exportable N={n} slot {s}")
        a, _, _ = ws.read_heights(0, slots)
        assert abs(a[0] - o.compute_waves(times[0])[0]) <= SCALAR_REL_TOL * a[0]


def test_exportable_survives_tile_size_change_and_can_be_switched_off(wso):
    p, o, xi = _oracle(128)
    with wso.WSTessendorf(64, 125.0) as ws:
        ws.set_exportable(True)
        ws.SetTileSize(128)
        ws.SetTileLength(p.tile_length)
        ws.PrepareWithGauss(xi)
        a = ws.ComputeWaves(1.0)
        fd, nbytes = ws.export_fd(0)
        try:
            seen = _read_through_fd(fd, nbytes)[: 128 * 128 * 16].view(np.float32).reshape(128, 128, 4)
        finally:
            os.close(fd)
        assert seen.tobytes() == ws.copy_map(0, 0).tobytes()
        a_ref, d_ref, n_ref = o.compute_waves(1.0)
        assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), d_ref, n_ref, "after size change")
        assert abs(a - a_ref) <= SCALAR_REL_TOL * a_ref
        ws.set_exportable(False)
        from watersurfacerendering_b200 import _lib as L
        with pytest.raises(L.WsoError):
            ws.export_fd(0)
        ws.ComputeWaves(1.0)
        assert_maps_close(ws.GetDisplacements(), ws.GetNormals(), d_ref, n_ref, "back on cudaMalloc")


def test_import_of_foreign_memory(wso):
    """wso_import_external_fd: K2 writes into memory owned by someone else.  The 'someone else' here is a second
    context's exportable allocation; a driver that only accepts Vulkan/other-API handles for
    cudaImportExternalMemory reports it and the test is skipped (the Vulkan hand-over itself cannot run in this image)."""
    from watersurfacerendering_b200 import _lib as L
    n = 128
    p, o, xi = _oracle(n)
    with wso.WSTessendorf(n, p.tile_length) as owner, wso.WSTessendorf(n, p.tile_length) as ws:
        owner.set_exportable(True)
        fd, nbytes = owner.export_fd(0)
        try:
            ws.import_external_fd(0, fd, nbytes, 0)
        except L.WsoError as e:
            os.close(fd)
            pytest.skip(f"driver does not import a CUDA-exported fd as external memory: {e}")
        ws.PrepareWithGauss(xi)
        ws.ComputeWaves(2.0)
        # the owner sees the displacement map the second context computed
        a_ref, d_ref, n_ref = o.compute_waves(2.0)
        assert_maps_close(owner.copy_map(0, 0), ws.copy_map(1, 0), d_ref, n_ref, "imported memory")


def test_bad_arguments(wso):
    from watersurfacerendering_b200 import _lib as L
    with wso.WSTessendorf(64, 125.0) as ws:
        with pytest.raises(L.WsoError):
            ws.import_external_fd(0, 0, 16, 0)        # far too small
        with pytest.raises(L.WsoError):
            ws.signal_semaphore(0, 1)                 # nothing imported
        with pytest.raises(L.WsoError):
            ws.export_fd(7)
