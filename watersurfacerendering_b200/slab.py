"""One large ocean grid over several B200s: slab-decomposed 2-D transform (BASELINE.json configs[4]).

One process per GPU (torchrun).  Rank r owns the column pairs (n, N-n), n in [r*N/2P, (r+1)*N/2P), for the first
transform (along m, fused with the spectrum evolve: kernel K1) and the row pairs (m', N-m') of the same range for
the second (along n, fused with the packing: K2h, K2).  In between sits the ONE exchange step of the path:

K1 stores its output coalesced; ONE transposing exchange kernel (``wso_slab_exchange``) carries the blocks to the owners:

  * ``fused=True``  — straight into the owners' receive buffers through NVLink peer mappings (CUDA IPC): the exchange IS
    the kernel's stores; a stream-ordered collective on a scratch word orders the ranks afterwards.  The receive buffers
    alternate by frame parity, so ``compute()`` may be called back to back without host synchronisation.
  * ``fused=False`` — into the blocks of a local send buffer that ``torch.distributed.all_to_all_single`` (NCCL) moves.

and one 2-float all-reduce, because the normalisation amplitude A = max(|min|,|max|) of the height field is global
(reference: WSTessendorf.cpp:440-455).  torch is plumbing here (process group, streams); every kernel is in
libwsocean.so.  ``SlabBackend`` is the seam the CPU tests use to drive the same orchestration over gloo with the
emulated kernel bodies (tests/emu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


class _DeviceView:
    """Expose a raw device allocation to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class SlabBackend:
    """Per-rank compute behind the orchestration: the C-ABI CUDA library."""

    def __init__(self, tile_size: int, tile_length: float, rank: int, world: int, device: int, **params):
        import torch
        self._lib = L.load()
        p = L.WsoParams()
        L.check(self._lib.wso_default_params(C.byref(p)))
        p.tile_size, p.tile_length = int(tile_size), float(tile_length)
        for k, v in params.items():
            setattr(p, k, v)
        self.params = p
        h = C.c_void_p()
        L.check_slab(self._lib.wso_slab_create(C.byref(p), int(device), int(rank), int(world), C.byref(h)))
        self._h = h
        self.n, self.rank, self.world, self.device = int(tile_size), int(rank), int(world), int(device)
        self.hl = self.n // 2 // self.world
        send, recv, mm, bb = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_size_t()
        L.check_slab(self._lib.wso_slab_buffers(h, C.byref(send), C.byref(recv), C.byref(bb), C.byref(mm)), h)
        self.block_bytes = int(bb.value)
        nfl = self.block_bytes // 4 * self.world
        dev = torch.device("cuda", self.device)
        self.send = torch.as_tensor(_DeviceView(send.value, nfl), device=dev)
        self.recv = torch.as_tensor(_DeviceView(recv.value, nfl), device=dev)
        self.minmax = torch.as_tensor(_DeviceView(mm.value, 2), device=dev)
        self.fused = False

    def close(self):
        if getattr(self, "_h", None):
            self.send = self.recv = self.minmax = None
            self._lib.wso_slab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- spectrum
    def import_h0(self, h0_full: np.ndarray):
        h0 = np.ascontiguousarray(h0_full)
        assert h0.nbytes == self.n * self.n * 20
        L.check_slab(self._lib.wso_slab_import_h0(self._h, h0.ctypes.data_as(C.c_void_p)), self._h)

    def prepare_counter(self, seed: int):
        L.check_slab(self._lib.wso_slab_prepare_counter(self._h, int(seed)), self._h)

    def prepare_counter_device(self, seed: int):
        """The same spectrum, built by the device Prepare kernel (this rank's column pairs only)."""
        L.check_slab(self._lib.wso_slab_prepare_counter_device(self._h, int(seed)), self._h)

    def set_lambda(self, lam: float):
        L.check_slab(self._lib.wso_slab_set_lambda(self._h, float(lam)), self._h)

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        L.check_slab(self._lib.wso_slab_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)), self._h)

    def force_pair(self, on: bool):
        L.check_slab(self._lib.wso_slab_force_pair(self._h, 1 if on else 0), self._h)

    # ---- fused exchange plumbing
    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        L.check_slab(self._lib.wso_slab_ipc_handle(self._h, buf), self._h)
        return buf.raw

    def open_peers(self, handles):
        for r, hd in enumerate(handles):
            buf = C.create_string_buffer(hd, 64) if r != self.rank else None
            L.check_slab(self._lib.wso_slab_open_peer(self._h, r, buf), self._h)
        L.check_slab(self._lib.wso_slab_set_fused(self._h, 1), self._h)
        self.fused = True

    # ---- phases (asynchronous on the slab's stream)
    def pass1(self, t: float):
        L.check_slab(self._lib.wso_slab_pass1(self._h, float(t)), self._h)

    def exchange(self):
        """The transposing exchange kernel (peer stores in fused mode, else it fills the send blocks)."""
        L.check_slab(self._lib.wso_slab_exchange(self._h), self._h)

    def pass1_fields(self, t: float, field0: int, nfields: int):
        L.check_slab(self._lib.wso_slab_pass1_fields(self._h, float(t), int(field0), int(nfields)), self._h)

    def exchange_fields(self, field0: int, nfields: int, cuda_stream_ptr: Optional[int] = None):
        L.check_slab(self._lib.wso_slab_exchange_fields(self._h, int(field0), int(nfields),
                                                        C.c_void_p(cuda_stream_ptr or 0)), self._h)

    def fields_per_group(self) -> int:
        return int(self._lib.wso_slab_fields_per_group(self._h))

    def heights(self):
        L.check_slab(self._lib.wso_slab_heights(self._h), self._h)

    def pass2(self):
        L.check_slab(self._lib.wso_slab_pass2(self._h), self._h)

    def sync(self):
        L.check_slab(self._lib.wso_slab_sync(self._h), self._h)

    def read_heights(self):
        a, mn, mx = C.c_float(), C.c_float(), C.c_float()
        L.check_slab(self._lib.wso_slab_read_heights(self._h, C.byref(a), C.byref(mn), C.byref(mx)), self._h)
        return np.float32(a.value), np.float32(mn.value), np.float32(mx.value)

    def local_rows(self, which: int) -> np.ndarray:
        out = np.empty((2 * self.hl, self.n, 4), np.float32)
        L.check_slab(self._lib.wso_slab_copy_rows(self._h, which, out.ctypes.data_as(C.c_void_p)), self._h)
        return out

    def row_index(self) -> np.ndarray:
        rows = np.zeros(2 * self.hl, np.uint32)
        L.check_slab(self._lib.wso_slab_row_index(self._h, rows.ctypes.data_as(C.c_void_p)), self._h)
        return rows


def counter_h0(params: "L.WsoParams", seed: int, m0: int, rows: int) -> np.ndarray:
    """Reference-layout h0 records of rows [m0, m0+rows) of the counter-generated spectrum (for checkers)."""
    from .surface import H0_DTYPE
    out = np.zeros((rows, int(params.tile_size)), H0_DTYPE)
    L.check(L.load().wso_counter_h0(C.byref(params), int(seed), int(m0), int(rows), out.ctypes.data_as(C.c_void_p)))
    return out


class SlabOcean:
    """ComputeWaves(t) of one N x N grid over the ranks of a torch.distributed process group."""

    def __init__(self, backend, group=None, fused: bool = False, pipeline: bool = False):
        import torch.distributed as dist
        self.b = backend
        self.group = group
        self.world = backend.world
        self._bound_stream = None
        self._xs = None
        self.pipeline = pipeline
        self.fused = bool(fused) and self.world > 1
        if self.world > 1 and not dist.is_initialized():
            raise RuntimeError("SlabOcean over several ranks needs an initialised torch.distributed process group")
        if self.fused:
            handles = [None] * self.world
            dist.all_gather_object(handles, backend.ipc_handle(), group=group)
            backend.open_peers(handles)
            dist.barrier(group=group)

    def compute(self, t: float):
        """Enqueue one tile-frame (asynchronous w.r.t. the host except for what the collectives impose)."""
        import torch
        import torch.distributed as dist
        b = self.b
        # the library's kernels and the collectives below must share one stream: bind the backend to torch's current
        # stream (a change of stream drains the old one first)
        cur = torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else None
        if cur is not None and cur != self._bound_stream and hasattr(b, "set_stream"):
            b.set_stream(cur)
            self._bound_stream = cur
        if self.fused and self.pipeline and hasattr(b, "pass1_fields") and b.fields_per_group() == 1 and b.hl >= 32:
            # field by field: the peer stores of field f (second stream) overlap the transform of field f + 1
            main = torch.cuda.current_stream()
            if self._xs is None:
                self._xs = torch.cuda.Stream(device=main.device)
                self._evs = [torch.cuda.Event() for _ in range(5)]
            for f in range(4):
                b.pass1_fields(t, f, 1)
                self._evs[f].record(main)
                self._xs.wait_event(self._evs[f])
                b.exchange_fields(f, 1, self._xs.cuda_stream)
            self._evs[4].record(self._xs)
            main.wait_event(self._evs[4])
        else:
            b.pass1(t)
            if hasattr(b, "exchange"):  # (the emulated CPU backend's pass1 fills the send blocks itself)
                b.exchange()
        if self.world > 1 and not self.fused:
            dist.all_to_all_single(b.recv, b.send, group=self.group)
        b_needs_order = self.world > 1 and self.fused
        if b_needs_order:
            # every rank's K1 (peer stores) must have completed before anyone transforms rows: a stream-ordered
            # collective on a scratch word does it
            dist.all_reduce(self._scratch(), group=self.group)
        b.heights()
        if self.world > 1:
            mm = b.minmax
            v = torch.stack((mm[0], -mm[1]))
            dist.all_reduce(v, op=dist.ReduceOp.MIN, group=self.group)
            mm[0] = v[0]
            mm[1] = -v[1]
        b.pass2()

    def _scratch(self):
        import torch
        if not hasattr(self, "_scr"):
            self._scr = torch.zeros(1, dtype=torch.float32, device=self.b.send.device)
        return self._scr

    def gather_maps(self):
        """(disp, norm) as full (N,N,4) arrays on every rank - verification helper, not a hot path."""
        import torch.distributed as dist
        b = self.b
        n = b.n
        out = []
        rows = b.row_index()
        for which in (L.WSO_MAP_DISPLACEMENT, L.WSO_MAP_NORMAL):
            loc = b.local_rows(which)
            if self.world > 1:
                parts = [None] * self.world
                dist.all_gather_object(parts, (rows, loc), group=self.group)
            else:
                parts = [(rows, loc)]
            full = np.zeros((n, n, 4), np.float32)
            for r_rows, r_loc in parts:
                full[r_rows] = r_loc
            out.append(full)
        return out[0], out[1]
