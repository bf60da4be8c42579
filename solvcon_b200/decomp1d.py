"""Slab decomposition of the 1D scalar march (``spacetime.LinearScalarSolver`` / ``InviscidBurgersSolver``) over
the GPUs of one box (BASELINE.json configs[1]).

The reference has no decomposition; its grid is periodic (``treat_boundary_so0/so1``, core.hpp:982-1004: the two
odd-plane ghost solution elements are copies of the first / last interior odd-plane element).  A slab owns a
contiguous range of conservation elements ``[c0, c1)`` of the global grid and keeps the reference's local layout
(``nx = 2 n + 5`` slots with two ghost coordinates per side); its odd-plane ghosts are filled by its ring
neighbours instead of the periodic copy:

    half 1 (even -> odd plane)  ->  pack {so0, so1} of the first / last odd-plane point  ->  send left / right,
    receive  ->  unpack (also evaluates the ghosts' cfl like ``update_cfl`` does)  ->  half 2 (odd -> even plane)

The even-plane point on a slab boundary is computed by both neighbours from the same operands, so the march is
bit-identical to the undecomposed one for any number of slabs (``tests/test_gpu_1d.py``, ``tests/test_decomp_gloo.py``).
The message is 2 doubles per side per step: latency only, never bandwidth.

Transport: :class:`LocalRing` (all slabs in one process on the library stream) or :class:`TorchRing`
(``torch.distributed`` send/recv of the 4-double device buffers: NCCL between GPUs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .spacetime import Grid, InviscidBurgersSolver, LinearScalarSolver


def slab_ranges(ncelm, nparts):
    """Contiguous, nearly equal CE ranges ``[(c0, c1), ...]`` of a grid of ``ncelm`` conservation elements."""
    cuts = [(ncelm * r) // nparts for r in range(nparts + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(nparts)]


class ScalarSlab:
    """One slab: the device solver of the local CE range plus the 2+2 double halo buffers.

    ``xcoord_global``: the coordinate array of the GLOBAL grid (``Grid.xcoord``, ``2 ncelm + 5`` entries; only the
    slice ``[2 c0, 2 c1 + 5)`` is read, so a caller may pass any array-like that supports that slice).  The local
    ghost coordinates are the global neighbours' coordinates (mirrored ghosts at the two global ends, exactly the
    ones the reference builds there)."""

    def __init__(self, eq, xcoord_global, c0, c1, time_increment, device_buffers=None):
        n = int(c1 - c0)
        if n < 1:
            raise ValueError("ScalarSlab: empty slab")
        xs = np.ascontiguousarray(xcoord_global[2 * c0:2 * c1 + 5], dtype=np.float64)
        grid = Grid(xloc=xs[2:2 * n + 3:2])
        grid.xcoord.ndarray[...] = xs       # interior entries are identical; the ghosts become the neighbours' points
        self.c0, self.c1, self.n = int(c0), int(c1), n
        cls = InviscidBurgersSolver if eq else LinearScalarSolver
        self.svr = cls(grid, time_increment)
        self._lib = _lib.load()
        _lib.check(self._lib.scg_st_set_halo_mode(self.svr._h, 1))
        if device_buffers is None:
            p = C.c_void_p()
            _lib.check(self._lib.scg_dev_alloc(C.byref(p), 8 * 8))
            self._own = p.value
            device_buffers = (p.value, p.value + 32)
        else:
            self._own = None
        self.send_ptr, self.recv_ptr = device_buffers   # 4 doubles each: [left0, left1, right0, right1]

    def __del__(self):
        if getattr(self, "_own", None):
            self._lib.scg_dev_free(C.c_void_p(self._own))
            self._own = None

    # -- state in/out (global slices) ---------------------------------------------------------------
    def load_global(self, so0, so1, cfl=None):
        """Take this slab's part of global ``nx``-slot arrays (interior and ghost slots alike)."""
        sl = slice(2 * self.c0, 2 * self.c1 + 5)
        self.svr.so0.ndarray[:, 0] = so0[sl]
        self.svr.so1.ndarray[:, 0] = so1[sl]
        if cfl is not None:
            self.svr.cfl.ndarray[...] = cfl[sl]
        self.svr.resident(True)

    def fetch(self):
        self.svr.sync_to_host()
        return self.svr.so0.ndarray[:, 0], self.svr.so1.ndarray[:, 0], self.svr.cfl.ndarray

    # -- march pieces --------------------------------------------------------------------------------
    def half(self, which, alpha):
        _lib.check(self._lib.scg_st_march_half(self.svr._h, int(which), int(alpha)))

    def pack(self):
        _lib.check(self._lib.scg_st_halo_pack(self.svr._h, 1, C.c_void_p(self.send_ptr), C.c_void_p(self.send_ptr + 16)))

    def unpack(self, from_left_ptr, from_right_ptr):
        _lib.check(self._lib.scg_st_halo_unpack(self.svr._h, 1, C.c_void_p(from_left_ptr), C.c_void_p(from_right_ptr)))


class LocalRing:
    """All slabs in one process (tests, one GPU): everything runs in order on the library stream."""

    def __init__(self, slabs):
        self.slabs = list(slabs)

    def exchange(self):
        P = len(self.slabs)
        for s in self.slabs:
            s.pack()
        for r, s in enumerate(self.slabs):
            left, right = self.slabs[(r - 1) % P], self.slabs[(r + 1) % P]
            # my left ghost is the left neighbour's LAST odd-plane point (its "right" message) and vice versa
            s.unpack(left.send_ptr + 16, right.send_ptr)

    def march(self, alpha, steps):
        for _ in range(int(steps)):
            for s in self.slabs:
                s.half(1, alpha)
            self.exchange()
            for s in self.slabs:
                s.half(2, alpha)


class NcclRing:
    """One slab per process over the LIBRARY's communicator (``scg_comm_init``): the whole ring march is one C call
    (``scg_st_march_ring``: half 1, ncclSend/ncclRecv of the two boundary points on the library's communication stream,
    half 2); no torch on this path."""

    def __init__(self, eq, xcoord_global, c0, c1, time_increment, rank, world):
        self.rank, self.world = int(rank), int(world)
        self.slab = ScalarSlab(eq, xcoord_global, c0, c1, time_increment)

    def march(self, alpha, steps):
        P, r = self.world, self.rank
        _lib.check(_lib.load().scg_st_march_ring(self.slab.svr._h, int(alpha), int(steps), (r - 1) % P, (r + 1) % P))


class TorchRing:
    """One slab per process: ``torch.distributed`` P2P of the two 2-double messages with the ring neighbours."""

    def __init__(self, eq, xcoord_global, c0, c1, time_increment, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        self.sendbuf = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.recvbuf = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.slab = ScalarSlab(eq, xcoord_global, c0, c1, time_increment,
                               device_buffers=(self.sendbuf.data_ptr(), self.recvbuf.data_ptr()))
        self._lib_stream = None

    def exchange(self):
        torch, dist, s = self.torch, self.dist, self.slab
        s.pack()
        P, r = self.world, self.rank
        if 1 == P:
            _lib.sync()
            self.recvbuf[0:2] = self.sendbuf[2:4]
            self.recvbuf[2:4] = self.sendbuf[0:2]
            torch.cuda.synchronize(self.device)
        else:
            if self._lib_stream is None:
                self._lib_stream = torch.cuda.ExternalStream(int(_lib.load().scg_stream_handle()), device=self.device)
            left, right = (r - 1) % P, (r + 1) % P
            # with two ranks both messages go to the same peer: the orders on the two sides must mirror each other
            ops = [dist.P2POp(dist.isend, self.sendbuf[0:2], left), dist.P2POp(dist.isend, self.sendbuf[2:4], right),
                   dist.P2POp(dist.irecv, self.recvbuf[2:4], right), dist.P2POp(dist.irecv, self.recvbuf[0:2], left)]
            with torch.cuda.stream(self._lib_stream):   # NCCL orders itself against the library stream: no host sync
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        s.unpack(self.recvbuf.data_ptr(), self.recvbuf.data_ptr() + 16)

    def march(self, alpha, steps):
        for _ in range(int(steps)):
            self.slab.half(1, alpha)
            self.exchange()
            self.slab.half(2, alpha)
