"""Domain decomposition and halo exchange for the multidim march (one process per GPU).

The reference has no decomposition; this is the B200-side addition that lets one mesh
run on the 8 GPUs of a box.  The march is a pure gather, so a partition only needs, besides its own
cells, one layer of *halo copies* of the remote face neighbours (refreshed every substep) and a second,
passive layer behind it:

* a partition is an ordinary ``StaticMesh`` built from the owned cells plus that halo layer, in
  ascending GLOBAL cell order (so the owner cell of every face -- and with it the node order of the
  face, the sub-face order and every rounding -- is the same as in the undecomposed mesh);
* the second layer is never exchanged or read by an owned cell's update; it only exists so that every
  face of a first-layer copy has both of its cells present -- then the face keeps its global owner and
  node order, and the copy's mesh centroid ``clcnd`` (which the CE geometry of the owned interface cells
  reads, ``euler_ce.cpp:96-101,214-216``) comes out with the same bits as in the undecomposed mesh;
* the march kernels skip all copies (``owned`` mask, ``scg_mesh_create_part``);
* after every substep each partition sends the new state (``so0n``+``so1n``) of the cells its
  neighbours hold as halo copies (``scg_md_halo_pack/unpack``), in ascending global cell order on both
  sides; ``so0t`` of a halo copy is recomputed locally;
* once at setup the halo copies receive their solution point and gamma from the owner (a halo copy's
  own CE centroid cannot be computed locally: its outer neighbours are missing).

Results are bit-identical to the single-domain run for any number of parts
(``tests/test_gpu_decomp.py``, ``tests/test_decomp_gloo.py``).

Transport: :class:`NcclExchange` (the product path: the library's own communicator, ``csrc/scg_comm.cu`` -- ncclSend /
ncclRecv issued from C++ on the library's communication stream, the whole overlapped march one C call, no torch),
:class:`TorchExchange` (``torch.distributed`` isend/irecv of the packed buffers: NCCL between GPUs, gloo for the
CPU tests of the exchange logic) or :class:`LocalExchange` (all partitions in one process: tests, one GPU).

Overlap: the owned cells whose state a neighbour needs (*interface* cells, ``owned == 2``) are stored first in
their type group and marched first (``EulerCore.substep_interface``); their message is packed on the compute
stream, an event hands it to a communication stream where NCCL sends it while the compute stream marches the
interior cells (``substep_rest``); a second event makes the unpack -- and with it the next substep -- wait for
the receive.  No host synchronisation is left inside ``march``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import meshgen
from .mesh import StaticMesh, make_mesh


# ----------------------------------------------------------------------------- partitioning ---

def rcb_partition(centroids, nparts):
    """Recursive coordinate bisection of the cell centroids into ``nparts`` parts of (nearly) equal
    size; deterministic (stable sorts), any ``nparts >= 1``.  Returns ``part[ncell]`` (int32)."""
    n = len(centroids)
    part = np.zeros(n, dtype=np.int32)

    def split(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        c = centroids[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = np.argsort(c[:, axis], kind="stable")
        nl = np_ // 2
        cut = (len(idx) * nl) // np_
        split(idx[order[:cut]], p0, nl)
        split(idx[order[cut:]], p0 + nl, np_ - nl)

    split(np.arange(n), 0, int(nparts))
    return part


@dataclass
class Partition:
    rank: int
    ndim: int
    ndcrd: np.ndarray          # local nodes
    cltpn: np.ndarray          # local cells, ascending global id
    clnds: np.ndarray
    owned: np.ndarray          # int8 mask over the local cells
    gid: np.ndarray            # global id of every local cell
    send: dict = field(default_factory=dict)  # neighbour rank -> local cell ids (ascending global id)
    recv: dict = field(default_factory=dict)  # neighbour rank -> local cell ids of the halo copies
    mesh: StaticMesh = None

    @property
    def neighbours(self):
        return sorted(set(self.send) | set(self.recv))

    def build_mesh(self):
        if self.mesh is None:
            self.mesh = make_mesh(self.ndim, self.ndcrd, self.cltpn, self.clnds)
        return self.mesh

    def send_list(self):
        return np.concatenate([self.send[q] for q in self.neighbours]) if self.send else np.zeros(0, np.int32)

    def recv_list(self):
        return np.concatenate([self.recv[q] for q in self.neighbours]) if self.recv else np.zeros(0, np.int32)

    def physical_boundary_groups(self, lo, hi):
        """(inlet, outlet, wall) boundary faces of OWNED cells; faces whose interior cell is a halo copy are
        artefacts of the cut and get no boundary condition."""
        m = self.build_mesh()
        inlet, outlet, wall = meshgen.boundary_groups(self.ndim, m.fccnd.body, m.fccls.body, m.nface, lo, hi)
        own = self.owned.astype(bool)
        keep = lambda f: f[own[m.fccls.body[f, 0]]]  # noqa: E731
        return keep(inlet), keep(outlet), keep(wall)


def make_partitions(ndim, ndcrd, cltpn, clnds, part, ranks=None):
    """Cut a global mesh (nodes + cells) into partitions with one halo layer.  ``part[ncell]`` assigns every
    cell to a rank.  Returns ``{rank: Partition}`` for ``ranks`` (default: all)."""
    ndcrd = np.asarray(ndcrd)
    nparts = int(part.max()) + 1
    g = StaticMesh(ndim, len(ndcrd), 0, len(cltpn))
    g.ndcrd.ndarray[...] = ndcrd
    g.cltpn.ndarray[...] = cltpn
    g.clnds.ndarray[...] = clnds
    g.build_interior(False, False)
    fccls = g.fccls.ndarray[:, :2]
    inner = fccls[fccls[:, 1] >= 0]
    a, b = inner[:, 0], inner[:, 1]
    cross = part[a] != part[b]
    a, b = a[cross], b[cross]
    out = {}
    ia, ib = inner[:, 0], inner[:, 1]
    for r in (range(nparts) if ranks is None else ranks):
        # cells of other ranks that share a face with a cell of rank r, and the other way round
        halo_of_r = np.unique(np.concatenate([b[part[a] == r], a[part[b] == r]]))
        own = np.nonzero(part == r)[0]
        # second (passive) layer: face neighbours of the first layer
        in1 = np.zeros(len(cltpn), dtype=bool)
        in1[halo_of_r] = True
        layer2 = np.unique(np.concatenate([ib[in1[ia]], ia[in1[ib]]]))
        local = np.union1d(np.union1d(own, halo_of_r), layer2)  # ascending global id
        lid = np.full(len(cltpn), -1, dtype=np.int64)
        lid[local] = np.arange(len(local))
        lcl = clnds[local].copy()
        valid = lcl[:, 1:] >= 0
        nodes = np.unique(lcl[:, 1:][valid])
        nmap = np.full(len(ndcrd), -1, dtype=np.int64)
        nmap[nodes] = np.arange(len(nodes))
        lcl[:, 1:][valid] = nmap[lcl[:, 1:][valid]]
        p = Partition(rank=r, ndim=ndim, ndcrd=np.ascontiguousarray(ndcrd[nodes]), cltpn=cltpn[local].astype(np.int32),
                      clnds=lcl.astype(np.int32), owned=(part[local] == r).astype(np.int8), gid=local)
        for q in np.unique(part[halo_of_r]):
            p.recv[int(q)] = lid[halo_of_r[part[halo_of_r] == q]].astype(np.int32)
            # my cells that rank q holds as halo copies: owned by r, adjacent to a cell of q
            mine = np.unique(np.concatenate([a[(part[a] == r) & (part[b] == q)], b[(part[b] == r) & (part[a] == q)]]))
            p.send[int(q)] = lid[mine].astype(np.int32)
        out[r] = p
    return out


def slab_partition_3d(rank, nparts, nx, ny_per_rank, nz, h, x1, z1, kinds="hpt", jitter=0.1):
    """Partition ``rank`` of the weak-scaling stepped channel: the domain is
    extended in y with the number of ranks (``ny = nparts * ny_per_rank`` boxes) and rank r owns the box
    layers ``r*ny_per_rank <= j < (r+1)*ny_per_rank``.  Only the own layers plus two halo layers per side
    (the exchanged one and the passive one) are generated, in global lattice coordinates, so no rank ever
    holds the global mesh."""
    ny = nparts * ny_per_rank
    j0, j1 = rank * ny_per_rank, (rank + 1) * ny_per_rank
    g0, g1 = max(0, j0 - 2), min(ny, j1 + 2)
    ndcrd, cltpn, clnds, bj = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=ny * h, lz=nz * h, x1=x1, z1=z1, kinds=kinds,
                                               jitter=jitter, j_range=(g0, g1), return_box_j=True)
    owned = ((bj >= j0) & (bj < j1)).astype(np.int8)
    ids = np.arange(len(cltpn), dtype=np.int32)
    p = Partition(rank=rank, ndim=3, ndcrd=ndcrd, cltpn=cltpn, clnds=clnds, owned=owned, gid=ids.astype(np.int64))
    # Every cell of a box touches the box's own y-faces?  No: only some cells of a box have a face on the
    # slab interface, but every face neighbour across the interface lies in the adjacent box layer, so
    # exchanging whole layers is a (slightly generous) superset and keeps both sides' lists trivially
    # consistent: ascending local order == ascending global order on both ranks.
    if rank > 0:
        p.recv[rank - 1] = ids[bj == j0 - 1]
        p.send[rank - 1] = ids[bj == j0]
    if rank < nparts - 1:
        p.recv[rank + 1] = ids[bj == j1]
        p.send[rank + 1] = ids[bj == j1 - 1]
    return p


# Relative cost of one cell-substep by cell type (measured kernel times per cell, profiles/r01f_md3d_launches_summary:
# hexahedron 0.92 us, prism 0.78 us, tetrahedron 0.58 us per 148-SM wave): the weights the strong-scaling cut balances.
CELL_COST = {"hex": 1.0, "prism": 0.85, "tet": 0.63}


def lattice_weight(rng, x1, z1, cost=None):
    """Cost-weighted cell count of the boxes ``rng = ((i0,i1),(j0,j1),(k0,k1))`` of the stepped channel of
    ``meshgen.mixed_3d`` (closed form: the cell kind of a box depends on ``i < x1`` and ``k < z1`` only)."""
    c = cost or {"hex": 1.0, "prism": 1.0, "tet": 1.0}
    (i0, i1), (j0, j1), (k0, k1) = rng
    ilo, ihi = max(0, min(i1, x1) - i0), max(0, i1 - max(i0, x1))      # boxes with i < x1 / i >= x1
    klo, khi = max(0, min(k1, z1) - k0), max(0, k1 - max(k0, z1))      # boxes with k < z1 / k >= z1
    dj = max(0, j1 - j0)
    return dj * (c["hex"] * ilo * klo + 2.0 * c["prism"] * ihi * klo + 6.0 * c["tet"] * ihi * khi)


def lattice_rcb(nx, ny, nz, x1, z1, nparts, cost=None, spacing=(1.0, 1.0, 1.0)):
    """Recursive coordinate bisection of the stepped channel's box lattice into ``nparts`` box ranges of (nearly)
    equal cost-weighted cell count: the RCB of SURVEY.md 8d config 4 at box granularity, evaluated in closed form so
    that every rank can compute all cuts without any rank holding the global mesh.  Cuts are perpendicular to the
    physically longest axis of the range being cut.  Returns ``[((i0,i1),(j0,j1),(k0,k1)), ...]`` (index = rank)."""
    out = [None] * int(nparts)

    def bbox(rng):
        """Shrink a range to the non-solid boxes it contains (the region i < x1, k >= z1 holds no cells)."""
        (i0, i1), (j0, j1), (k0, k1) = rng
        if k0 >= z1:
            i0 = max(i0, x1)
        if i1 <= x1:
            k1 = min(k1, z1)
        return ((i0, i1), (j0, j1), (k0, k1))

    def split(rng, p0, np_):
        rng = bbox(rng)
        if np_ == 1:
            out[p0] = rng
            return
        nl = np_ // 2
        total = lattice_weight(rng, x1, z1, cost)
        target = total * nl / np_
        ext = [(b - a) * sp for (a, b), sp in zip(rng, spacing)]
        for axis in sorted(range(3), key=lambda a: -ext[a]):
            a0, a1 = rng[axis]
            if a1 - a0 < 2:
                continue
            best, bestd = None, None
            for c in range(a0 + 1, a1):
                lo = list(rng)
                lo[axis] = (a0, c)
                d = abs(lattice_weight(tuple(lo), x1, z1, cost) - target)
                if bestd is None or d < bestd:
                    best, bestd = c, d
            lo, hi = list(rng), list(rng)
            lo[axis], hi[axis] = (a0, best), (best, a1)
            if lattice_weight(tuple(lo), x1, z1, cost) > 0 and lattice_weight(tuple(hi), x1, z1, cost) > 0:
                split(tuple(lo), p0, nl)
                split(tuple(hi), p0 + nl, np_ - nl)
                return
        raise ValueError(f"lattice_rcb: cannot cut {rng} into {np_} non-empty parts")

    split(((0, nx), (0, ny), (0, nz)), 0, int(nparts))
    return out


def lattice_partition_3d(rank, ranges, nx, ny, nz, h, x1, z1, kinds="hpt", jitter=0.1):
    """Partition ``rank`` of the stepped channel cut into the box ranges ``ranges`` (``lattice_rcb``).  The rank
    generates its own boxes plus two box layers around them (the exchanged halo layer and the passive one) straight
    from the lattice formula; send / receive lists are the cells that really share a face across a cut (found in the
    LOCAL face table), ordered by ``meshgen.cell_key`` -- the same ascending global order on both sides of a rank
    pair.  No rank ever holds the global mesh."""
    (i0, i1), (j0, j1), (k0, k1) = ranges[rank]
    grow = ((i0 - 2, i1 + 2), (j0 - 2, j1 + 2), (k0 - 2, k1 + 2))
    ndcrd, cltpn, clnds, box = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=ny * h, lz=nz * h, x1=x1, z1=z1, kinds=kinds,
                                                jitter=jitter, box_range=grow, return_box=True)
    bi, bj, bk, _ = box
    key = meshgen.cell_key(box, ny, nz)

    def inside(r):
        (a0, a1), (b0, b1), (c0, c1) = r
        return (bi >= a0) & (bi < a1) & (bj >= b0) & (bj < b1) & (bk >= c0) & (bk < c1)

    owner = np.full(len(cltpn), -1, dtype=np.int32)
    for q, r in enumerate(ranges):
        owner[inside(r)] = q
    owned = (owner == rank).astype(np.int8)
    p = Partition(rank=rank, ndim=3, ndcrd=ndcrd, cltpn=cltpn, clnds=clnds, owned=owned, gid=key)
    fccls = p.build_mesh().fccls.body
    a, b = fccls[:, 0], fccls[:, 1]
    inner = b >= 0
    a, b = a[inner], b[inner]
    cut = (owner[a] == rank) != (owner[b] == rank)
    mine = np.where(owner[a[cut]] == rank, a[cut], b[cut])
    theirs = np.where(owner[a[cut]] == rank, b[cut], a[cut])
    for q in np.unique(owner[theirs]):
        m = owner[theirs] == q
        p.recv[int(q)] = np.unique(theirs[m]).astype(np.int32)      # ascending local id == ascending key
        p.send[int(q)] = np.unique(mine[m]).astype(np.int32)
    return p


# ----------------------------------------------------------------------------------- transport ---

class LocalExchange:
    """All partitions live in this process (tests; several partitions on one GPU): a 'message' is the
    sender's packed device buffer, read directly by the receiver's unpack kernel."""

    def __init__(self, solvers):
        self.solvers = solvers  # {rank: DistEulerCore}

    def exchange(self, what):
        packed = {}
        for r, s in self.solvers.items():
            ptr, n = s.core.halo_pack(what)
            packed[r] = (ptr, s.send_offsets)
        from . import _lib
        _lib.sync()
        for r, s in self.solvers.items():
            w = s.core.halo_width(what)
            # the receiver's unpack wants ONE buffer ordered like its recv list: gather the segments
            segs = []
            for q in s.part.neighbours:
                ptr, offs = packed[q]
                o0, o1 = offs[r]
                segs.append((ptr + o0 * w * 8, (o1 - o0) * w))
            s.unpack_segments(segs, what)


class LocalSplitExchange(LocalExchange):
    """LocalExchange driving the two-stage substep (interface cells, pack, interior cells, unpack) of every
    partition in turn -- the kernel ranges and the ordering of the overlapped march without the streams."""

    def march(self, steps):
        for _ in range(2 * int(steps)):
            for s in self.solvers.values():
                s.core.substep_interface()
            packed = {}
            for r, s in self.solvers.items():
                ptr, n = s.core.halo_pack(0)
                packed[r] = (ptr, s.send_offsets)
            for s in self.solvers.values():
                s.core.substep_rest()
            from . import _lib
            _lib.sync()
            for r, s in self.solvers.items():
                w = s.core.halo_width(0)
                segs = []
                for q in s.part.neighbours:
                    ptr, offs = packed[q]
                    o0, o1 = offs[r]
                    segs.append((ptr + o0 * w * 8, (o1 - o0) * w))
                s.unpack_segments(segs, 0)


class TorchExchange:
    """``torch.distributed`` transport (NCCL between GPUs, gloo on the CPU): grouped isend/irecv of the
    packed buffers with every neighbour rank."""

    def __init__(self, solver, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.s, self.device = torch, dist, solver, torch.device(device)
        self.recvbuf = {}
        self.cuda = self.device.type == "cuda"

    def _view(self, ptr, n):
        """A tensor over ``n`` doubles at ``ptr`` without a copy: device memory of the library on a GPU, host memory
        of the solver object when the process group is gloo (the CPU tests of the exchange logic)."""
        if not self.cuda:
            import ctypes as C
            return self.torch.from_numpy(np.ctypeslib.as_array((C.c_double * int(n)).from_address(int(ptr))))

        class _Dev:
            pass
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
        return self.torch.as_tensor(d, device=self.device)

    def _recvbuf(self, what, n):
        """Receive buffer of message kind ``what``; allocated once, and the allocation (made on torch's current
        stream) is complete before the library's stream or the communication stream may touch it."""
        if what not in self.recvbuf:
            self.recvbuf[what] = self.torch.zeros(max(1, n), dtype=self.torch.float64, device=self.device)
            if self.cuda:
                self.torch.cuda.current_stream(self.device).synchronize()
        return self.recvbuf[what]

    def _sync_library(self):
        if self.cuda:
            from . import _lib
            _lib.sync()  # the library packs on its own stream

    def exchange(self, what):
        ptr, n = self.s.core.halo_pack(what)
        return self._exchange_packed(what, ptr, n)

    def _exchange_packed(self, what, ptr, n):
        torch, dist, s = self.torch, self.dist, self.s
        w = s.core.halo_width(what)
        self._sync_library()
        ops = []
        rb = self._recvbuf(what, s.nrecv * w)
        sendall = self._view(ptr, max(1, n)) if n else None
        ro = 0
        for q in s.part.neighbours:
            o0, o1 = s.send_offsets[q]
            if o1 > o0:
                ops.append(dist.P2POp(dist.isend, sendall[o0 * w:o1 * w], q))
            nr = len(s.part.recv.get(q, ()))
            if nr:
                ops.append(dist.P2POp(dist.irecv, rb[ro * w:(ro + nr) * w], q))
            ro += nr
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if self.cuda:
                torch.cuda.synchronize(self.device)
        if s.nrecv:
            s.core.halo_unpack(rb.data_ptr(), what)

    def exchange_split(self, what, rest):
        """Overlapped exchange (CUDA only): pack on the library stream -> event -> NCCL on a communication stream
        while ``rest()`` (the interior march) is enqueued on the library stream -> event -> unpack."""
        torch, dist, s = self.torch, self.dist, self.s
        if not self.cuda:
            # no streams on the CPU: same order of operations, run back to back
            ptr, n = s.core.halo_pack(what)
            rest()
            return self._exchange_packed(what, ptr, n)
        if getattr(self, "_lib_stream", None) is None:
            from . import _lib
            self._lib_stream = torch.cuda.ExternalStream(int(_lib.load().scg_stream_handle()), device=self.device)
            self._comm = torch.cuda.Stream(device=self.device)
            self._ev_packed = torch.cuda.Event()
            self._ev_done = torch.cuda.Event()
        w = s.core.halo_width(what)
        ptr, n = s.core.halo_pack(what)
        self._ev_packed.record(self._lib_stream)
        rest()  # interior cells + boundary kernels, enqueued behind the pack on the library stream
        rb = self._recvbuf(what, s.nrecv * w)
        sendall = self._view(ptr, max(1, n)) if n else None
        ops, ro = [], 0
        for q in s.part.neighbours:
            o0, o1 = s.send_offsets[q]
            if o1 > o0:
                ops.append(dist.P2POp(dist.isend, sendall[o0 * w:o1 * w], q))
            nr = len(s.part.recv.get(q, ()))
            if nr:
                ops.append(dist.P2POp(dist.irecv, rb[ro * w:(ro + nr) * w], q))
            ro += nr
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(self._ev_packed)
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()  # stream-ordered for NCCL: the communication stream waits, not the host
            self._ev_done.record(self._comm)
        self._lib_stream.wait_event(self._ev_done)
        if s.nrecv:
            s.core.halo_unpack(rb.data_ptr(), what)


def _libnccl_path():
    """Where this environment keeps libnccl (the wheel torch depends on), for ``scg_comm_*``'s dlopen; None lets the
    library try the default soname (a copy already mapped into the process is reused either way)."""
    import glob
    import os
    import sys
    for root in sys.path:
        hits = glob.glob(os.path.join(root, "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return sorted(hits)[0]
    return None


def comm_init(rank, world, dist=None, id_file=None, timeout=120.0):
    """Create the LIBRARY's communicator (``scg_comm_init``: NCCL inside libscg_cuda, csrc/scg_comm.cu).  Only the
    128-byte unique id travels out of band: through ``torch.distributed`` if a process group is given (any backend),
    else through ``id_file`` (rank 0 writes it, the others poll) -- the product path needs no torch."""
    import ctypes as C
    import os
    import time
    from . import _lib
    lib = _lib.load()
    path = _libnccl_path()
    cpath = path.encode() if path else None
    buf = (C.c_char * 128)()
    if rank == 0:
        _lib.check(lib.scg_comm_unique_id(buf, 128, cpath))
    if dist is not None:
        box = [bytes(buf.raw) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        raw = box[0]
    else:
        if id_file is None:
            raise ValueError("comm_init: give a torch.distributed module or an id_file")
        if rank == 0:
            with open(id_file + ".tmp", "wb") as f:
                f.write(bytes(buf.raw))
            os.replace(id_file + ".tmp", id_file)
            raw = bytes(buf.raw)
        else:
            t0 = time.monotonic()
            while not os.path.exists(id_file):
                if time.monotonic() - t0 > timeout:
                    raise TimeoutError(f"comm_init: {id_file} did not appear")
                time.sleep(0.05)
            with open(id_file, "rb") as f:
                raw = f.read()
    idbuf = (C.c_char * 128).from_buffer_copy(raw)
    _lib.check(lib.scg_comm_init(int(rank), int(world), idbuf, cpath))
    ver = C.c_int()
    lib.scg_comm_info(None, None, C.byref(ver))
    return int(ver.value)


class NcclExchange:
    """Transport through the library's own communicator: every exchange is ONE C call (pack kernel, one ncclGroup of
    ncclSend / ncclRecv with all neighbour ranks on the library's communication stream, unpack kernel; stream-ordered,
    no host synchronisation), and ``march`` is ``scg_md_march_dist`` -- the whole overlapped march in C++.  No torch
    on this path.  Bit-identical to :class:`TorchExchange` and to the undecomposed march."""

    def __init__(self, solver):
        from . import _lib
        self.s = solver
        self._lib = _lib.load()
        nb = solver.part.neighbours
        ranks = np.array(nb, dtype=np.int32)
        sb = np.array([solver.send_offsets[q][0] for q in nb], dtype=np.int64)
        se = np.array([solver.send_offsets[q][1] for q in nb], dtype=np.int64)
        rl = np.array([len(solver.part.recv.get(q, ())) for q in nb], dtype=np.int64)
        re = np.cumsum(rl)
        rb = re - rl
        p = lambda a: _lib.ptr(a) if len(a) else None  # noqa: E731
        _lib.check(self._lib.scg_md_halo_peers(solver.core._h, len(nb), p(ranks), p(sb), p(se), p(rb), p(re)))

    def exchange(self, what):
        from . import _lib
        _lib.check(self._lib.scg_md_halo_exchange(self.s.core._h, int(what)))

    def march(self, steps):
        from . import _lib
        _lib.check(self._lib.scg_md_march_dist(self.s.core._h, int(steps)))
        self.s.core._pull()     # the device holds newer data than any host copy


# ------------------------------------------------------------------------------------- solver ---

class HaloSolver:
    """One partition's solver plus its halo bookkeeping.  ``core`` is anything with the halo interface of
    :class:`solvcon_b200.multidim.EulerCore` (``halo_set, halo_width, halo_pack, halo_unpack, substep_local,
    substep_interface, substep_rest``): the device core in the product (:class:`DistEulerCore`), a host-memory double
    over the CPU checker in the gloo tests of the exchange logic.
    ``march(steps)`` = per substep: local substep on the owned cells, then the halo exchange."""

    def __init__(self, part: Partition, core):
        self.part = part
        self.core = core
        snd, rcv = part.send_list(), part.recv_list()
        self.nsend, self.nrecv = len(snd), len(rcv)
        self.core.halo_set(snd, rcv)
        self.send_offsets, o = {}, 0
        for q in part.neighbours:
            k = len(part.send.get(q, ()))
            self.send_offsets[q] = (o, o + k)
            o += k
        self.transport = None
        self._scratch = None

    def march(self, steps):
        if hasattr(self.transport, "march"):       # the library marches the partition itself (NcclExchange)
            return self.transport.march(steps)
        split = getattr(self.transport, "exchange_split", None)
        for _ in range(2 * int(steps)):
            if split is not None:
                self.core.substep_interface()
                split(0, self.core.substep_rest)
            else:
                self.core.substep_local()
                self.transport.exchange(0)


class DistEulerCore(HaloSolver):
    """``EulerCore`` of one partition: the device solver of the local mesh plus the halo bookkeeping."""

    def __init__(self, part: Partition, time_increment, order="morton"):
        from .multidim import EulerCore
        mesh = part.build_mesh()
        own = part.owned.astype(np.int8).copy()
        snd0 = part.send_list()
        if len(snd0):
            own[snd0] = 2  # interface cells: marched first so that their message travels during the interior march
        core = EulerCore(mesh, time_increment, order=order, owned=own)
        core.resident(True)
        super().__init__(part, core)

    def unpack_segments(self, segs, what):
        """LocalExchange helper: unpack from several source buffers (one per neighbour, in neighbour order)."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        if len(segs) == 1:
            self.core.halo_unpack(segs[0][0], what)
            return
        # more than one neighbour: stage the segments contiguously (device-to-device copies)
        total = sum(n for _, n in segs)
        if self._scratch is None or self._scratch[1] < total:
            p = C.c_void_p()
            _lib.check(lib.scg_dev_alloc(C.byref(p), total * 8))
            self._scratch = (p.value, total)
        o = 0
        for ptr, n in segs:
            _lib.check(lib.scg_dev_copy(C.c_void_p(self._scratch[0] + o * 8), C.c_void_p(ptr), n * 8))
            o += n
        self.core.halo_unpack(self._scratch[0], what)

    def upload_state(self):
        """Push host-side edits of the field (initial condition) to the device."""
        self.core._push(force=True)

    def download_state(self):
        self.core.sync_to_host()
