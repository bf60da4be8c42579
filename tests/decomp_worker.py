"""Worker of tests/test_decomp_gloo.py: one rank of a world_size-N gloo job.  Each rank marches ITS partition through
the PRODUCT's ``solvcon_b200.decomp.HaloSolver`` + ``TorchExchange`` (gloo): the very classes, message layout
(neighbour order, send offsets, one grouped isend/irecv per exchange) and march loop (``exchange`` and the two-stage
``exchange_split`` ordering) that run over NCCL on the GPUs.  Only the compute engine differs: a host-memory double of
the device core's halo interface over the CPU checker (``HostHaloCore`` below; test infrastructure).  Every rank
compares its owned cells with the undecomposed run, bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ic(so0_body, clcnd, ndim, ctr, span, gamma=1.4):
    r2 = (((clcnd - ctr) / (0.25 * span)) ** 2).sum(axis=1)
    rho = 1.4 * (1.0 + 0.3 * np.exp(-r2))
    p = 1.0 * (1.0 + 0.3 * np.exp(-r2))
    v = np.zeros((len(clcnd), ndim))
    v[:, 0] = 2.0
    v[:, 1] = 0.1 * np.sin(3.0 * clcnd[:, 0])
    so0_body[:, 0] = rho
    for d in range(ndim):
        so0_body[:, 1 + d] = rho * v[:, d]
    so0_body[:, ndim + 1] = p / (gamma - 1.0) + 0.5 * rho * (v * v).sum(axis=1)


def setup(b, svr, mesh, groups, ndim, ctr, span):
    inlet, outlet, wall = groups
    vel = [2.0, 0.0, 0.0][:ndim]
    svr.init_solution(1.4, 1.4, vel, 1.0)
    ic(svr.body("so0n"), mesh.clcnd.body, ndim, ctr, span)
    svr.add_bc(b.BC_INLET, inlet, [1.4] + vel + [1.0, 1.4])
    svr.add_bc(b.BC_NONREFL, outlet)
    svr.add_bc(b.BC_SLIPWALL, wall)
    svr.phase(b.MD_BC_SOLN)
    svr.phase(b.MD_BC_DSOLN)


class HostHaloCore:
    """The halo interface of ``solvcon_b200.multidim.EulerCore`` (halo_set / halo_width / halo_pack / halo_unpack /
    substep_local / substep_interface / substep_rest) over the CPU checker, buffers in host memory."""

    def __init__(self, b, svr, ndim, owned):
        self.b, self.svr, self.ndim, self.neq = b, svr, ndim, ndim + 2
        self.owned = np.asarray(owned)
        self.calls = []

    def halo_set(self, send_cells, recv_cells):
        self.snd, self.rcv = np.asarray(send_cells, dtype=np.int64), np.asarray(recv_cells, dtype=np.int64)
        self.buf = {w: np.zeros(max(1, len(self.snd) * self.halo_width(w))) for w in (0, 1)}

    def halo_width(self, what=0):
        return self.neq * (1 + self.ndim) if what == 0 else self.ndim + 1

    def _fields(self, what):
        if what == 0:
            return [("so0n", slice(None)), ("so1n", slice(None))]
        return [("cecnd", slice(0, self.ndim)), ("gamma", slice(None))]

    def halo_pack(self, what=0):
        rows = np.concatenate([self.svr.body(n)[self.snd].reshape(len(self.snd), -1)[:, sl] for n, sl in self._fields(what)], axis=1)
        self.buf[what][:rows.size] = rows.ravel()
        self.calls.append(("pack", what))
        return self.buf[what].ctypes.data, rows.size

    def halo_unpack(self, ptr, what=0):
        import ctypes as C
        w = self.halo_width(what)
        a = np.ctypeslib.as_array((C.c_double * (len(self.rcv) * w)).from_address(int(ptr))).reshape(len(self.rcv), w)
        col = 0
        for n, sl in self._fields(what):
            tgt = self.svr.body(n)
            k = tgt[self.rcv].reshape(len(self.rcv), -1)[:, sl].shape[1]
            tgt.reshape(len(tgt), -1)[self.rcv, sl] = a[:, col:col + k]
            col += k
        self.calls.append(("unpack", what))

    def substep_local(self):
        self.svr.phase(self.b.MD_SUBSTEP)
        self.calls.append(("substep",))

    # the checker has no cell ranges: the whole substep runs in the first stage, the second is empty.  What the test
    # pins is the ORDER the product's march drives (interface -> pack -> rest -> exchange -> unpack).
    def substep_interface(self):
        self.svr.phase(self.b.MD_SUBSTEP)
        self.calls.append(("iface",))

    def substep_rest(self):
        self.calls.append(("rest",))


def main():
    import torch
    import torch.distributed as dist
    from cases import md_mesh_case
    from oracle import bridge as b
    from solvcon_b200 import meshgen
    from solvcon_b200.decomp import HaloSolver, TorchExchange, make_partitions, rcb_partition
    from solvcon_b200.mesh import make_mesh

    name, steps = sys.argv[1], int(sys.argv[2])
    split = len(sys.argv) > 3 and sys.argv[3] == "split"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    h = (np.prod(hi - lo) / len(cltpn)) ** (1.0 / ndim)
    dt = 0.05 * h
    gm = make_mesh(ndim, ndcrd, cltpn, clnds)
    ctr, span = gm.clcnd.body.mean(axis=0), hi - lo
    # the undecomposed run
    ref = b.PortEuler(gm.tables(), ndim, dt)
    ref.phase(b.MD_PREPARE_CE)
    setup(b, ref, gm, meshgen.boundary_groups(ndim, gm.fccnd.body, gm.fccls.body, gm.nface, lo, hi), ndim, ctr, span)
    ref.march(steps)
    # my partition, marched by the product's HaloSolver over the product's TorchExchange
    part = rcb_partition(gm.clcnd.body, world)
    p = make_partitions(ndim, ndcrd, cltpn, clnds, part, ranks=[rank])[rank]
    lm = p.build_mesh()
    svr = b.PortEuler(lm.tables(), ndim, dt)
    svr.phase(b.MD_PREPARE_CE)
    setup(b, svr, lm, p.physical_boundary_groups(lo, hi), ndim, ctr, span)
    core = HostHaloCore(b, svr, ndim, p.owned)
    hs = HaloSolver(p, core)
    hs.transport = TorchExchange(hs, "cpu")
    hs.transport.exchange(1)   # solution point + gamma of the halo copies, once
    hs.transport.exchange(0)   # state
    if not split:
        hs.transport.exchange_split = None   # plain path: substep_local + exchange
    core.calls.clear()
    hs.march(steps)
    want = ([("iface",), ("pack", 0), ("rest",), ("unpack", 0)] if split else [("substep",), ("pack", 0), ("unpack", 0)])
    if not p.recv:
        want = [c for c in want if c[0] != "unpack"]
    ok = core.calls == want * (2 * steps)
    if not ok:
        print(f"rank {rank}: unexpected call order {core.calls[:8]}", flush=True)
    own = p.owned.astype(bool)
    for n in ("so0n", "so1n", "cflo", "cflc"):
        if not np.array_equal(svr.body(n)[own], ref.body(n)[p.gid[own]]):
            ok = False
            print(f"rank {rank}: {n} differs, max abs {np.abs(svr.body(n)[own] - ref.body(n)[p.gid[own]]).max()}", flush=True)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
