"""Worker of tests/test_decomp_gloo.py: one rank of a world_size-N gloo job.  Each rank marches ITS partition
with the CPU checker (oracle port) as the compute engine and exchanges halos through torch.distributed, then
compares its owned cells with the undecomposed run.  Exercises the host side of solvcon_b200.decomp:
partitioning, the two halo layers, list ordering on both sides of every rank pair, and the exchange."""
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


def main():
    import torch
    import torch.distributed as dist
    from cases import md_mesh_case
    from oracle import bridge as b
    from solvcon_b200 import meshgen
    from solvcon_b200.decomp import make_partitions, rcb_partition
    from solvcon_b200.mesh import make_mesh

    name, steps = sys.argv[1], int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    neq = ndim + 2
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
    # my partition
    part = rcb_partition(gm.clcnd.body, world)
    p = make_partitions(ndim, ndcrd, cltpn, clnds, part, ranks=[rank])[rank]
    lm = p.build_mesh()
    svr = b.PortEuler(lm.tables(), ndim, dt)
    svr.phase(b.MD_PREPARE_CE)
    setup(b, svr, lm, p.physical_boundary_groups(lo, hi), ndim, ctr, span)

    def exchange(fields):
        """fields: list of (array name, column slice or None); one message per neighbour."""
        reqs, recvs = [], []
        for q in p.neighbours:
            snd = np.concatenate([svr.body(n)[p.send[q]].reshape(len(p.send[q]), -1)[:, sl] for n, sl in fields], axis=1)
            t = torch.from_numpy(np.ascontiguousarray(snd))
            reqs.append(dist.isend(t, q))
            r = torch.empty((len(p.recv[q]), snd.shape[1]), dtype=torch.float64)
            reqs.append(dist.irecv(r, q))
            recvs.append((q, r, t))
        for r in reqs:
            r.wait()
        for q, r, _ in recvs:
            a, col = r.numpy(), 0
            for n, sl in fields:
                tgt = svr.body(n)
                w = tgt[p.recv[q]].reshape(len(p.recv[q]), -1)[:, sl].shape[1]
                tgt.reshape(len(tgt), -1)[p.recv[q], sl] = a[:, col:col + w]
                col += w

    exchange([("cecnd", slice(0, ndim)), ("gamma", slice(None))])   # what = 1: solution point + gamma
    exchange([("so0n", slice(None)), ("so1n", slice(None))])          # what = 0: state
    for _ in range(2 * steps):
        svr.phase(b.MD_SUBSTEP)
        exchange([("so0n", slice(None)), ("so1n", slice(None))])
    own = p.owned.astype(bool)
    ok = True
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
