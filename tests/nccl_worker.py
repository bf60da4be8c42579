"""Worker of tests/test_gpu_nccl.py: one rank of a 2-GPU job.  A small stepped channel is cut into one part per rank
and marched over the real transports -- the library's own NCCL communicator (scg_md_march_dist) and the
torch.distributed one -- and compared on rank 0, bit for bit, with the undecomposed device march."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ring_parity(rank, world, dist, ncelm=4001, steps=25):
    """1D scalar ring over the library's communicator (scg_st_march_ring) against the undecomposed device solver."""
    import numpy as np
    from solvcon_b200 import spacetime as st
    from solvcon_b200.decomp1d import NcclRing, slab_ranges
    ok = True
    for eq in (0, 1):
        grid = st.Grid(0.0, 8.0, ncelm)
        x = grid.xcoord.ndarray
        dt = 0.4 * 8.0 / ncelm
        so0, so1 = np.zeros(len(x)), np.zeros(len(x))
        so0[2::2] = 0.5 * np.sin(x[2::2] * np.pi / 4.0)
        so1[2::2] = 0.5 * np.cos(x[2::2] * np.pi / 4.0) * np.pi / 4.0
        c0, c1 = slab_ranges(ncelm, world)[rank]
        ring = NcclRing(eq, x, c0, c1, dt, rank, world)
        ring.slab.load_global(so0, so1)
        ring.march(2, steps)
        a, b, c = ring.slab.fetch()
        mine = {"c0": c0, "c1": c1, "so0": a.copy(), "so1": b.copy(), "cfl": c.copy()}
        got = [None] * world if rank == 0 else None
        dist.gather_object(mine, got, dst=0)
        if rank == 0:
            ref = (st.InviscidBurgersSolver if eq else st.LinearScalarSolver)(grid, dt)
            ref.so0.ndarray[:, 0] = so0
            ref.so1.ndarray[:, 0] = so1
            ref.march_alpha2(steps)
            for g in got:
                sl = slice(2 * g["c0"] + 1, 2 * g["c1"] + 4)     # every slot but the two outermost ghost coordinates
                loc = slice(1, -1)
                same = (np.array_equal(g["so0"][loc], ref.so0.ndarray[sl, 0]) and
                        np.array_equal(g["so1"][loc], ref.so1.ndarray[sl, 0]) and
                        np.array_equal(g["cfl"][loc], ref.cfl.ndarray[sl]))
                ok = ok and same
            print("ring eq", eq, {"bit_equal": ok}, flush=True)
    return ok


def main():
    import torch
    import torch.distributed as dist
    from solvcon_b200 import _lib, bench_md
    from solvcon_b200.decomp import comm_init
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.init(local)
    comm_init(rank, world, dist=dist)
    ok = True
    for transport in ("lib", "torch"):
        res, good = bench_md.nccl_parity_3d(rank, world, local, dist, "rcm", ncell_target=60_000, steps=4, transport=transport)
        ok = ok and good
        if rank == 0:
            print(transport, res, flush=True)
    ok = ring_parity(rank, world, dist) and ok
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
