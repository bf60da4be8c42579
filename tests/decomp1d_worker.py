"""Worker of tests/test_decomp_gloo.py (1D): one rank of a world_size-N gloo job marches ITS slab of the periodic 1D
scalar grid with the CPU checker (oracle port) as the compute engine, exchanges the odd-plane end points with its
ring neighbours through torch.distributed, and compares every slot with the undecomposed march.  Exercises the host
side of solvcon_b200.decomp1d (slab ranges, global coordinate slices, message order on both sides of a rank pair)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PH_SO0, PH_CFL, PH_SO1 = 0, 1, 2


def main():
    import torch
    import torch.distributed as dist
    from cases import scalar_ic
    from oracle import bridge as b
    from solvcon_b200.decomp1d import slab_ranges
    from solvcon_b200.spacetime import Grid

    eq, n, steps, alpha = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 2
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    xmax = 8 * np.pi
    dt = 0.5 * xmax / n
    ref = b.PortScalar1D(eq, 0.0, xmax, n, dt)
    xg = Grid(0.0, xmax, n).xcoord.ndarray
    assert np.array_equal(xg, ref.x)
    scalar_ic(ref.x, ref.so0, ref.so1, ref.cfl)
    c0, c1 = slab_ranges(n, world)[rank]
    m = c1 - c0
    sl = slice(2 * c0, 2 * c1 + 5)
    svr = b.PortScalar1D(eq, dt=dt, xloc=xg[sl][2:2 * m + 3:2])
    svr.x[...] = xg[sl]
    svr.so0[...], svr.so1[...], svr.cfl[...] = ref.so0[sl], ref.so1[sl], ref.cfl[sl]
    ref.march(alpha, steps)
    left, right = (rank - 1) % world, (rank + 1) % world
    for _ in range(steps):
        svr.phase(PH_SO0, False, alpha)
        svr.phase(PH_SO1, False, alpha)
        snd_l = torch.tensor([svr.so0[3], svr.so1[3]], dtype=torch.float64)
        snd_r = torch.tensor([svr.so0[2 * m + 1], svr.so1[2 * m + 1]], dtype=torch.float64)
        rcv_l, rcv_r = torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
        if world == 1:
            rcv_l, rcv_r = snd_r, snd_l
        else:
            # same op order as solvcon_b200.decomp1d.TorchRing.exchange
            reqs = [dist.isend(snd_l, left), dist.isend(snd_r, right), dist.irecv(rcv_r, right), dist.irecv(rcv_l, left)]
            for q in reqs:
                q.wait()
        svr.so0[1], svr.so1[1] = float(rcv_l[0]), float(rcv_l[1])
        svr.so0[2 * m + 3], svr.so1[2 * m + 3] = float(rcv_r[0]), float(rcv_r[1])
        svr.phase(PH_CFL, True, alpha)
        svr.phase(PH_SO0, True, alpha)
        svr.phase(PH_CFL, False, alpha)
        svr.phase(PH_SO1, True, alpha)
    ok = True
    lo, hi = 2 * c0 + 1, 2 * c1 + 4
    for name in ("so0", "so1", "cfl"):
        a, g = getattr(svr, name)[1:-1], getattr(ref, name)[lo:hi]
        if not np.array_equal(a, g):
            ok = False
            print(f"rank {rank}: {name} differs, max abs {np.abs(a - g).max()}", flush=True)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
