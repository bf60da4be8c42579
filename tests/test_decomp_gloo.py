"""World_size-2/3 gloo jobs on the CPU: the product's HaloSolver + TorchExchange (solvcon_b200/decomp.py: the classes
that run over NCCL on the GPUs) reproduce the undecomposed march bit for bit, plain and two-stage (overlap) order.
Compute engine: the CPU checker behind a host double of the device core's halo interface (tests/decomp_worker.py);
the single-GPU twin of this test is tests/test_gpu_decomp.py, the NCCL one is bench.py's multi-rank parity leg."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_rcb_partition_balanced_and_deterministic():
    from solvcon_b200.decomp import rcb_partition
    rng = np.random.RandomState(0)
    c = rng.rand(1001, 3)
    for n in (1, 2, 3, 5, 8):
        part = rcb_partition(c, n)
        cnt = np.bincount(part, minlength=n)
        assert cnt.sum() == 1001 and cnt.max() - cnt.min() <= n
        assert np.array_equal(part, rcb_partition(c, n))


def test_partition_lists_are_consistent():
    from cases import md_mesh_case
    from solvcon_b200.decomp import make_partitions, rcb_partition
    from solvcon_b200.mesh import make_mesh
    ndim, ndcrd, cltpn, clnds = md_mesh_case("mixed3d")
    gm = make_mesh(ndim, ndcrd, cltpn, clnds)
    part = rcb_partition(gm.clcnd.body, 4)
    ps = make_partitions(ndim, ndcrd, cltpn, clnds, part)
    owned_total = 0
    for r, p in ps.items():
        own = p.owned.astype(bool)
        owned_total += own.sum()
        assert np.all(np.diff(p.gid) > 0)                       # ascending global order
        assert np.array_equal(part[p.gid[own]], np.full(own.sum(), r))
        for q in p.neighbours:                                   # what I send is what q expects, in the same order
            assert np.array_equal(p.gid[p.send[q]], ps[q].gid[ps[q].recv[r]])
            assert p.owned[p.send[q]].all() and not p.owned[p.recv[q]].any()
        lm = p.build_mesh()
        # mesh centroids of owned cells and first-layer copies carry the global mesh's bits
        first = np.concatenate([p.recv[q] for q in p.neighbours])
        chk = np.concatenate([np.nonzero(own)[0], first])
        assert np.array_equal(lm.clcnd.body[chk], gm.clcnd.body[p.gid[chk]])
    assert owned_total == len(cltpn)


@pytest.mark.parametrize("name,world,mode", [("mixed3d", 2, "plain"), ("mixed2d", 3, "plain"), ("mixed3d", 3, "split")])
def test_gloo_march_bit_identical(bridge, name, world, mode):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29400 + world + len(name) + (7 if mode == "split" else 0)),
           os.path.join(HERE, "decomp_worker.py"), name, "6", mode]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_slab_ranges_cover_the_grid():
    from solvcon_b200.decomp1d import slab_ranges
    for n, p in ((10, 1), (10, 3), (1_000_003, 8), (7, 7)):
        r = slab_ranges(n, p)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("eq,world", [(0, 2), (1, 3)])
def test_gloo_1d_ring_bit_identical(bridge, eq, world):
    """1D scalar slabs over a gloo ring: same message order as TorchRing, CPU checker as the engine."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29440 + world + eq),
           os.path.join(HERE, "decomp1d_worker.py"), str(eq), "2003", "30"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_lattice_rcb_covers_and_balances():
    """The closed-form box-lattice RCB: parts tile the lattice, every part holds cells, cost-weighted sizes agree
    within a box layer, and the result does not depend on who computes it."""
    from solvcon_b200.bench_md import dims_3d
    from solvcon_b200.decomp import CELL_COST, lattice_rcb, lattice_weight
    for target, nparts in ((40_000, 3), (1_000_000, 8), (100_000_000, 8), (100_000_000, 2)):
        nx, ny, nz, x1, z1 = dims_3d(target)
        rs = lattice_rcb(nx, ny, nz, x1, z1, nparts, cost=CELL_COST)
        assert rs == lattice_rcb(nx, ny, nz, x1, z1, nparts, cost=CELL_COST)
        cells = [lattice_weight(r, x1, z1) for r in rs]
        assert sum(cells) == lattice_weight(((0, nx), (0, ny), (0, nz)), x1, z1) and min(cells) > 0
        cost = [lattice_weight(r, x1, z1, CELL_COST) for r in rs]
        assert max(cost) / (sum(cost) / nparts) < (1.01 if target >= 100_000_000 else 1.05 if target >= 1_000_000 else 1.25)


def test_lattice_partition_lists_are_consistent():
    from solvcon_b200 import meshgen
    from solvcon_b200.bench_md import dims_3d
    from solvcon_b200.decomp import CELL_COST, lattice_partition_3d, lattice_rcb
    nx, ny, nz, x1, z1 = dims_3d(30000)
    h = 1.0 / ny
    rs = lattice_rcb(nx, ny, nz, x1, z1, 5, cost=CELL_COST)
    ps = [lattice_partition_3d(r, rs, nx, ny, nz, h, x1, z1) for r in range(5)]
    nd, tp, cl, box = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=1.0, lz=nz * h, x1=x1, z1=z1, jitter=0.1, return_box=True)
    key = meshgen.cell_key(box, ny, nz)
    gm = __import__("solvcon_b200.mesh", fromlist=["make_mesh"]).make_mesh(3, nd, tp, cl)
    assert sum(int(p.owned.sum()) for p in ps) == len(tp)
    for p in ps:
        assert np.all(np.diff(p.gid) > 0)
        for q in p.neighbours:
            assert np.array_equal(p.gid[p.send[q]], ps[q].gid[ps[q].recv[p.rank]])
            assert p.owned[p.send[q]].all() and not p.owned[p.recv[q]].any()
        # mesh centroids of the owned cells and of the exchanged copies carry the global mesh's bits
        first = np.concatenate([p.recv[q] for q in p.neighbours])
        chk = np.concatenate([np.nonzero(p.owned)[0], first])
        assert np.array_equal(p.build_mesh().clcnd.body[chk], gm.clcnd.body[np.searchsorted(key, p.gid[chk])])
