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
