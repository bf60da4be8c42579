"""Domain decomposition on the device: several partitions of one mesh (RCB and y-slabs), marched with the halo
exchange of solvcon_b200.decomp, are bit-identical to the undecomposed run (and therefore to the checker)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ("so0n", "so1n", "cflo", "cflc")


def _ic(body_so0, clcnd, ndim, ctr, span, gamma=1.4):
    r2 = (((clcnd - ctr) / (0.25 * span)) ** 2).sum(axis=1)
    rho = 1.4 * (1.0 + 0.3 * np.exp(-r2))
    p = 1.0 * (1.0 + 0.3 * np.exp(-r2))
    v = np.zeros((len(clcnd), ndim))
    v[:, 0] = 2.0
    v[:, 1] = 0.1 * np.sin(3.0 * clcnd[:, 0])
    body_so0[:, 0] = rho
    for d in range(ndim):
        body_so0[:, 1 + d] = rho * v[:, d]
    body_so0[:, ndim + 1] = p / (gamma - 1.0) + 0.5 * rho * (v * v).sum(axis=1)


def _setup(core, mesh, groups, ndim, ctr, span, push):
    inlet, outlet, wall = groups
    vel = [2.0, 0.0, 0.0][:ndim]
    core.init_solution(1.4, 1.4, vel, 1.0)
    _ic(core.so0n.body, mesh.clcnd.body, ndim, ctr, span)
    push()
    core.add_inlet(inlet, [1.4] + vel + [1.0, 1.4])
    core.add_nonrefl(outlet)
    core.add_slipwall(wall)
    core.bc_soln()
    core.bc_dsoln()


def _run_global(ndim, ndcrd, cltpn, clnds, dt, steps):
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import EulerCore
    mesh = make_mesh(ndim, ndcrd, cltpn, clnds)
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    groups = meshgen.boundary_groups(ndim, mesh.fccnd.body, mesh.fccls.body, mesh.nface, lo, hi)
    core = EulerCore(mesh, dt)
    ctr, span = mesh.clcnd.body.mean(axis=0), hi - lo
    _setup(core, mesh, groups, ndim, ctr, span, lambda: None)
    core.march(steps)
    return {n: getattr(core, n).body.copy() for n in STATE}, (lo, hi, ctr, span)


def _run_parts(parts, dt, steps, geo, split=False):
    from solvcon_b200.decomp import DistEulerCore, LocalExchange, LocalSplitExchange
    lo, hi, ctr, span = geo
    solvers = {r: DistEulerCore(p, dt) for r, p in parts.items()}
    for r, s in solvers.items():
        _setup(s.core, s.part.mesh, s.part.physical_boundary_groups(lo, hi), s.part.ndim, ctr, span, s.upload_state)
    ex = LocalSplitExchange(solvers) if split else LocalExchange(solvers)
    for s in solvers.values():
        s.transport = ex
    ex.exchange(1)
    ex.exchange(0)
    if split:
        # the two-stage substep of the overlapped march: interface cells, pack, interior cells, unpack
        ex.march(steps)
    else:
        # lock-step march: every partition does its substep, then one exchange
        for _ in range(2 * steps):
            for s in solvers.values():
                s.core.substep_local()
            ex.exchange(0)
    out = {}
    for r, s in solvers.items():
        s.download_state()
        own = s.part.owned.astype(bool)
        out[r] = (s.part.gid[own], {n: getattr(s.core, n).body[own].copy() for n in STATE})
    return out


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("name,nparts", [("mixed3d", 2), ("mixed3d", 3), ("mixed2d", 4), ("hex3d", 2)])
def test_rcb_partitions_bit_identical(gpu, name, nparts, split):
    from cases import md_mesh_case
    from solvcon_b200.decomp import make_partitions, rcb_partition
    from solvcon_b200.mesh import make_mesh
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    h = (np.prod(ndcrd.max(axis=0) - ndcrd.min(axis=0)) / len(cltpn)) ** (1.0 / ndim)
    dt, steps = 0.05 * h, 12
    ref, geo = _run_global(ndim, ndcrd, cltpn, clnds, dt, steps)
    gm = make_mesh(ndim, ndcrd, cltpn, clnds)
    part = rcb_partition(gm.clcnd.body, nparts)
    assert len(np.unique(part)) == nparts
    parts = make_partitions(ndim, ndcrd, cltpn, clnds, part)
    res = _run_parts(parts, dt, steps, geo, split)
    seen = np.zeros(len(cltpn), dtype=int)
    for r, (gid, arrs) in res.items():
        seen[gid] += 1
        for n in STATE:
            assert np.array_equal(arrs[n], ref[n][gid]), f"rank {r}: {n} differs from the single-domain run"
    assert (seen == 1).all()
    assert np.abs(ref["so1n"]).max() > 1e-3


def test_slab_partitions_bit_identical(gpu):
    """The weak-scaling slab generator: each rank generates only its own box layers + one halo layer."""
    from solvcon_b200 import meshgen
    from solvcon_b200.decomp import slab_partition_3d
    nx, nyr, nz, nparts, h = 6, 3, 4, 3, 0.25
    x1, z1 = 4, 3
    ny = nyr * nparts
    ndcrd, cltpn, clnds, bj = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=ny * h, lz=nz * h, x1=x1, z1=z1, jitter=0.1,
                                               return_box_j=True)
    dt, steps = 0.05 * h, 10
    ref, geo = _run_global(3, ndcrd, cltpn, clnds, dt, steps)
    parts = {r: slab_partition_3d(r, nparts, nx, nyr, nz, h, x1, z1) for r in range(nparts)}
    # global ids of the slab cells: cells are generated box-major with j in the middle, so map through (i, j, k)-order
    order = np.arange(len(cltpn))
    for r, p in parts.items():
        g0, g1 = max(0, r * nyr - 2), min(ny, (r + 1) * nyr + 2)
        p.gid = order[(bj >= g0) & (bj < g1)]
    res = _run_parts(parts, dt, steps, geo)
    for r, (gid, arrs) in res.items():
        for n in STATE:
            assert np.array_equal(arrs[n], ref[n][gid]), f"slab {r}: {n} differs from the single-domain run"


@pytest.mark.parametrize("nparts,split", [(3, False), (5, True), (8, True)])
def test_lattice_rcb_partitions_bit_identical(gpu, nparts, split):
    """BASELINE config 4's decomposition: cost-weighted RCB of the stepped channel's box lattice, every rank generating
    only its own box range plus two layers (no global mesh on any rank), send / receive lists from the local face
    table -- bit-identical to the undecomposed run for 3, 5 and 8 parts."""
    from solvcon_b200 import meshgen
    from solvcon_b200.decomp import CELL_COST, lattice_partition_3d, lattice_rcb
    nx, ny, nz, x1, z1, h = 12, 9, 8, 8, 6, 0.125
    ndcrd, cltpn, clnds, box = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=ny * h, lz=nz * h, x1=x1, z1=z1, jitter=0.1,
                                                return_box=True)
    key = meshgen.cell_key(box, ny, nz)
    dt, steps = 0.05 * h, 10
    ref, geo = _run_global(3, ndcrd, cltpn, clnds, dt, steps)
    ranges = lattice_rcb(nx, ny, nz, x1, z1, nparts, cost=CELL_COST)
    parts = {r: lattice_partition_3d(r, ranges, nx, ny, nz, h, x1, z1) for r in range(nparts)}
    for p in parts.values():
        for q in p.neighbours:     # what I send is what q expects, in the same order
            assert np.array_equal(p.gid[p.send[q]], parts[q].gid[parts[q].recv[p.rank]])
    for p in parts.values():
        p.gid = np.searchsorted(key, p.gid)     # cell_key -> dense global id of the undecomposed mesh
    res = _run_parts(parts, dt, steps, geo, split)
    seen = np.zeros(len(cltpn), dtype=int)
    for r, (gid, arrs) in res.items():
        seen[gid] += 1
        for n in STATE:
            assert np.array_equal(arrs[n], ref[n][gid]), f"part {r}: {n} differs from the single-domain run"
    assert (seen == 1).all()
