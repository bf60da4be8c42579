"""CPU reference arm of bench.py for the multidim workloads (``cpu_baseline`` and ``--impl reference``): the reference's
own ``EulerCore.march`` (``oracle/_ref``: the unmodified reference sources compiled in place; else the pinned plain-C
port) on the bench's mesh generators at a size that finishes in seconds, single thread (the reference is
single-threaded).  Together with bench.py's 1D workloads this is the only code outside ``tests/`` and ``smoke()`` that
executes anything under ``oracle/`` -- as the thing timed BESIDE the product, never inside it."""
from __future__ import annotations

import time

from solvcon_b200 import meshgen
from solvcon_b200.bench_md import UNIT, dims_2d, dims_3d
from solvcon_b200.mesh import make_mesh


def cpu_reference_3d(budget_s=15.0, kinds="hpt", ncell_target=60000):
    """The reference's own EulerCore.march (oracle/_ref, else the pinned C port) on the same generator at a
    size that finishes in seconds, single thread (the reference is single-threaded)."""
    from oracle import bridge as b
    nx, ny, nz, x1, z1 = dims_3d(ncell_target)
    h = 1.0 / ny
    ndcrd, cltpn, clnds = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=1.0, lz=nz * h, x1=x1, z1=z1, kinds=kinds, jitter=0.1)
    dt = 0.05 * h
    if b.have_ref():
        mesh = b.make_ref_mesh(3, ndcrd, cltpn, clnds)
        svr = b.RefEuler(mesh, dt)
        fccnd, fccls, nface = mesh.body("fccnd"), mesh.body("fccls"), mesh.nface
        kind = "reference"
    else:
        pm = make_mesh(3, ndcrd, cltpn, clnds)
        svr = b.PortEuler(pm.tables(), 3, dt)
        svr.phase(b.MD_PREPARE_CE)
        fccnd, fccls, nface = pm.fccnd.body, pm.fccls.body, pm.nface
        kind = "port"
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(3, fccnd, fccls, nface, lo, hi)
    svr.init_solution(1.4, 1.4, [3.0, 0.0, 0.0], 1.0)
    svr.add_bc(b.BC_INLET, inlet, [1.4, 3.0, 0.0, 0.0, 1.0, 1.4])
    svr.add_bc(b.BC_NONREFL, outlet)
    svr.add_bc(b.BC_SLIPWALL, wall)
    svr.phase(b.MD_BC_SOLN)
    svr.phase(b.MD_BC_DSOLN)
    svr.march(1)
    ncell = len(cltpn)
    t0 = time.perf_counter()
    steps = 0
    while True:
        svr.march(1)
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 200:
            break
    upd = 2 * steps * ncell
    return ({"value": upd / el, "unit": UNIT, "cores": 1, "kind": kind,
             "sample": f"3D Euler mixed tet/hex/prism channel, {ncell} cells, march({steps}) in {el:.2f} s, single "
                       f"thread (the reference is single-threaded)"}, upd, el)


def cpu_reference_2d(budget_s=15.0, ncell_target=120000):
    """The reference's EulerCore.march on the same 2D generator at a size that finishes in seconds, single thread."""
    from oracle import bridge as b
    nx, ny = dims_2d(ncell_target)
    h = 1.0 / ny
    ndcrd, cltpn, clnds = meshgen.mixed_2d(nx, ny, lx=3.0, ly=1.0, tri_every=4, jitter=0.3, step=(0.6, 0.2))
    dt = 0.05 * h
    if b.have_ref():
        mesh = b.make_ref_mesh(2, ndcrd, cltpn, clnds)
        svr = b.RefEuler(mesh, dt)
        fccnd, fccls, nface = mesh.body("fccnd"), mesh.body("fccls"), mesh.nface
        kind = "reference"
    else:
        pm = make_mesh(2, ndcrd, cltpn, clnds)
        svr = b.PortEuler(pm.tables(), 2, dt)
        svr.phase(b.MD_PREPARE_CE)
        fccnd, fccls, nface = pm.fccnd.body, pm.fccls.body, pm.nface
        kind = "port"
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(2, fccnd, fccls, nface, lo, hi)
    svr.init_solution(1.4, 1.4, [3.0, 0.0], 1.0)
    svr.add_bc(b.BC_INLET, inlet, [1.4, 3.0, 0.0, 1.0, 1.4])
    svr.add_bc(b.BC_NONREFL, outlet)
    svr.add_bc(b.BC_SLIPWALL, wall)
    svr.phase(b.MD_BC_SOLN)
    svr.phase(b.MD_BC_DSOLN)
    svr.march(1)
    ncell = len(cltpn)
    t0 = time.perf_counter()
    steps = 0
    while True:
        svr.march(1)
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 400:
            break
    upd = 2 * steps * ncell
    return ({"value": upd / el, "unit": UNIT, "cores": 1, "kind": kind,
             "sample": f"2D Euler mixed tri/quad forward step, {ncell} cells, march({steps}) in {el:.2f} s, single "
                       f"thread (the reference is single-threaded)"}, upd, el)
