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


def _checker_3d(ncell_target, kinds="hpt"):
    """The CPU checker (oracle/_ref, else the pinned C port) set up on the bench's 3D generator, IC and BCs."""
    from oracle import bridge as b
    from solvcon_b200.bench_md import GAMMA, MACH_U, P0, RHO0, pocket_at_rest
    nx, ny, nz, x1, z1 = dims_3d(ncell_target)
    h = 1.0 / ny
    ndcrd, cltpn, clnds = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=1.0, lz=nz * h, x1=x1, z1=z1, kinds=kinds, jitter=0.1)
    dt = 0.05 * h
    if b.have_ref():
        mesh = b.make_ref_mesh(3, ndcrd, cltpn, clnds)
        svr = b.RefEuler(mesh, dt)
        fccnd, fccls, nface, clcnd = mesh.body("fccnd"), mesh.body("fccls"), mesh.nface, mesh.body("clcnd")
        kind = "reference"
    else:
        pm = make_mesh(3, ndcrd, cltpn, clnds)
        svr = b.PortEuler(pm.tables(), 3, dt)
        svr.phase(b.MD_PREPARE_CE)
        fccnd, fccls, nface, clcnd = pm.fccnd.body, pm.fccls.body, pm.nface, pm.clcnd.body
        kind = "port"
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(3, fccnd, fccls, nface, lo, hi)
    svr.init_solution(GAMMA, RHO0, [MACH_U, 0.0, 0.0], P0)
    if kinds == "hpt":
        pocket_at_rest(svr.body("so0n"), clcnd, z1 * h)
    svr.add_bc(b.BC_INLET, inlet, [RHO0, MACH_U, 0.0, 0.0, P0, GAMMA])
    svr.add_bc(b.BC_NONREFL, outlet)
    svr.add_bc(b.BC_SLIPWALL, wall)
    svr.phase(b.MD_BC_SOLN)
    svr.phase(b.MD_BC_DSOLN)
    return svr, kind, len(cltpn)


def cpu_reference_3d(budget_s=15.0, kinds="hpt", ncell_target=60000):
    """The reference's own EulerCore.march (oracle/_ref, else the pinned C port) on the same generator at a
    size that finishes in seconds, single thread (the reference is single-threaded)."""
    svr, kind, ncell = _checker_3d(ncell_target, kinds)
    svr.march(1)
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


def _compare(dev, chk, kind, ncell, steps, strict):
    """Bit comparison of the device field with the checker's after the same march (reference cell order)."""
    import numpy as np
    worst, equal = 0.0, True
    for name in ("so0n", "so1n", "cflo", "cflc"):
        a, c = getattr(dev, name).ndarray, chk.arr(name)
        equal = equal and bool(np.array_equal(a, c))
        scale = float(np.abs(c).max()) + 1e-300
        worst = max(worst, float(np.abs(a - c).max()) / scale)
    return {"cells": int(ncell), "steps": int(steps), "bit_equal": equal, "max_rel_linf": worst, "checker": kind,
            "arrays": ["so0n", "so1n", "cflo", "cflc"], "build": "strict (-fmad=false)" if strict else "fma variant"}


def parity_3d(order, kinds="hpt", ncell_target=60000, steps=12):
    """In-run parity leg of bench.py: the bench's own generator, IC and BCs at a size the CPU checker finishes in
    seconds, marched by the device EulerCore and by the checker; every element of so0n, so1n, cflo, cflc compared."""
    from solvcon_b200 import _lib
    from solvcon_b200.bench_md import build_case_3d
    chk, kind, ncell = _checker_3d(ncell_target, kinds)
    _, dev, info = build_case_3d(ncell_target, order, kinds)
    assert info["ncell"] == ncell
    chk.march(steps)
    dev.march(steps)
    dev.materialise()
    dev.sync_to_host()
    return _compare(dev, chk, kind, ncell, steps, bool(_lib.load().scg_build_is_strict()))


def _checker_2d(ncell_target):
    from oracle import bridge as b
    from solvcon_b200.bench_md import GAMMA, MACH_U, P0, RHO0
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
    svr.init_solution(GAMMA, RHO0, [MACH_U, 0.0], P0)
    svr.add_bc(b.BC_INLET, inlet, [RHO0, MACH_U, 0.0, P0, GAMMA])
    svr.add_bc(b.BC_NONREFL, outlet)
    svr.add_bc(b.BC_SLIPWALL, wall)
    svr.phase(b.MD_BC_SOLN)
    svr.phase(b.MD_BC_DSOLN)
    return svr, kind, len(cltpn)


def cpu_reference_2d(budget_s=15.0, ncell_target=120000):
    """The reference's EulerCore.march on the same 2D generator at a size that finishes in seconds, single thread."""
    svr, kind, ncell = _checker_2d(ncell_target)
    svr.march(1)
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


def parity_2d(order, ncell_target=120000, steps=25):
    from solvcon_b200 import _lib
    from solvcon_b200.bench_md import build_case_2d
    chk, kind, ncell = _checker_2d(ncell_target)
    _, dev, info = build_case_2d(ncell_target, order)
    assert info["ncell"] == ncell
    chk.march(steps)
    dev.march(steps)
    dev.materialise()
    dev.sync_to_host()
    return _compare(dev, chk, kind, ncell, steps, bool(_lib.load().scg_build_is_strict()))
