"""Generate tests/golden/march_*.npz: marched fields of the UNMODIFIED reference (oracle/_ref, the reference's own sources
compiled in place) from inputs stored next to them, so that the parity tests also hold on a box where neither
/root/reference nor oracle/_ref exists.  Inputs are stored as arrays (no transcendental function is re-evaluated by the
tests, numpy's sin/exp differ between CPUs).  Run from the repo root in the container that has /root/reference:

    python tests/golden/make_march_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import bridge as b  # noqa: E402
from cases import md_bump_ic, md_mesh_case, smooth_e1d_ic, sod_ic  # noqa: E402
from solvcon_b200 import meshgen  # noqa: E402

assert b.have_ref(), "needs oracle/_ref (python __graft_entry__.py)"


def e1d(name, n, dt, steps, ic):
    out = {}
    for alpha in (1, 2):
        chk = b.RefEuler1D(n, dt)
        ic(chk)
        if alpha == 1:
            out.update(coord=chk.coord.copy(), gamma=chk.gamma.copy(), so0_in=chk.so0.copy(), so1_in=chk.so1.copy(),
                       cfl_in=chk.cfl.copy(), dt=dt, steps=steps)
        chk.march(alpha, steps)
        out.update({f"so0_a{alpha}": chk.so0.copy(), f"so1_a{alpha}": chk.so1.copy(), f"cfl_a{alpha}": chk.cfl.copy()})
    np.savez_compressed(os.path.join(HERE, f"march_e1d_{name}.npz"), **out)


def scalar(eq, ncelm, steps):
    xmax = 8 * np.pi
    dt = 0.5 * xmax / ncelm
    out = {}
    for alpha in (0, 1, 2):
        chk = b.RefScalar1D(eq, 0.0, xmax, ncelm, dt)
        chk.so0[...] = 0.0
        chk.so1[...] = 0.0
        chk.cfl[...] = 0.0
        chk.so0[2::2] = 0.5 * np.sin(chk.x[2::2])
        chk.so1[2::2] = 0.5 * np.cos(chk.x[2::2])
        if alpha == 0:
            out.update(x=chk.x.copy(), so0_in=chk.so0.copy(), so1_in=chk.so1.copy(), dt=dt, steps=steps, xmax=xmax, ncelm=ncelm)
        chk.march(alpha, steps)
        out.update({f"so0_a{alpha}": chk.so0.copy(), f"so1_a{alpha}": chk.so1.copy(), f"cfl_a{alpha}": chk.cfl.copy()})
    np.savez_compressed(os.path.join(HERE, f"march_scalar_eq{eq}.npz"), **out)


def md(name, steps):
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    mesh = b.make_ref_mesh(ndim, ndcrd, cltpn, clnds)
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(ndim, mesh.body("fccnd"), mesh.body("fccls"), mesh.nface, lo, hi)
    h = (np.prod(hi - lo) / len(cltpn)) ** (1.0 / ndim)
    dt = 0.05 * h
    svr = b.RefEuler(mesh, dt)
    vel = [2.0, 0.0, 0.0][:ndim]
    svr.init_solution(1.4, 1.4, vel, 1.0)
    md_bump_ic(svr.body, mesh.body("clcnd"), ndim)
    inval = [1.4] + vel + [1.0, 1.4]
    svr.add_bc(b.BC_INLET, inlet, inval)
    svr.add_bc(b.BC_NONREFL, outlet)
    svr.add_bc(b.BC_SLIPWALL, wall)
    out = dict(ndim=ndim, ndcrd=ndcrd, cltpn=cltpn, clnds=clnds, dt=dt, steps=steps, so0n_in=svr.body("so0n").copy(),
               inlet=np.asarray(inlet, dtype=np.int32), outlet=np.asarray(outlet, dtype=np.int32),
               wall=np.asarray(wall, dtype=np.int32), inval=np.asarray(inval), vel=np.asarray(vel))
    svr.phase(b.MD_BC_SOLN)
    svr.phase(b.MD_BC_DSOLN)
    svr.march(steps)
    for n in ("so0c", "so0n", "so0t", "so1c", "so1n", "cflo", "cflc"):
        out[n] = svr.arr(n).copy()
    np.savez_compressed(os.path.join(HERE, f"march_md_{name}.npz"), **out)


if __name__ == "__main__":
    e1d("sod", 201, 1.6e-3 * 5, 40, sod_ic)
    e1d("smooth", 129, 2.0e-3, 30, smooth_e1d_ic)
    scalar(0, 96, 24)
    scalar(1, 96, 24)
    for name in ("mixed2d", "mixed3d", "pyramid3d"):
        md(name, 10)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("march_"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
