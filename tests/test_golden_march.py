"""The plain-C restatement (oracle/cese_oracle.c) against marched fields of the UNMODIFIED reference committed as
fixtures (tests/golden/march_*.npz, written by tests/golden/make_march_golden.py from oracle/_ref): the oracle stays
pinned on a box that has neither /root/reference nor oracle/_ref.  Bit for bit.  CPU only."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("case", ["sod", "smooth"])
@pytest.mark.parametrize("alpha", [1, 2])
def test_port_e1d_matches_golden(bridge, case, alpha):
    g = gold(f"march_e1d_{case}.npz")
    chk = bridge.PortEuler1D(len(g["coord"]), float(g["dt"]))
    for n, src in (("coord", "coord"), ("gamma", "gamma"), ("so0", "so0_in"), ("so1", "so1_in"), ("cfl", "cfl_in")):
        getattr(chk, n)[...] = g[src]
    chk.march(alpha, int(g["steps"]))
    for n in ("so0", "so1", "cfl"):
        assert np.array_equal(getattr(chk, n), g[f"{n}_a{alpha}"]), n
    assert np.abs(g[f"so1_a{alpha}"]).max() > 1e-3       # the fixture is not a trivial field


@pytest.mark.parametrize("eq", [0, 1])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_port_scalar_matches_golden(bridge, eq, alpha):
    g = gold(f"march_scalar_eq{eq}.npz")
    chk = bridge.PortScalar1D(eq, 0.0, float(g["xmax"]), int(g["ncelm"]), float(g["dt"]))
    assert np.array_equal(chk.x, g["x"])
    chk.so0[...] = g["so0_in"]
    chk.so1[...] = g["so1_in"]
    chk.cfl[...] = 0.0
    chk.march(alpha, int(g["steps"]))
    for n in ("so0", "so1", "cfl"):
        assert np.array_equal(getattr(chk, n), g[f"{n}_a{alpha}"]), n


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d", "pyramid3d"])
def test_port_md_matches_golden(bridge, name):
    from solvcon_b200.mesh import make_mesh
    b = bridge
    g = gold(f"march_md_{name}.npz")
    ndim = int(g["ndim"])
    pm = make_mesh(ndim, g["ndcrd"], g["cltpn"], g["clnds"])     # this repo's native table builder
    chk = b.PortEuler(pm.tables(), ndim, float(g["dt"]))
    chk.phase(b.MD_PREPARE_CE)
    chk.init_solution(1.4, 1.4, g["vel"].tolist(), 1.0)
    chk.body("so0n")[...] = g["so0n_in"]
    chk.add_bc(b.BC_INLET, g["inlet"], g["inval"].tolist())
    chk.add_bc(b.BC_NONREFL, g["outlet"])
    chk.add_bc(b.BC_SLIPWALL, g["wall"])
    chk.phase(b.MD_BC_SOLN)
    chk.phase(b.MD_BC_DSOLN)
    chk.march(int(g["steps"]))
    for n in ("so0c", "so0n", "so0t", "so1c", "so1n", "cflo", "cflc"):
        assert np.array_equal(chk.arr(n), g[n]), n
    assert np.abs(g["so1n"]).max() > 1e-3
