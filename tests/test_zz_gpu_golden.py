"""The device path against marched fields of the UNMODIFIED reference committed as fixtures (tests/golden/march_*.npz,
written by tests/golden/make_march_golden.py from oracle/_ref), inputs taken from the same files: parity that needs no
checker at run time.  Through the C-ABI like every GPU test; bit for bit in the strict build."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def _same(a, b, name, strict):
    a, b = np.asarray(a), np.asarray(b)
    if strict:
        assert np.array_equal(a, b), f"{name}: max abs diff {np.abs(a - b).max()}"
    else:   # FMA speed variant, not the product default
        assert np.abs(a - b).max() <= 1e-9 * (np.abs(b).max() + 1e-300), name


def _strict(gpu):
    return bool(gpu.load().scg_build_is_strict())


@pytest.mark.parametrize("case", ["sod", "smooth"])
@pytest.mark.parametrize("alpha", [1, 2])
def test_e1d_matches_golden(gpu, case, alpha):
    from solvcon_b200.onedim import Euler1DCore
    g = gold(f"march_e1d_{case}.npz")
    dev = Euler1DCore(len(g["coord"]), float(g["dt"]))
    dev.coord[...] = g["coord"]
    dev.gamma[...] = g["gamma"]
    dev.so0[...] = g["so0_in"]
    dev.so1[...] = g["so1_in"]
    dev.cfl[...] = g["cfl_in"]
    getattr(dev, f"march_alpha{alpha}")(int(g["steps"]))
    for n in ("so0", "so1", "cfl"):
        _same(getattr(dev, n), g[f"{n}_a{alpha}"], n, _strict(gpu))


@pytest.mark.parametrize("eq", [0, 1])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_scalar_matches_golden(gpu, eq, alpha):
    from solvcon_b200 import spacetime as st
    g = gold(f"march_scalar_eq{eq}.npz")
    grid = st.Grid(0.0, float(g["xmax"]), int(g["ncelm"]))
    assert np.array_equal(grid.xcoord.ndarray, g["x"])
    dev = (st.LinearScalarSolver if eq == 0 else st.InviscidBurgersSolver)(grid, float(g["dt"]))
    dev.so0.ndarray[:, 0] = g["so0_in"]
    dev.so1.ndarray[:, 0] = g["so1_in"]
    dev.cfl.ndarray[...] = 0.0
    getattr(dev, f"march_alpha{alpha}")(int(g["steps"]))
    strict = _strict(gpu)
    _same(dev.so0.ndarray[:, 0], g[f"so0_a{alpha}"], "so0", strict)
    _same(dev.so1.ndarray[:, 0], g[f"so1_a{alpha}"], "so1", strict)
    _same(dev.cfl.ndarray, g[f"cfl_a{alpha}"], "cfl", strict)


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d", "pyramid3d"])
@pytest.mark.parametrize("order", ["morton", "rcm"])
def test_md_matches_golden(gpu, name, order):
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import EulerCore
    g = gold(f"march_md_{name}.npz")
    ndim = int(g["ndim"])
    mesh = make_mesh(ndim, g["ndcrd"], g["cltpn"], g["clnds"])
    dev = EulerCore(mesh, float(g["dt"]), order=order)
    dev.init_solution(1.4, 1.4, g["vel"].tolist(), 1.0)
    dev.so0n.body[...] = g["so0n_in"]
    dev.add_inlet(g["inlet"], g["inval"].tolist())
    dev.add_nonrefl(g["outlet"])
    dev.add_slipwall(g["wall"])
    dev.bc_soln()
    dev.bc_dsoln()
    dev.march(int(g["steps"]))
    strict = _strict(gpu)
    for n in ("so0c", "so0n", "so0t", "so1c", "so1n", "cflo", "cflc"):
        _same(getattr(dev, n).ndarray, g[n], n, strict)
