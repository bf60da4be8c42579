"""Row f4 on the GPU: the oblique-shock-reflection application over the device EulerCore -- runs through on every
mesh flavour (the reference's driver tests, tests/apps/obsrefl/test_driver.py:322-353), marches bit for bit like the
CPU checker, and converges to the analytic three-zone solution (the physics-level acceptance check)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cell_type", ["quad", "triangle", "unstructured"])
def test_build_and_march(gpu, cell_type):
    from solvcon_b200.apps.obsrefl import ObliqueShock
    shock = ObliqueShock()
    shock.build_constant()
    shock.build_numerical(cell_type=cell_type, nx=24, ny=8)
    assert shock.mesh.ncell == shock.svr.ncell and shock.svr.ndim == 2
    shock.march(10)
    assert np.isfinite(shock.svr.so0n.ndarray).all()


def test_unstructured_marches_like_the_checker(gpu, bridge):
    from solvcon_b200.apps.obsrefl import ObliqueShock
    b = bridge
    shock = ObliqueShock()
    shock.build_constant()
    shock.build_numerical(cell_type="unstructured", nx=24, ny=8)
    mesh = shock.mesh
    chk = b.PortEuler(mesh.tables(), 2, 2.e-3)
    chk.phase(b.MD_PREPARE_CE)
    chk.init_solution(shock.gamma, shock.density, [shock.velocity, 0.0], shock.pressure)
    left, top, bottom, right = (mesh.bc(n).facn.ndarray[:, 0] for n in shock.mesher.BOUNDARY_NAMES)
    chk.add_bc(b.BC_INLET, left, [shock.density, shock.velocity, 0.0, shock.pressure, shock.gamma])
    chk.add_bc(b.BC_INLET, top, [shock.density2, shock.velocity2[0], shock.velocity2[1], shock.pressure2, shock.gamma])
    chk.add_bc(b.BC_SLIPWALL, bottom)
    chk.add_bc(b.BC_NONREFL, right)
    chk.phase(b.MD_BC_SOLN)
    chk.phase(b.MD_BC_DSOLN)
    chk.march(40)
    shock.march(40)
    strict = bool(gpu.load().scg_build_is_strict())
    for n in ("so0n", "so1n", "cflo", "cflc"):
        a, c = getattr(shock.svr, n).ndarray, chk.arr(n)
        if strict:
            assert np.array_equal(a, c), n
        else:
            assert np.abs(a - c).max() / (np.abs(c).max() + 1e-300) < 1e-9, n


@pytest.mark.parametrize("cell_type", ["quad", "unstructured"])
def test_converges_to_the_three_analytic_zones(gpu, cell_type):
    """Acceptance: the march reaches a steady state (unchanged between 1500 and 3000 steps) whose mean state well inside
    each zone -- 0.2 away from the analytic shock lines, i.e. outside the 3-4 cells over which the 64 x 16 mesh smears
    a shock -- is the analytic one within 1.5 % in density, pressure and velocity (measured: zone 2 within 0.1 % / 0.6 %,
    zone 3 within 1.0 % / 0.6 %), with the flow turned back to horizontal behind the reflection."""
    from solvcon_b200.apps.obsrefl import ObliqueShock, measure_zones
    shock = ObliqueShock()
    shock.build_constant()
    shock.build_numerical(cell_type=cell_type)          # the reference defaults: 64 x 16 boxes, dt = 2e-3
    shock.svr.resident(True)
    shock.march(1500)
    mid = measure_zones(shock, margin=0.2)
    shock.march(1500)
    lo, hi = shock.svr.reduce()[:2]
    assert hi < 1.0, "CFL above one"
    got = measure_zones(shock, margin=0.2)
    for a, c in zip(mid, got):
        assert np.allclose(a[1:], c[1:], rtol=1e-3, atol=1e-3), "not steady"
    want = shock.zone_states()
    for iz, ((n, rho, vx, vy, p), (wr, wvx, wvy, wp)) in enumerate(zip(got, want)):
        assert n >= 20, f"zone {iz + 1}: only {n} cells sampled"
        tol = 1.5e-2
        assert abs(rho - wr) <= tol * wr, (iz, rho, wr)
        assert abs(p - wp) <= tol * wp, (iz, p, wp)
        assert abs(vx - wvx) <= tol * wvx and abs(vy - wvy) <= tol * wvx, (iz, vx, vy)
