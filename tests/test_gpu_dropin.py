"""The reference's OWN front ends, unmodified (``dropin/_stage/solvcon``, staged by ``dropin/stage.py``), running on
the device cores through ``dropin/_solvcon`` -- bit for bit against the CPU checker on identical inputs, plus the
physics-level acceptance checks of SURVEY.md 8f (Sod tube against its analytic solution, oblique-shock reflection
against its three analytic zones).  The reference's own unit tests for the same classes run from ``tests/_refsuite``."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.isdir(os.path.join(os.path.dirname(HERE), "dropin", "_stage", "solvcon"))
needs_stage = pytest.mark.skipif(not STAGED, reason="dropin/_stage not staged (dropin/stage.py needs /root/reference)")


def _assert_same(a, c, name, strict, tol=1e-10):
    a, c = np.asarray(a), np.asarray(c)
    if strict:
        assert np.array_equal(a, c), name
    else:
        assert np.abs(a - c).max() <= tol * (np.abs(c).max() + 1e-300), name


@needs_stage
def test_reference_package_is_served_by_the_device_library(gpu):
    import solvcon
    import solvcon_b200.multidim
    import solvcon_b200.onedim
    assert solvcon.__file__.startswith(os.path.join(os.path.dirname(HERE), "dropin", "_stage"))
    assert solvcon.EulerCore is solvcon_b200.multidim.EulerCore
    from solvcon.onedim import euler1d
    assert euler1d._impl.Euler1DCore is solvcon_b200.onedim.Euler1DCore


@needs_stage
def test_shocktube_front_end(gpu, bridge):
    """``solvcon.onedim.euler1d.ShockTube`` (reference file, BASELINE config 1's front end): numerical Sod tube on the
    device core, same initial condition bit-exact against the CPU checker, and against the analytic solution."""
    from solvcon.onedim.euler1d import ShockTube
    n, dt, steps = 1001, 1.6e-3, 100
    st = ShockTube()
    st.build_constant(gamma=1.4, pressure1=1.0, density1=1.0, pressure5=0.1, density5=0.125)
    st.build_numerical(xmin=-1, xmax=1, ncoord=n, time_increment=dt)
    svr = st.svr
    chk = bridge.RefEuler1D(n, dt) if bridge.have_ref() else bridge.PortEuler1D(n, dt)
    for name in ("coord", "so0", "so1", "cfl"):      # cfl: build_numerical already ran setup_march on the device
        getattr(chk, name)[...] = getattr(svr, name)
    chk.gamma.fill(1.4)
    svr.march_alpha2(steps)
    chk.march(2, steps)
    strict = bool(gpu.load().scg_build_is_strict())
    for name in ("so0", "so1", "cfl"):
        _assert_same(getattr(svr, name), getattr(chk, name), name, strict)
    st.build_field(t=dt * steps)
    xi = svr.xindices
    # first-order-accurate capture of the three waves: mean error of a few 1e-3 at 500 cells
    assert np.abs(svr.density[xi] - st.density_field).mean() < 4e-3
    assert np.abs(svr.velocity[xi] - st.velocity_field).mean() < 6e-3
    assert np.abs(svr.pressure[xi] - st.pressure_field).mean() < 4e-3


@needs_stage
def test_linear_scalar_front_end(gpu, bridge):
    """``solvcon.onedim.linear_scalar.LinearScalarSolver`` (reference file; the pilot's default case of 4 x 64 elements
    on [0, 8 pi]): march bit-exact against the checker and the CFL = 1 property -- the pulse travels one element per
    step without changing shape."""
    from solvcon.onedim.linear_scalar import LinearScalarSolver
    ncelm, xmax = 4 * 64, 4 * 2 * np.pi
    svr = LinearScalarSolver(xmin=0, xmax=xmax, ncelm=ncelm, cfl=1)
    svr.setup_march()
    u0 = np.asarray(svr.get_so0(0)).copy()
    dt = (svr.grid.xmax - svr.grid.xmin) / ncelm
    chk = (bridge.RefScalar1D if bridge.have_ref() else bridge.PortScalar1D)(0, 0.0, xmax, ncelm, dt)
    chk.so0[...] = svr.so0.ndarray[:, 0]
    chk.so1[...] = svr.so1.ndarray[:, 0]
    chk.cfl[...] = svr.cfl.ndarray
    steps = 32
    svr.march_alpha2(steps)
    chk.march(2, steps)
    strict = bool(gpu.load().scg_build_is_strict())
    _assert_same(svr.so0.ndarray[:, 0], chk.so0, "so0", strict)
    _assert_same(svr.so1.ndarray[:, 0], chk.so1, "so1", strict)
    _assert_same(svr.cfl.ndarray, chk.cfl, "cfl", strict)
    u = np.asarray(svr.get_so0(0))
    np.testing.assert_allclose(u[:-1], np.roll(u0[:-1], steps), rtol=0, atol=1e-12)


def _checker_for(shock, b):
    """The CPU checker set up exactly like ``ObliqueShock.build_numerical`` (``_driver.py:470-509``) on the tables of
    the driver's own mesh."""
    mesh = shock.mesh
    chk = b.PortEuler(mesh.tables(), 2, shock.svr.time_increment)
    chk.phase(b.MD_PREPARE_CE)
    chk.init_solution(shock.gamma, shock.density, [shock.velocity, 0.0], shock.pressure)
    left, top, bottom, right = (mesh.bc(n).facn.ndarray[:, 0] for n in shock.mesher.BOUNDARY_NAMES)
    chk.add_bc(b.BC_INLET, left, [shock.density, shock.velocity, 0.0, shock.pressure, shock.gamma])
    chk.add_bc(b.BC_INLET, top, [shock.density2, shock.velocity2[0], shock.velocity2[1], shock.pressure2, shock.gamma])
    chk.add_bc(b.BC_SLIPWALL, bottom)
    chk.add_bc(b.BC_NONREFL, right)
    chk.phase(b.MD_BC_SOLN)
    chk.phase(b.MD_BC_DSOLN)
    return chk


@needs_stage
@pytest.mark.parametrize("cell_type", ["quad", "triangle", "unstructured"])
def test_obsrefl_driver_marches_like_the_checker(gpu, bridge, cell_type):
    """Row f4: the reference's ``_driver.py`` (mesher + ObliqueShock), unmodified, over the device EulerCore: 40 steps
    bit for bit like the checker on the same mesh tables, every mesh flavour."""
    from solvcon.pilot.apps.obsrefl import ObliqueShock
    shock = ObliqueShock()
    shock.build_constant()
    shock.build_numerical(cell_type=cell_type, nx=24, ny=8)
    chk = _checker_for(shock, bridge)
    chk.march(40)
    shock.march(40)
    strict = bool(gpu.load().scg_build_is_strict())
    for n in ("so0n", "so1n", "cflo", "cflc"):
        _assert_same(getattr(shock.svr, n).ndarray, chk.arr(n), n, strict, tol=1e-9)


@needs_stage
def test_obsrefl_reaches_the_analytic_zones(gpu):
    """Acceptance: the reference's driver + its ``Reflection`` analysis (``_analytic.py``), unmodified, on the device:
    the default 64 x 16 case marched to steady state recovers the three analytic zone states within 1.5 %."""
    from solvcon.pilot.apps.obsrefl import ObliqueShock, Reflection
    shock = ObliqueShock()
    shock.build_constant()
    shock.build_numerical()
    shock.svr.resident(True)
    shock.march(3000)
    shock.svr.resident(False)        # back to host-visible arrays: the analysis reads them through EulerField
    refl = Reflection(shock, margin=0.2)
    lo, hi = refl.field.calc_cfl_range()
    assert 0.0 < lo <= hi < 1.0
    masks = refl.zone_masks()
    want = refl.zone_fields()
    for iz, mask in enumerate(masks):
        assert mask.sum() >= 20, f"zone {iz + 1}: only {mask.sum()} cells sampled"
        for name in ("density", "pressure", "velx"):
            got = float(refl.field.field(name)[mask].mean())
            assert abs(got - want[name][iz]) <= 1.5e-2 * abs(want[name][iz]), (iz, name, got, want[name][iz])


@needs_stage
def test_euler_field_reads_the_device_core(gpu, bridge):
    """``solvcon.multidim.euler.EulerField`` (reference file) over the device EulerCore agrees with the device-side
    derived-field kernels (EulerCore.derived) element for element."""
    from cases import md_mesh_case
    from solvcon.multidim.euler import EulerField
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import EulerCore
    ndim, ndcrd, cltpn, clnds = md_mesh_case("mixed2d")
    mesh = make_mesh(ndim, ndcrd, cltpn, clnds)
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(ndim, mesh.fccnd.body, mesh.fccls.body, mesh.nface, lo, hi)
    svr = EulerCore(mesh, 0.05 * (np.prod(hi - lo) / mesh.ncell) ** 0.5)
    svr.init_solution(1.4, 1.4, [3.0, 0.0], 1.0)
    svr.add_inlet(inlet, [1.4, 3.0, 0.0, 1.0, 1.4])
    svr.add_nonrefl(outlet)
    svr.add_slipwall(wall)
    svr.bc_soln()
    svr.bc_dsoln()
    svr.march(10)
    fld = EulerField(svr, mesh)
    strict = bool(gpu.load().scg_build_is_strict())
    for name in fld.FIELDS:
        _assert_same(svr.derived(name), fld.field(name), name, strict, tol=1e-12)
    lo_, hi_, mass = svr.reduce(mesh.clvol.ndarray)
    assert (lo_, hi_) == tuple(fld.calc_cfl_range())
    assert abs(mass - fld.calc_overall_mass()) <= 1e-13 * abs(mass)
