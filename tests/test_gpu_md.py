"""GPU parity of the multidim device path against the CPU checkers, through the C-ABI of include/scg.h
(via solvcon_b200.multidim.EulerCore).  Strict build: bit-equal, array by array, phase by phase."""
import numpy as np
import pytest

from cases import md_bump_ic, md_mesh_case

pytestmark = pytest.mark.gpu

MESHES = ["quad2d", "tri2d", "mixed2d", "step2d", "hex3d", "tet3d", "prism3d", "mixed3d", "pyramid3d"]
STATE = ("so0c", "so0n", "so0t", "so1c", "so1n", "cflo", "cflc", "gamma")
GEOM = ("cevol", "cecnd", "sfcnd", "sfnml")


class Case:
    """One mesh + IC + BCs set up identically on the checker and on the device solver."""

    def __init__(self, b, name, order="morton", dt_scale=0.05):
        from solvcon_b200 import meshgen
        from solvcon_b200.mesh import TableMesh
        from solvcon_b200.multidim import EulerCore
        ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
        self.ndim = ndim
        if b.have_ref():
            rmesh = b.make_ref_mesh(ndim, ndcrd, cltpn, clnds)
            tabs = rmesh.tables()
            self.rmesh = rmesh
        else:
            pytest.skip("needs oracle/_ref (mesh tables come from the reference's StaticMesh)")
        ng = tabs["fccls"][1]
        nface = len(tabs["fccls"][0]) - ng
        fccnd = tabs["fccnd"][0][tabs["fccnd"][1]:]
        fccls = tabs["fccls"][0][ng:]
        clcnd = tabs["clcnd"][0][tabs["clcnd"][1]:]
        lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
        self.groups = meshgen.boundary_groups(ndim, fccnd, fccls, nface, lo, hi)
        inlet, outlet, wall = self.groups
        h = (np.prod(hi - lo) / len(cltpn)) ** (1.0 / ndim)
        dt = dt_scale * h
        self.chk = b.PortEuler(tabs, ndim, dt)
        self.chk.phase(b.MD_PREPARE_CE)
        self.mesh = TableMesh(ndim, tabs)
        self.dev = EulerCore(self.mesh, dt, order=order)
        vel = [2.0, 0.0, 0.0][:ndim]
        self.chk.init_solution(1.4, 1.4, vel, 1.0)
        self.dev.init_solution(1.4, 1.4, vel, 1.0)
        md_bump_ic(self.chk.body, clcnd, ndim)
        md_bump_ic(lambda n: getattr(self.dev, n).body, clcnd, ndim)
        inval = [1.4] + vel + [1.0, 1.4]
        self.chk.add_bc(b.BC_INLET, inlet, inval)
        self.chk.add_bc(b.BC_NONREFL, outlet)
        self.chk.add_bc(b.BC_SLIPWALL, wall)
        self.dev.add_inlet(inlet, inval)
        self.dev.add_nonrefl(outlet)
        self.dev.add_slipwall(wall)

    def same(self, tag, names=STATE, strict=True):
        for n in names:
            a, c = getattr(self.dev, n).ndarray, self.chk.arr(n)
            if strict:
                assert np.array_equal(a, c), f"{tag}: {n} differs, max abs {np.abs(a - c).max()}"
            else:
                scale = np.abs(c).max() + 1e-300
                assert np.abs(a - c).max() / scale < 1e-10, f"{tag}: {n}"


def _strict(gpu):
    return bool(gpu.load().scg_build_is_strict())


@pytest.mark.parametrize("name", MESHES)
def test_md_geometry_and_phases(gpu, bridge, name):
    """prepare_ce and every phase of three substeps, like the reference's per-phase tests
    (tests/test_multidim.py:632-770) but against the checker's arrays, bit for bit."""
    b = bridge
    c = Case(b, name)
    strict = _strict(gpu)
    c.same("prepare_ce", GEOM, strict)
    c.chk.phase(b.MD_BC_SOLN)
    c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    c.same("prime", STATE, strict)
    calls = (("update", b.MD_UPDATE), ("calc_solt", b.MD_SOLT), ("calc_soln", b.MD_SOLN), ("bc_soln", b.MD_BC_SOLN),
             ("calc_cfl", b.MD_CFL), ("calc_dsoln", b.MD_DSOLN), ("bc_dsoln", b.MD_BC_DSOLN))
    for it in range(3):
        for fn, ph in calls:
            c.chk.phase(ph)
            getattr(c.dev, fn)()
            c.same(f"substep {it} {fn}", STATE, strict)
    assert np.abs(c.dev.so1n.body).max() > 1e-3


@pytest.mark.parametrize("name", ["mixed2d", "step2d", "mixed3d", "hex3d", "tet3d", "prism3d", "tri2d", "pyramid3d"])
@pytest.mark.parametrize("order", ["none", "morton", "rcm"])
@pytest.mark.parametrize("fused", [3, 2, 1, 0])
def test_md_march_matches(gpu, bridge, name, order, fused):
    """march(N) with every cell ordering, fused and per-phase kernels: identical to the checker."""
    b = bridge
    c = Case(b, name, order=order)
    strict = _strict(gpu)
    c.dev.set_fused(fused)
    for s in (c.chk,):
        s.phase(b.MD_BC_SOLN)
        s.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    n = 25
    c.chk.march(n)
    c.dev.march(n)
    c.same(f"march({n}) order={order} fused={fused}", STATE, strict)
    assert np.isfinite(c.dev.so0n.ndarray).all()
    # conserved totals equal to round-off
    vol = c.mesh.clvol.body
    tot_d = (c.dev.so0n.body * vol[:, None]).sum(axis=0)
    tot_c = (c.chk.body("so0n") * vol[:, None]).sum(axis=0)
    assert np.allclose(tot_d, tot_c, rtol=1e-13, atol=1e-13)


def test_md_substep_and_swap(gpu, bridge):
    """march_substep swaps the current/next buffers on both sides (euler.hpp:123-127)."""
    b = bridge
    c = Case(b, "mixed2d")
    c.dev.bc_soln(); c.dev.bc_dsoln()
    c.chk.phase(b.MD_BC_SOLN); c.chk.phase(b.MD_BC_DSOLN)
    view_n = c.dev.so0n.ndarray
    for it in range(3):
        c.chk.phase(b.MD_SUBSTEP)
        c.dev.march_substep()
        c.same(f"march_substep {it}", STATE, _strict(gpu))
    # three swaps later the old so0n buffer is the current one
    assert view_n is c.dev.so0c.ndarray


def test_md_reduce_resident_and_params(gpu, bridge):
    b = bridge
    c = Case(b, "mixed3d")
    c.dev.sigma0, c.dev.taumin, c.dev.tauscale = 2.0, 0.1, 0.7
    c.chk.set_params(2.0, 0.1, 0.7)
    c.chk.phase(b.MD_BC_SOLN); c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln(); c.dev.bc_dsoln()
    c.dev.resident(True)
    c.dev.march(4)
    mn, mx, mass = c.dev.reduce(c.mesh.clvol.ndarray)
    c.dev.sync_to_host()
    c.chk.march(4)
    c.same("resident march(4) with sigma0=2, taumin=.1, tauscale=.7", STATE, _strict(gpu))
    cflo = c.dev.cflo.body
    assert mn == cflo.min() and mx == cflo.max()
    ref_mass = (c.dev.so0n.body[:, 0] * c.mesh.clvol.body).sum()
    assert abs(mass - ref_mass) <= 1e-12 * abs(ref_mass)
    assert c.dev.algorithmic_bytes() > 0


def test_md_bc_validation_and_normal_matrix(gpu, bridge):
    b = bridge
    c = Case(b, "hex3d")
    inlet, outlet, wall = c.groups
    interior = int(np.nonzero(c.mesh.fccls.body[:, 1] >= 0)[0][0])
    with pytest.raises(ValueError, match="is not a boundary face"):
        c.dev.add_slipwall([interior])
    with pytest.raises(ValueError, match="out of range"):
        c.dev.add_nonrefl([c.mesh.nface])
    with pytest.raises(ValueError, match="inlet value size"):
        c.dev.add_inlet(inlet, [1.0, 2.0])
    with pytest.raises(ValueError, match="gamma .* must be > 1"):
        c.dev.init_solution(1.0, 1.0, [0, 0, 0], 1.0)
    with pytest.raises(ValueError, match="velocity size"):
        c.dev.init_solution(1.4, 1.0, [0, 0], 1.0)
    with pytest.raises(IndexError):
        c.dev.get_normal_matrix(c.mesh.nface)
    for ifc in list(wall[:4]) + list(inlet[:2]):
        out = np.zeros((3, 3))
        nv = np.ascontiguousarray(c.mesh.fcnml.body[ifc])
        b.port_lib().orc_normal_matrix(3, b._p(nv), b._p(out))
        assert np.array_equal(np.array(c.dev.get_normal_matrix(int(ifc))), out)


def test_md_overlapping_bc_last_wins(gpu, bridge):
    """Boundaries apply in registration order; a face listed twice is overwritten by the later one."""
    b = bridge
    c = Case(b, "mixed2d")
    inlet, outlet, wall = c.groups
    extra = wall[: len(wall) // 2]
    c.chk.add_bc(b.BC_NONREFL, extra)
    c.dev.add_nonrefl(extra)
    c.chk.phase(b.MD_BC_SOLN); c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln(); c.dev.bc_dsoln()
    c.chk.march(5)
    c.dev.march(5)
    c.same("overlapping bc", STATE, _strict(gpu))
    c.dev.clear_bc(); c.chk.clear_bc()
    c.chk.march(2)
    c.dev.march(2)
    c.same("no bc", STATE, _strict(gpu))


def test_shared_reciprocal_division_is_ieee(gpu):
    """DivBy (scg_md_kernels.cuh) must return exactly a / b wherever its range test passes, and the range
    test must pass for everything in the normal range."""
    rng = np.random.RandomState(7)
    n = 1 << 20
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    # hard cases: mantissas of all ones / near powers of two, exact quotients, tiny and huge operands, zeros
    edge = np.array([1.0, 3.0, 1.0 / 3.0, np.nextafter(2.0, 0), np.nextafter(1.0, 2), 1e-300, 1e300, 4.9e-324,
                     2.2250738585072014e-308, 1.7976931348623157e308, 0.0, -0.0, 7.0, 1e-160, 1e160])
    ea, eb = np.meshgrid(edge, edge)
    a = np.concatenate([a, ea.ravel(), -ea.ravel()])
    b = np.concatenate([b, eb.ravel(), eb.ravel()])
    n = len(a)
    fast, ref = np.zeros(n), np.zeros(n)
    ok = np.zeros(n, dtype=np.int32)
    gpu.check(gpu.load().scg_selftest_div(gpu.ptr(a), gpu.ptr(b), gpu.ptr(fast), gpu.ptr(ref), ok.ctypes.data, n))
    m = ok == 1
    assert np.array_equal(fast[m].view(np.int64), ref[m].view(np.int64))
    with np.errstate(all="ignore"):
        host = a / b
    fin = np.isfinite(host) & (host != 0)
    assert np.array_equal(ref[fin].view(np.int64), host[fin].view(np.int64))   # device '/' is IEEE too
    assert m[: 1 << 20].mean() > 0.999                                          # the fast path is the common path
    # exact zeros over a sane denominator are answered by the fast path too (uniform flow regions are full of them)
    zs = (a == 0) & np.isfinite(b) & (np.abs(b) > 1e-250) & (np.abs(b) < 1e250)
    assert zs.sum() > 10 and m[zs].all()


def test_fast_sqrt_and_reciprocal_are_ieee(gpu):
    """sqrt_fast / rcp_fast (scg_md_kernels.cuh) must return exactly sqrt(x) / (1.0 / x) wherever their
    range tests pass, the tests must pass over the whole normal range the march uses, and the device
    operators themselves must agree with the host's IEEE results."""
    rng = np.random.RandomState(11)
    n0 = 1 << 20
    x = np.abs(rng.standard_normal(n0)) * 10.0 ** rng.uniform(-250, 250, n0)
    edge = np.array([1.0, 2.0, 3.0, 4.0, 1.0 / 3.0, np.nextafter(2.0, 0), np.nextafter(1.0, 2), np.nextafter(4.0, 0),
                     1e-200, 1e-100, 1e-300, 1e300, 4.9e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                     0.0, 7.0, 1e-160, 1e160, 0.25, 0.5, np.nextafter(0.5, 0)])
    x = np.concatenate([x, edge, -edge])
    n = len(x)
    fast, ref = np.zeros(2 * n), np.zeros(2 * n)
    ok = np.zeros(2 * n, dtype=np.int32)
    gpu.check(gpu.load().scg_selftest_rsq(gpu.ptr(x), gpu.ptr(fast), gpu.ptr(ref), ok.ctypes.data, n))
    m = ok == 1
    assert np.array_equal(fast[m].view(np.int64), ref[m].view(np.int64))
    with np.errstate(all="ignore"):
        host = np.concatenate([np.sqrt(x), 1.0 / x])
    fin = np.isfinite(host) & (host != 0) & (np.abs(host) > 1e-300)
    assert np.array_equal(ref[fin].view(np.int64), host[fin].view(np.int64))
    assert m[:n0].mean() > 0.999 and m[n:n + n0].mean() > 0.999


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d"])
def test_derived_fields_match_numpy(gpu, bridge, name):
    """Row f2: the device-side derived fields and reductions (EulerCore.derived / EulerCore.reduce: scg_md_derived,
    scg_md_reduce) against the numpy expressions of the reference's solvcon/multidim/euler/_field.py:56-148 applied to
    the downloaded state.  (The reference's own EulerField, unmodified, runs over the device core in
    tests/test_gpu_dropin.py and tests/_refsuite/multidim/euler/test_field.py.)"""
    c = Case(bridge, name, order="morton")
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    c.dev.march(10)
    nd = c.dev.ndim
    cons = c.dev.so0n.body.copy()
    ga = c.dev.gamma.body.copy()
    vel = cons[:, 1:1 + nd] / cons[:, 0][:, None]
    speed = np.sqrt((vel ** 2).sum(axis=1))
    pr = (ga - 1.0) * (cons[:, 1 + nd] - 0.5 * cons[:, 0] * speed ** 2)
    want = {"density": cons[:, 0], "velx": vel[:, 0], "vely": vel[:, 1], "speed": speed, "pressure": pr,
            "mach": speed / np.sqrt(ga * pr / cons[:, 0]), "total_energy": cons[:, 1 + nd], "cfl": c.dev.cflo.body}
    strict = _strict(gpu)
    for k, w in want.items():
        got = c.dev.derived(k)
        if strict:
            assert np.array_equal(got, w), k
        else:
            assert np.allclose(got, w, rtol=1e-12, atol=0), k
    with pytest.raises(ValueError, match="unknown field"):
        c.dev.derived("vorticity")
    lo, hi, mass = c.dev.reduce(c.mesh.clvol.ndarray)
    assert abs(mass - float(np.sum(cons[:, 0] * c.mesh.clvol.body))) <= 1e-13 * abs(mass)
    assert (lo, hi) == (float(c.dev.cflo.body.min()), float(c.dev.cflo.body.max()))


@pytest.mark.parametrize("name", ["step2d", "mix3d"])
def test_gmsh_mesh_marches_like_the_checker(gpu, bridge, name):
    """Row f3 end to end: a .msh file read by solvcon_b200.gmsh.Gmsh goes straight to the device EulerCore and marches
    bit for bit like the CPU checker on the tables of the same mesh (the reference's own use of its reader:
    tests/test_euler_multidim.py:191-224)."""
    import os
    from solvcon_b200 import meshgen
    from solvcon_b200.gmsh import Gmsh
    from solvcon_b200.multidim import EulerCore
    b = bridge
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".msh"), "rb") as f:
        mesh = Gmsh(f.read()).to_block()
    ndim = mesh.ndim
    nd = mesh.ndcrd.body
    lo, hi = nd.min(axis=0), nd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(ndim, mesh.fccnd.body, mesh.fccls.body, mesh.nface, lo, hi)
    h = (np.prod(hi - lo) / mesh.ncell) ** (1.0 / ndim)
    dt = 0.05 * h
    chk = b.PortEuler(mesh.tables(), ndim, dt)
    chk.phase(b.MD_PREPARE_CE)
    dev = EulerCore(mesh, dt, order="morton")
    vel = [3.0, 0.0, 0.0][:ndim]
    inval = [1.4] + vel + [1.0, 1.4]
    for s in (chk,):
        s.init_solution(1.4, 1.4, vel, 1.0)
        s.add_bc(b.BC_INLET, inlet, inval)
        s.add_bc(b.BC_NONREFL, outlet)
        s.add_bc(b.BC_SLIPWALL, wall)
        s.phase(b.MD_BC_SOLN)
        s.phase(b.MD_BC_DSOLN)
    dev.init_solution(1.4, 1.4, vel, 1.0)
    dev.add_inlet(inlet, inval)
    dev.add_nonrefl(outlet)
    dev.add_slipwall(wall)
    dev.bc_soln()
    dev.bc_dsoln()
    chk.march(20)
    dev.march(20)
    strict = _strict(gpu)
    for n in STATE:
        a, c = getattr(dev, n).ndarray, chk.arr(n)
        if strict:
            assert np.array_equal(a, c), n
        else:
            assert np.abs(a - c).max() / (np.abs(c).max() + 1e-300) < 1e-10, n
    assert np.isfinite(dev.so0n.ndarray).all()


def test_march_sees_host_edits_between_calls(gpu, bridge):
    """Host arrays stay the API-visible storage: edits of so0n / so1n / gamma and of GHOST rows of the
    other arrays between two march() calls must reach the device -- march() uploads exactly that set."""
    b = bridge
    c = Case(b, "mixed3d", order="morton")
    c.chk.phase(b.MD_BC_SOLN)
    c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    c.chk.march(3)
    c.dev.march(3)
    for s, get in ((c.chk, c.chk.arr), (c.dev, lambda n: getattr(c.dev, n).ndarray)):
        ng = c.dev.ngstcell
        get("so0n")[ng + 5:ng + 9, 0] *= 1.01          # body rows of the newest solution
        get("so1n")[ng + 2, 1, 0] += 1e-3
        get("gamma")[ng + 7] = 1.39
        get("so0c")[0] = 0.123                          # a ghost row of an array march() overwrites in the body
        get("cflc")[1] = 0.5
        get("so0t")[ng + 3] = 9.0                       # dead: recomputed by calc_solt
    c.chk.march(3)
    c.dev.march(3)
    c.same("march after host edits", STATE, _strict(gpu))


def test_only_observable_bytes_travel(gpu, bridge):
    """The coherence protocol of EulerCore (solvcon_b200/multidim.py): an array crosses the bus only when somebody can
    observe it -- uploaded iff the caller may have written it (handed out since the last upload, or still held),
    downloaded at once iff the caller holds it or a view of it, else at the next access.  Results never depend on it."""
    b = bridge
    c = Case(b, "mixed3d", order="morton")
    dev, chk = c.dev, c.chk
    chk.phase(b.MD_BC_SOLN)
    chk.phase(b.MD_BC_DSOLN)
    dev.bc_soln()
    dev.bc_dsoln()
    T = dev.ncell + dev.ngstcell
    # 1. nothing held, nothing handed out: march() moves no bytes at all
    dev.sync_to_host()
    dev._exposed.clear()
    h0, d0 = dev.h2d_bytes, dev.d2h_bytes
    chk.march(2)
    dev.march(2)
    assert (dev.h2d_bytes, dev.d2h_bytes) == (h0, d0)
    # 2. the next access refreshes exactly the array asked for
    rho = dev.so0n.ndarray
    assert dev.d2h_bytes - d0 == rho.nbytes == T * 5 * 8
    assert np.array_equal(rho, chk.arr("so0n"))
    # 3. a HELD view is uploaded before and refreshed after every march, like a view of the reference's live memory
    d1, h1 = dev.d2h_bytes, dev.h2d_bytes
    rho[dev.ngstcell + 3, 0] *= 1.001
    chk.arr("so0n")[dev.ngstcell + 3, 0] *= 1.001
    chk.march(2)
    dev.march(2)
    assert dev.h2d_bytes - h1 == rho.nbytes and dev.d2h_bytes - d1 == rho.nbytes
    assert np.array_equal(rho, chk.arr("so0n"))
    del rho
    # 4. arrays nobody looked at are still right when finally read
    c.same("lazy refresh", STATE, _strict(gpu))
    # 5. resident mode: no implicit transfer in either direction, property access still returns current data
    dev.resident(True)
    h2, d2 = dev.h2d_bytes, dev.d2h_bytes
    chk.march(1)
    dev.march(1)
    assert (dev.h2d_bytes, dev.d2h_bytes) == (h2, d2)
    assert np.array_equal(dev.cflo.ndarray, chk.arr("cflo"))
    dev.resident(False)
    c.same("after resident", STATE, _strict(gpu))


@pytest.mark.parametrize("name", ["one_quad", "two_tri", "one_hex", "six_tet", "two_prism", "six_pyramid"])
@pytest.mark.parametrize("fused", [3, 0])
def test_md_smallest_meshes(gpu, bridge, name, fused):
    """Edge sizes: one box of every cell type -- all neighbours ghosts, a handful of cells per type group."""
    b = bridge
    c = Case(b, name, order="morton")
    c.dev.set_fused(fused)
    strict = _strict(gpu)
    c.same("prepare_ce", GEOM, strict)
    c.chk.phase(b.MD_BC_SOLN)
    c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    c.chk.march(6)
    c.dev.march(6)
    c.same(f"march(6) fused={fused}", STATE, strict)


@pytest.mark.parametrize("name,order", [("mixed3d_150k", "rcm"), ("mixed3d_150k", "morton"), ("mixed2d_200k", "rcm")])
def test_md_many_waves_matches(gpu, bridge, name, order):
    """Meshes of 1.5e5 - 2e5 cells: several waves of CTAs per kernel, ragged last CTAs in every type group, both
    production orderings -- still bit-equal to the checker after 3 steps."""
    b = bridge
    c = Case(b, name, order=order)
    assert c.dev.ncell > 140_000
    c.chk.phase(b.MD_BC_SOLN)
    c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    c.chk.march(3)
    c.dev.march(3)
    c.same(f"march(3) {name} {order}", STATE, _strict(gpu))
    vol = c.mesh.clvol.body
    assert np.isfinite(c.dev.so0n.ndarray).all() and (c.dev.so0n.body[:, 0] > 0).all()
    assert abs((c.dev.so0n.body[:, 0] * vol).sum() - (c.chk.body("so0n")[:, 0] * vol).sum()) <= 1e-13 * vol.sum()


def test_ce_geometry_hand_values(gpu):
    """The reference's own CE-geometry known answers (tests/test_euler_multidim.py:11-144), written against the
    mirrored classes exactly as the reference writes them: a fan of 3 triangles and a unit-square quadrilateral."""
    import solvcon_b200 as solvcon
    mh = solvcon.StaticMesh(ndim=2, nnode=4, nface=0, ncell=3)
    mh.ndcrd[:, :] = [(0, 0), (-1, -1), (1, -1), (0, 1)]
    mh.cltpn.fill(solvcon.StaticMesh.TRIANGLE)
    mh.clnds[:, :4] = (3, 0, 1, 2), (3, 0, 2, 3), (3, 0, 3, 1)
    mh.build_interior(do_metric=True)
    mh.build_boundary()
    mh.build_ghost()
    ec = solvcon.EulerCore(mesh=mh, time_increment=0.01)
    assert (ec.ndim, ec.ncell, ec.ngstcell) == (2, 3, 3)
    total = ec.ngstcell + ec.ncell
    CLMFC, FCMND = solvcon.StaticMesh.CLMFC, solvcon.StaticMesh.FCMND
    assert ec.cevol.shape == (total, CLMFC + 1) and ec.cecnd.shape == (total, (CLMFC + 1) * 2)
    assert ec.sfcnd.shape == (3, CLMFC * FCMND, 2) and ec.sfnml.shape == (3, CLMFC * FCMND, 2)
    for icl in range(3):
        assert ec.cevol[icl, 0] > 0.0
        nfc = mh.clfcs[icl, 0]
        assert abs(sum(ec.cevol[icl, ifl] for ifl in range(1, nfc + 1)) - ec.cevol[icl, 0]) < 1e-12
    assert abs(ec.cevol[0, 1] - 0.5) < 1e-10 and abs(ec.cevol[0, 2] - 2.0 / 3.0) < 1e-10 and abs(ec.cevol[0, 3] - 0.5) < 1e-10
    assert abs(ec.cevol[0, 0] - (0.5 + 2.0 / 3.0 + 0.5)) < 1e-10
    before = ec.cevol[0, 0]
    ec.prepare_ce()
    assert ec.cevol[0, 0] == before

    mq = solvcon.StaticMesh(ndim=2, nnode=4, nface=0, ncell=1)
    mq.ndcrd[:, :] = [(0, 0), (1, 0), (1, 1), (0, 1)]
    mq.cltpn.fill(solvcon.StaticMesh.QUADRILATERAL)
    mq.clnds[:, :5] = [(4, 0, 1, 2, 3)]
    mq.build_interior(do_metric=True)
    mq.build_boundary()
    mq.build_ghost()
    eq = solvcon.EulerCore(mesh=mq, time_increment=0.01)
    assert (eq.ndim, eq.ncell, eq.ngstcell) == (2, 1, 4)
    for ifl in range(1, 5):
        assert abs(eq.cevol[0, ifl] - 0.5) < 1e-10
    assert abs(eq.cevol[0, 0] - 2.0) < 1e-10
    mags = [np.linalg.norm(eq.sfnml.ndarray[0, ifc * FCMND + ind]) for ifc in range(4) for ind in range(FCMND)]
    assert any(m > 1e-14 for m in mags)
    for m in mags:
        if m > 1e-14:
            assert abs(m - 0.5 * np.sqrt(2.0)) < 1e-10
    assert np.any(eq.sfcnd.ndarray[0] != 0)


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d", "pyramid3d"])
def test_reference_invariants(gpu, name):
    """Invariants the reference's tests/test_multidim.py pins without any second implementation: init_solution
    validation (:505-515), calc_cfl on a uniform field (:517-541), the negative-pressure clamp (:543-558), update()
    swapping the buffers (:560-577), free-stream preservation of calc_soln (:657-675)."""
    from cases import md_mesh_case
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import EulerCore
    nd, ndcrd, cltpn, clnds = md_mesh_case(name)
    mesh = make_mesh(nd, ndcrd, cltpn, clnds)
    GAMMA, RHO, PRES = 1.4, 1.2, 0.9
    v = [0.3, -0.2, 0.1][:nd]
    ec = EulerCore(mesh, 1.e-3)
    good = dict(gamma=GAMMA, rho=RHO, v=v, p=PRES)
    for bad in (dict(v=[0.1] * (nd - 1)), dict(gamma=1.0), dict(rho=0.0), dict(p=-1.0)):
        with pytest.raises(ValueError):
            ec.init_solution(**dict(good, **bad))
    # ---- calc_cfl, uniform field
    ec.init_solution(**good)
    ec.calc_cfl()
    hdt = ec.time_increment / 2.0
    vsq = sum(c * c for c in v)
    wspd = np.sqrt(GAMMA * PRES / RHO) + np.sqrt(vsq)
    energy0 = PRES / (GAMMA - 1.0) + 0.5 * RHO * vsq
    ce = ec.cecnd.body
    nfc = mesh.clfcs.body[:, 0]
    for icl in range(0, ec.ncell, max(1, ec.ncell // 50)):
        dist = min(np.sqrt(((ce[icl, f * nd:(f + 1) * nd] - ce[icl, :nd]) ** 2).sum()) for f in range(1, nfc[icl] + 1))
        assert abs(ec.cflo[icl] - hdt * wspd / dist) < 1e-7 * ec.cflo[icl] + 1e-12
        assert abs(ec.cflc[icl] - ec.cflo[icl]) < 1e-7
        assert abs(ec.so0n[icl, nd + 1] - energy0) < 1e-7
    # ---- negative pressure: clamped CFL forced to 1, energy rebuilt from the kinetic part
    ec.init_solution(**good)
    ec.so0n.body[:, nd + 1] = 0.0
    ec.calc_cfl()
    ke = sum((RHO * c) ** 2 for c in v) / (2.0 * RHO)
    assert np.allclose(ec.cflc.body, 1.0, rtol=0, atol=1e-7) and np.allclose(ec.so0n.body[:, nd + 1], ke, rtol=0, atol=1e-7)
    # ---- update() swaps
    ec.init_solution(**good)
    ec.so1n.ndarray[...] = np.arange(ec.so1n.ndarray.size, dtype="float64").reshape(ec.so1n.ndarray.shape)
    before = {n: getattr(ec, n).ndarray.copy() for n in ("so0n", "so0c", "so1n", "so1c")}
    ec.update()
    assert np.array_equal(ec.so0c.ndarray, before["so0n"]) and np.array_equal(ec.so0n.ndarray, before["so0c"])
    assert np.array_equal(ec.so1c.ndarray, before["so1n"]) and np.array_equal(ec.so1n.ndarray, before["so1c"])
    # ---- free stream: a uniform state with zero gradient is preserved by the flux integral
    ec2 = EulerCore(mesh, 1.e-3)
    row = [RHO] + [RHO * c for c in v] + [energy0]
    ec2.gamma.ndarray[...] = GAMMA
    ec2.so0c.ndarray[...] = row
    ec2.so0t.ndarray[...] = 0.0
    ec2.so1c.ndarray[...] = 0.0
    ec2.calc_solt()
    ec2.calc_soln()
    assert np.allclose(ec2.so0n.body, row, rtol=0, atol=1e-10)


@pytest.mark.parametrize("dims", [2, 3])
def test_channel_march_stays_uniform(gpu, dims):
    """tests/test_multidim.py:1028-1083 of the reference: a walled channel with a supersonic inlet, slip walls and a
    non-reflective outflow keeps the aligned uniform stream (density within 1e-6 of the inflow after 5 steps), and the
    inlet ghosts hold the prescribed state -- written like the reference writes it (bndfcs, fccnd, fccls, so0n[ghost])."""
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import EulerCore
    GAMMA, RHO, PRES, VX = 1.4, 1.0, 1.0, 2.0
    if dims == 2:
        lx = 4.0
        mh = make_mesh(2, *meshgen.mixed_2d(4, 2, lx=lx, ly=2.0, tri_every=0))
        vel, dt = [VX, 0.0], 0.04
    else:
        lx = 4.0
        mh = make_mesh(3, *meshgen.mixed_3d(4, 2, 2, lx=lx, ly=2.0, lz=2.0, kinds="h"))
        vel, dt = [VX, 0.0, 0.0], 0.04
    left, right, walls = [], [], []
    for ifc in mh.bndfcs.ndarray[:, 0]:
        cx = mh.fccnd[int(ifc), 0]
        (left if abs(cx) < 1e-9 else right if abs(cx - lx) < 1e-9 else walls).append(int(ifc))
    assert left and right and walls
    ec = EulerCore(mh, dt)
    ec.init_solution(gamma=GAMMA, rho=RHO, v=vel, p=PRES)
    ec.add_inlet(left, value=[RHO] + vel + [PRES, GAMMA])
    ec.add_nonrefl(right)
    ec.add_slipwall(walls)
    ec.bc_soln()
    ec.bc_dsoln()
    ec.march(steps=5)
    so0n = ec.so0n.ndarray
    assert np.all(np.isfinite(so0n)) and np.abs(so0n).max() < 1e6
    for icl in range(ec.ncell):
        assert ec.so0n[icl, 0] > 0.0 and abs(ec.so0n[icl, 0] - RHO) < 1e-6
    for ifc in left:
        assert abs(ec.so0n[int(mh.fccls[ifc, 1]), 0] - RHO) < 1e-7


@pytest.mark.parametrize("name", ["mixed2d", "step2d", "mixed3d", "pyramid3d"])
@pytest.mark.parametrize("order", ["none", "morton", "rcm"])
def test_device_layout_equals_host_layout(gpu, bridge, monkeypatch, name, order):
    """The layout stage (ordering keys incl. the level-by-level reverse Cuthill-McKee, the sort into type / class
    groups, perm) runs on the device by default; the host implementation is its checker: same permutation, with and
    without an ownership array (partitions: interface cells first, halo copies last inside a type group)."""
    from solvcon_b200.mesh import TableMesh
    from solvcon_b200.multidim import DeviceMesh
    if not bridge.have_ref():
        pytest.skip("needs oracle/_ref (mesh tables come from the reference's StaticMesh)")
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    mesh = TableMesh(ndim, bridge.make_ref_mesh(ndim, ndcrd, cltpn, clnds).tables())
    rng = np.random.default_rng(5)
    owned = rng.integers(0, 3, size=len(cltpn)).astype(np.int8)
    for own in (None, owned):
        monkeypatch.delenv("SCG_MESH_HOST_LAYOUT", raising=False)
        dev = DeviceMesh(mesh, order=order, owned=own).perm
        monkeypatch.setenv("SCG_MESH_HOST_LAYOUT", "1")
        host = DeviceMesh(mesh, order=order, owned=own).perm
        monkeypatch.delenv("SCG_MESH_HOST_LAYOUT", raising=False)
        assert np.array_equal(np.sort(dev), np.arange(len(cltpn)))
        assert np.array_equal(dev, host), f"{name} {order}: device and host layouts order the cells differently"


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d", "pyramid3d"])
def test_md_march_same_bit_patterns(gpu, bridge, name):
    """Stricter than array_equal (which calls -0.0 and +0.0 equal): the arrays a caller can read after march() carry the
    checker's bit patterns, signs of zeros included.  The default kernels drop two leading `0.0 +` of ordered sums
    (SCG_SKIP_ZERO_ADD) on the argument that the sign of those zeros cannot reach a result; the bump case keeps whole
    regions of exactly zero transverse momentum and zero gradients, which is where it would show."""
    b = bridge
    if not _strict(gpu):
        pytest.skip("bit patterns are a property of the strict build")
    c = Case(b, name, order="rcm")
    c.chk.phase(b.MD_BC_SOLN)
    c.chk.phase(b.MD_BC_DSOLN)
    c.dev.bc_soln()
    c.dev.bc_dsoln()
    c.chk.march(12)
    c.dev.march(12)
    for n in ("so0n", "so1n", "cflo", "cflc"):
        a = np.ascontiguousarray(getattr(c.dev, n).ndarray).view(np.int64)
        w = np.ascontiguousarray(c.chk.arr(n)).view(np.int64)
        assert np.array_equal(a, w), f"{name}: {n} differs in bit pattern at {np.argwhere(a != w)[:4].tolist()}"


def test_device_layout_rcm_on_a_chain_of_cells(gpu, monkeypatch):
    """A strip of cells is a graph with as many BFS levels as cells; the level-by-level device RCM hands such graphs back
    to the sequential host code after 16384 levels.  Same permutation as the host layout."""
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import DeviceMesh
    ndcrd, cltpn, clnds = meshgen.mixed_2d(18000, 1, lx=18000.0, ly=1.0)
    mesh = make_mesh(2, ndcrd, cltpn, clnds, device=False)
    monkeypatch.delenv("SCG_MESH_HOST_LAYOUT", raising=False)
    dev = DeviceMesh(mesh, order="rcm").perm
    monkeypatch.setenv("SCG_MESH_HOST_LAYOUT", "1")
    host = DeviceMesh(mesh, order="rcm").perm
    monkeypatch.delenv("SCG_MESH_HOST_LAYOUT", raising=False)
    assert len(dev) > 18000 and np.array_equal(dev, host)
