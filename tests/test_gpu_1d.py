"""GPU parity of the 1D device path against the CPU checkers (bit-exact in the strict build).

Every call goes through the C-ABI of include/scg.h (via solvcon_b200's ctypes layer)."""
import numpy as np
import pytest

from cases import scalar_ic, smooth_e1d_ic, sod_ic

pytestmark = pytest.mark.gpu


def _checker(bridge):
    """Prefer the compiled reference itself; fall back to the pinned C port."""
    return bridge.have_ref()


def _assert_same(a, b, name, strict):
    if strict:
        assert np.array_equal(a, b), f"{name}: max abs diff {np.abs(a - b).max()}"
    else:
        # FMA speed variant (not the product default): documented distance from the reference, DESIGN.md section 2
        scale = np.abs(b).max() + 1e-300
        assert np.abs(a - b).max() / scale < 1e-9, name


@pytest.mark.parametrize("alpha", [1, 2])
def test_e1d_sod_config1(gpu, bridge, alpha):
    """BASELINE config 1: Sod tube, 10^4 grid points, 1000 steps, exact-match check."""
    from solvcon_b200.onedim import Euler1DCore
    n, dt, steps = 10001, 1.6e-4, 1000
    chk = bridge.RefEuler1D(n, dt) if bridge.have_ref() else bridge.PortEuler1D(n, dt)
    dev = Euler1DCore(n, dt)
    sod_ic(chk)
    sod_ic(dev)
    chk.march(alpha, steps)
    getattr(dev, f"march_alpha{alpha}")(steps)
    strict = bool(gpu.load().scg_build_is_strict())
    for name in ("so0", "so1", "cfl"):
        _assert_same(getattr(dev, name), getattr(chk, name), name, strict)
    # conserved totals (even plane, uniform spacing) equal to round-off
    tot_d = dev.so0[2:-2:2].sum(axis=0)
    tot_c = chk.so0[2:-2:2].sum(axis=0)
    assert np.allclose(tot_d, tot_c, rtol=1e-14, atol=0)


def test_e1d_phases_match(gpu, bridge):
    """Fine-grained phase calls, like tests/test_onedim_euler.py:53-71 of the reference."""
    from solvcon_b200.onedim import Euler1DCore
    n, dt = 401, 2.0e-3
    chk = bridge.PortEuler1D(n, dt)
    dev = Euler1DCore(n, dt)
    smooth_e1d_ic(chk)
    smooth_e1d_ic(dev)
    strict = bool(gpu.load().scg_build_is_strict())
    for _ in range(3):
        for fn, ph, odd in (("march_half_so0", 0, False), ("update_cfl", 1, False),
                            ("march_half_so1_alpha2", 2, False), ("treat_boundary_so0", 3, False),
                            ("treat_boundary_so1", 4, False), ("march_half_so0", 0, True),
                            ("update_cfl", 1, True), ("march_half_so1_alpha2", 2, True)):
            chk.phase(ph, odd, 2)
            if fn.startswith("treat"):
                getattr(dev, fn)()
            else:
                getattr(dev, fn)(odd)
            for name in ("so0", "so1", "cfl"):
                _assert_same(getattr(dev, name), getattr(chk, name), f"{name} after {fn}({odd})", strict)


def test_e1d_fine_equals_march(gpu):
    """march_alpha2(1) == the eight fine-grained calls, bit for bit (reference test_onedim_euler.py:53-71)."""
    from solvcon_b200.onedim import Euler1DCore
    n, dt = 257, 1.0e-3
    a, b = Euler1DCore(n, dt), Euler1DCore(n, dt)
    smooth_e1d_ic(a)
    smooth_e1d_ic(b)
    for _ in range(10):
        a.march_alpha2(1)
        b.march_half1_alpha2()
        b.treat_boundary_so0()
        b.treat_boundary_so1()
        b.march_half2_alpha2()
        assert np.array_equal(a.so0, b.so0) and np.array_equal(a.so1, b.so1) and np.array_equal(a.cfl, b.cfl)


@pytest.mark.parametrize("n", [5, 7, 63, 65, 67, 1025])
def test_e1d_small_and_ragged_sizes(gpu, bridge, n):
    """Edge sizes around the 31-target warp tiling, down to the minimum ncoord."""
    from solvcon_b200.onedim import Euler1DCore
    dt = 0.2 / n
    chk = bridge.PortEuler1D(n, dt)
    dev = Euler1DCore(n, dt)
    smooth_e1d_ic(chk)
    smooth_e1d_ic(dev)
    chk.march(2, 7)
    dev.march_alpha2(7)
    strict = bool(gpu.load().scg_build_is_strict())
    for name in ("so0", "so1", "cfl"):
        _assert_same(getattr(dev, name), getattr(chk, name), name, strict)


def test_e1d_even_ncoord_rejected(gpu):
    from solvcon_b200.onedim import Euler1DCore
    with pytest.raises(ValueError, match="ncoord cannot be even"):
        Euler1DCore(10, 0.1)


@pytest.mark.parametrize("eq", [0, 1])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_scalar_march_matches(gpu, bridge, eq, alpha):
    """Config 2 at parity size: ncelm = 10^6, dt = dx/2, 20 steps, vs the reference."""
    from solvcon_b200 import spacetime as st
    n = 1_000_000 if alpha == 2 else 20_001
    steps = 20
    xmax = 8 * np.pi
    dt = 0.5 * xmax / n
    chk = (bridge.RefScalar1D if bridge.have_ref() else bridge.PortScalar1D)(eq, 0.0, xmax, n, dt)
    grid = st.Grid(0.0, xmax, n)
    assert np.array_equal(grid.xcoord.ndarray, chk.x)
    dev = (st.LinearScalarSolver if eq == 0 else st.InviscidBurgersSolver)(grid, dt)
    scalar_ic(chk.x, chk.so0, chk.so1, chk.cfl)
    scalar_ic(chk.x, dev.so0.ndarray[:, 0], dev.so1.ndarray[:, 0], dev.cfl.ndarray)
    chk.march(alpha, steps)
    getattr(dev, f"march_alpha{alpha}")(steps)
    strict = bool(gpu.load().scg_build_is_strict())
    _assert_same(dev.so0.ndarray[:, 0], chk.so0, "so0", strict)
    _assert_same(dev.so1.ndarray[:, 0], chk.so1, "so1", strict)
    _assert_same(dev.cfl.ndarray, chk.cfl, "cfl", strict)


def test_scalar_known_answer_cfl1(gpu):
    """Reference known answer (tests/test_spacetime_linear_scalar.py:15-38,106-113): 8 cells over one
    period at CFL = 1 reproduce sin(x) after 10 periods to 1e-14, and cfl == 1."""
    from solvcon_b200 import spacetime as st
    resolution, cycle = 8, 10
    xcrd = np.arange(resolution + 1) / resolution
    xcrd *= 2 * np.pi
    grid = st.Grid(xcrd)
    dx = (grid.xmax - grid.xmin) / grid.ncelm
    time_stop = 2 * np.pi
    nstep = int(np.ceil(time_stop / (dx * 1.0)))
    dt = time_stop / nstep
    svr = st.LinearScalarSolver(grid=grid, time_increment=dt)
    svr.set_so0(0, np.sin(xcrd))
    svr.set_so1(0, np.cos(xcrd))
    svr.setup_march()
    assert svr.xctr().tolist() == xcrd.tolist()
    assert svr.xctr().tolist() == [e.xctr for e in svr.selms(odd_plane=False)]
    svr.march_alpha2(nstep * cycle)
    np.testing.assert_allclose(svr.get_so0(0), np.sin(xcrd), rtol=0, atol=1e-14)
    np.testing.assert_allclose(svr.get_cfl(), np.ones(grid.nselm), rtol=0, atol=1e-14)


@pytest.mark.parametrize("eq", [0, 1])
def test_scalar_fine_interface_bit_equal(gpu, eq):
    """Fine-grained phases == march_alpha2, bit for bit (tests/test_spacetime_linear_scalar.py:115-136)."""
    from solvcon_b200 import spacetime as st
    xloc = np.cumsum(0.1 + 0.05 * np.sin(np.arange(130)))
    cls = st.LinearScalarSolver if eq == 0 else st.InviscidBurgersSolver
    a, b = cls(st.Grid(xloc=xloc), 0.02), cls(st.Grid(xloc=xloc), 0.02)
    for s in (a, b):
        scalar_ic(s.grid.xcoord.ndarray, s.so0.ndarray[:, 0], s.so1.ndarray[:, 0])
    for _ in range(6):
        a.march_alpha2(1)
        b.march_half_so0(odd_plane=False)
        b.treat_boundary_so0()
        b.update_cfl(odd_plane=True)
        b.march_half_so1_alpha2(odd_plane=False)
        b.treat_boundary_so1()
        b.march_half_so0(odd_plane=True)
        b.update_cfl(odd_plane=False)
        b.march_half_so1_alpha2(odd_plane=True)
        strict = bool(gpu.load().scg_build_is_strict())   # the FMA speed variant contracts the fused kernel differently
        for name in ("so0", "so1", "cfl"):
            _assert_same(getattr(a, name).ndarray, getattr(b, name).ndarray, name, strict)


@pytest.mark.parametrize("ncelm", [1, 2, 3, 255, 256, 257])
def test_scalar_tiny_and_ragged(gpu, bridge, ncelm):
    from solvcon_b200 import spacetime as st
    dt = 0.4 / ncelm
    for eq in (0, 1):
        chk = bridge.PortScalar1D(eq, 0.0, 1.0, ncelm, dt)
        dev = (st.LinearScalarSolver if eq == 0 else st.InviscidBurgersSolver)(st.Grid(0.0, 1.0, ncelm), dt)
        scalar_ic(chk.x * 6, chk.so0, chk.so1, chk.cfl)
        dev.so0.ndarray[:, 0] = chk.so0
        dev.so1.ndarray[:, 0] = chk.so1
        chk.march(2, 9)
        dev.march_alpha2(9)
        strict = bool(gpu.load().scg_build_is_strict())
        _assert_same(dev.so0.ndarray[:, 0], chk.so0, "so0", strict)
        _assert_same(dev.so1.ndarray[:, 0], chk.so1, "so1", strict)
        _assert_same(dev.cfl.ndarray, chk.cfl, "cfl", strict)


def test_scalar_reduce_and_resident(gpu):
    from solvcon_b200 import spacetime as st
    n = 4096
    grid = st.Grid(0.0, 8 * np.pi, n)
    svr = st.InviscidBurgersSolver(grid, 0.5 * 8 * np.pi / n)
    scalar_ic(grid.xcoord.ndarray, svr.so0.ndarray[:, 0], svr.so1.ndarray[:, 0])
    svr.resident(True)
    svr.march_alpha2(5)
    mn, mx, sm = svr.reduce()
    svr.sync_to_host()
    cfl = svr.cfl.ndarray
    ref_pts = np.r_[cfl[2:2 * n + 3:2], cfl[3:2 * n + 2:2]]
    assert mn == ref_pts.min() and mx == ref_pts.max()
    assert abs(sm - svr.so0.ndarray[2:2 * n + 3:2, 0].sum()) < 1e-9


@pytest.mark.parametrize("eq", [0, 1])
@pytest.mark.parametrize("nparts", [1, 2, 3, 5])
def test_scalar_slabs_bit_identical(gpu, bridge, eq, nparts):
    """BASELINE config 2 at parity size: the slab-decomposed march (ring halo exchange of the odd-plane end points,
    solvcon_b200/decomp1d.py) equals the reference's undecomposed periodic march in every slot of every slab."""
    from solvcon_b200 import spacetime as st
    from solvcon_b200.decomp1d import LocalRing, ScalarSlab, slab_ranges
    n, steps, alpha = 20_003, 40, 2
    xmax = 8 * np.pi
    dt = 0.5 * xmax / n
    chk = (bridge.RefScalar1D if bridge.have_ref() else bridge.PortScalar1D)(eq, 0.0, xmax, n, dt)
    xg = st.Grid(0.0, xmax, n).xcoord.ndarray
    assert np.array_equal(xg, chk.x)
    scalar_ic(chk.x, chk.so0, chk.so1, chk.cfl)
    so0, so1, cfl = chk.so0.copy(), chk.so1.copy(), chk.cfl.copy()
    slabs = [ScalarSlab(eq, xg, c0, c1, dt) for c0, c1 in slab_ranges(n, nparts)]
    assert slabs[0].c0 == 0 and slabs[-1].c1 == n and all(a.c1 == b.c0 for a, b in zip(slabs, slabs[1:]))
    for s in slabs:
        s.load_global(so0, so1, cfl)
    chk.march(alpha, steps)
    LocalRing(slabs).march(alpha, steps)
    strict = bool(gpu.load().scg_build_is_strict())
    for r, s in enumerate(slabs):
        a0, a1, ac = s.fetch()
        lo, hi = 2 * s.c0 + 1, 2 * s.c1 + 4       # every slot but the two outermost ghost coordinates
        _assert_same(a0[1:-1], chk.so0[lo:hi], f"slab {r} so0", strict)
        _assert_same(a1[1:-1], chk.so1[lo:hi], f"slab {r} so1", strict)
        _assert_same(ac[1:-1], chk.cfl[lo:hi], f"slab {r} cfl", strict)
    assert np.abs(chk.so1).max() > 0.1


def test_celm_selm_cursors_known_answers(gpu):
    """tests/test_spacetime_solution.py:75-152 of the reference: cursor strings (with the solver kind's prefix,
    spacetime/io.cpp:74-86) and the bounds messages."""
    from solvcon_b200 import spacetime as st
    sol = st.LinearScalarSolver(st.Grid(xmin=0.0, xmax=10.0, nelm=10), 0.2)
    with pytest.raises(IndexError, match=r"Field::celm_at\(ielm=-1, odd_plane=0\): xindex = 1 outside the interval \[2, 23\)"):
        sol.celm(-1, odd_plane=False)
    assert str(sol.celm(-1, odd_plane=True)) == "LinearScalarCelm(odd, index=-1, x=0, xneg=-0.5, xpos=0.5)"
    assert str(sol.celm(0)) == "LinearScalarCelm(even, index=0, x=0.5, xneg=0, xpos=1)"
    assert str(sol.celm(ielm=0, odd_plane=True)) == "LinearScalarCelm(odd, index=0, x=1, xneg=0.5, xpos=1.5)"
    assert str(sol.celm(9, odd_plane=False)) == "LinearScalarCelm(even, index=9, x=9.5, xneg=9, xpos=10)"
    assert str(sol.celm(9, odd_plane=True)) == "LinearScalarCelm(odd, index=9, x=10, xneg=9.5, xpos=10.5)"
    with pytest.raises(IndexError, match=r"Field::celm_at\(ielm=10, odd_plane=0\): xindex = 23 outside the interval \[2, 23\)"):
        sol.celm(10)
    with pytest.raises(IndexError, match=r"Field::selm_at\(ielm=-1, odd_plane=0\): xindex = 0 outside the interval \[1, 24\)"):
        sol.selm(-1, odd_plane=False)
    assert str(sol.selm(-1, odd_plane=True)) == "LinearScalarSelm(odd, index=-1, x=-0.5, xneg=-1, xpos=0)"
    assert str(sol.selm(0, odd_plane=False)) == "LinearScalarSelm(even, index=0, x=0, xneg=-0.5, xpos=0.5)"
    assert str(sol.selm(0, odd_plane=True)) == "LinearScalarSelm(odd, index=0, x=0.5, xneg=0, xpos=1)"
    assert str(sol.selm(10, odd_plane=False)) == "LinearScalarSelm(even, index=10, x=10, xneg=9.5, xpos=10.5)"
    assert str(sol.selm(10, odd_plane=True)) == "LinearScalarSelm(odd, index=10, x=10.5, xneg=10, xpos=11)"
    with pytest.raises(IndexError, match=r"Field::selm_at\(ielm=11, odd_plane=0\): xindex = 24 outside the interval \[1, 24\)"):
        sol.selm(11)
    nx = (sol.grid.ncelm + sol.grid.BOUND_COUNT) * 2 + 1
    assert sol.so0.shape == (nx, 1) and sol.so1.shape == (nx, 1) and sol.so0.ndarray.dtype == np.float64
    assert len(list(sol.celms())) == sol.grid.ncelm and len(list(sol.selms())) == sol.grid.nselm
    assert len(list(sol.celms(odd_plane=True))) == sol.grid.ncelm - 1
