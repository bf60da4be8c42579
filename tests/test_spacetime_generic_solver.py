"""The generic ``spacetime.Solver`` with a ``Kernel`` of Python callables (host-only by construction, like in the
reference): the reference's own tests/test_spacetime_solution.py:181-360 restated -- a linear-wave kernel written in
Python reproduces the sine after 10 periods to 1e-14 with cfl == 1, ``kernel.reset()`` gives the zero solver, the fine
interface equals ``march_alpha2`` bit for bit -- plus parity of the callback path with the C port.  CPU only."""
import numpy as np
import pytest


def _build(resolution=8):
    from solvcon_b200 import spacetime as libst
    xcrd = np.arange(resolution + 1) / resolution
    xcrd *= 2 * np.pi
    grid = libst.Grid(xcrd)
    dx = (grid.xmax - grid.xmin) / grid.ncelm
    time_stop = 2 * np.pi
    nstep = int(np.ceil(time_stop / (dx * 1.0)))
    dt = time_stop / nstep
    svr = libst.Solver(grid=grid, time_increment=dt, nvar=1)

    def xn(se, iv):
        displacement = 0.5 * (se.x + se.xneg) - se.xctr
        return se.dxneg * (se.get_so0(iv) + displacement * se.get_so1(iv))

    def xp(se, iv):
        displacement = 0.5 * (se.x + se.xpos) - se.xctr
        return se.dxpos * (se.get_so0(iv) + displacement * se.get_so1(iv))

    def tn(se, iv):
        ret = se.get_so0(iv)
        ret += (se.x - se.xctr) * se.get_so1(iv)
        ret += se.qdt * se.get_so1(iv)
        return se.hdt * ret

    def tp(se, iv):
        ret = se.get_so0(iv)
        ret += (se.x - se.xctr) * se.get_so1(iv)
        ret -= se.qdt * se.get_so1(iv)
        return se.hdt * ret

    def so0p(se, iv):
        ret = se.get_so0(iv)
        ret += (se.x - se.xctr) * se.get_so1(iv)
        ret -= se.hdt * se.get_so1(iv)
        return ret

    def cfl(se):
        se.set_cfl(se.hdt / min(se.dxneg, se.dxpos))

    k = svr.kernel
    k.xn_calc, k.xp_calc, k.tn_calc, k.tp_calc, k.so0p_calc, k.cfl_updater = xn, xp, tn, tp, so0p, cfl
    svr.set_so0(0, np.sin(xcrd))
    svr.set_so1(0, np.cos(xcrd))
    svr.setup_march()
    return nstep, xcrd, svr


def test_shapes_getters_and_kernel_hooks():
    nstep, xcrd, svr = _build()
    assert svr.nvar == 1
    assert svr.xctr().tolist() == xcrd.tolist() == [e.xctr for e in svr.selms(odd_plane=False)]
    assert svr.get_so0(0).ndarray.tolist() == np.sin(xcrd).tolist()
    assert svr.get_so1(0).ndarray.tolist() == np.cos(xcrd).tolist()
    assert [e.get_so0(0) for e in svr.selms()] == svr.get_so0(0).ndarray.tolist()
    for bad in (lambda: svr.get_so0(1), lambda: svr.get_so1(1, odd_plane=True)):
        with pytest.raises(IndexError, match="out of nvar range"):
            bad()
    svr.kernel.reset()
    se = svr.selm(0)
    assert se.xn(0) == 0
    svr.kernel.xn_calc = lambda se, iv: 1.2
    assert se.xn(0) == 1.2


def test_march_reproduces_the_sine_and_reset_gives_zero():
    nstep, xcrd, svr = _build()
    svr.march_alpha2(nstep * 10)
    np.testing.assert_allclose(svr.get_so0(0), np.sin(xcrd), rtol=0, atol=1.e-14)
    np.testing.assert_allclose(svr.get_cfl(), np.ones(svr.grid.nselm), rtol=0, atol=1.e-14)
    svr.kernel.reset()
    svr.march_alpha2(nstep * 10)
    np.testing.assert_array_equal(svr.get_so0(0), np.zeros_like(xcrd))


def test_fine_interface_equals_march():
    nstep, xcrd, svr = _build()
    svr2 = _build()[-1]
    for _ in range(nstep * 2):
        svr.march_half_so0(odd_plane=False)
        svr.treat_boundary_so0()
        svr.update_cfl(odd_plane=True)
        svr.march_half_so1_alpha2(odd_plane=False)
        svr.treat_boundary_so1()
        svr.march_half_so0(odd_plane=True)
        svr.update_cfl(odd_plane=False)
        svr.march_half_so1_alpha2(odd_plane=True)
        svr2.march_alpha2(steps=1)
        assert svr.get_so0(0).ndarray.tolist() == svr2.get_so0(0).ndarray.tolist()


def test_callback_path_equals_the_c_port(bridge):
    """The same linear-wave kernel through the Python callbacks and through the oracle's C restatement of
    LinearScalarSolver: every slot bit-equal after 12 steps on a non-uniform grid."""
    from solvcon_b200 import spacetime as libst
    xloc = np.cumsum(0.1 + 0.05 * np.sin(np.arange(40)))
    dt = 0.04
    nstep, _, proto = _build()
    svr = libst.Solver(grid=libst.Grid(xloc=xloc), time_increment=dt, nvar=1)
    for n in ("xn_calc", "xp_calc", "tn_calc", "tp_calc", "so0p_calc", "cfl_updater"):
        setattr(svr.kernel, n, getattr(proto.kernel, n))
    chk = bridge.PortScalar1D(0, dt=dt, xloc=xloc)
    for a in (chk.so0, chk.so1, chk.cfl):
        a[...] = 0.0
    chk.so0[2::2] = 0.5 * np.sin(chk.x[2::2])
    chk.so1[2::2] = 0.5 * np.cos(chk.x[2::2])
    svr.so0.ndarray[:, 0] = chk.so0
    svr.so1.ndarray[:, 0] = chk.so1
    chk.march(2, 12)
    svr.march_alpha2(12)
    assert np.array_equal(svr.so0.ndarray[:, 0], chk.so0)
    assert np.array_equal(svr.so1.ndarray[:, 0], chk.so1)
    assert np.array_equal(svr.cfl.ndarray, chk.cfl)


def test_solver_proxy_subclass_marches_like_the_plain_solver():
    """tests/test_spacetime_solution.py:361-460: a SolverProxy subclass supplying the linear-wave hooks."""
    from solvcon_b200 import spacetime as libst
    nstep, xcrd, plain = _build()

    class LinearProxy(libst.SolverProxy):
        def _xn_calc(self, se, iv): return plain.kernel.xn_calc(se, iv)
        def _xp_calc(self, se, iv): return plain.kernel.xp_calc(se, iv)
        def _tn_calc(self, se, iv): return plain.kernel.tn_calc(se, iv)
        def _tp_calc(self, se, iv): return plain.kernel.tp_calc(se, iv)
        def _so0p_calc(self, se, iv): return plain.kernel.so0p_calc(se, iv)
        def _cfl_updater(self, se): return plain.kernel.cfl_updater(se)

    svr = LinearProxy(grid=libst.Grid(xcrd), time_increment=plain.time_increment, nvar=1)
    svr.set_so0(0, np.sin(xcrd))
    svr.set_so1(0, np.cos(xcrd))
    svr.setup_march()
    svr.march_alpha2(nstep * 10)
    np.testing.assert_allclose(svr.get_so0(0), np.sin(xcrd), rtol=0, atol=1.e-14)
    np.testing.assert_allclose(svr.get_cfl(), np.ones(svr.grid.nselm), rtol=0, atol=1.e-14)
    plain.march_alpha2(nstep * 10)
    assert svr.get_so0(0).ndarray.tolist() == plain.get_so0(0).ndarray.tolist()
    svr.kernel.reset()
    svr.march_alpha2(nstep)
    np.testing.assert_array_equal(svr.get_so0(0), np.zeros_like(xcrd))
    base = libst.SolverProxy(grid=libst.Grid(xcrd), time_increment=0.1, nvar=1)
    assert base.selm(0).xn(0) == 1.0
