"""Host-side mirror of ``_solvcon.spacetime`` for the two concrete 1D scalar solvers.

Mirrors ``cpp/solvcon/spacetime/pymod/wrap_spacetime.{hpp,cpp}``: ``Grid``,
``LinearScalarSolver``, ``InviscidBurgersSolver`` and their ``Selm``/``Celm`` cursors, with the
reference's names, argument meaning and error text.  The march and every phase method run on
the device through ``libscg_cuda.so``; the cursor classes are per-element *inspection* helpers
over the host arrays (the reference's front ends use them to set initial conditions,
``solvcon/onedim/linear_scalar.py:36-45``) and are never on the march path.

The generic ``Solver`` whose kernel is a set of Python callables
(``spacetime/core.hpp:223-273``) is host-only by nature and is not part of the device path.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib
from .simplearray import SimpleArray

PH_SO0, PH_CFL, PH_SO1, PH_BC0, PH_BC1 = range(5)
_DBL_MIN = sys.float_info.min


def _fmt(v):
    """std::format("{}", double): shortest round-trip, integral values without '.0'."""
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    s = repr(float(v))
    return s[:-2] if s.endswith(".0") else s


class Grid:
    """``spacetime::Grid`` (core.hpp:98-221, core.cpp:14-99)."""

    BOUND_COUNT = 2

    def __init__(self, xmin=None, xmax=None, nelm=None, xloc=None):
        if xloc is None and xmin is not None and not np.isscalar(xmin):
            xloc, xmin = xmin, None
        if xloc is not None:
            self._init_from_array(np.ascontiguousarray(xloc, dtype=np.float64))
            return
        ncelm = int(nelm)
        xmin, xmax = float(xmin), float(xmax)
        if ncelm < 1:
            raise ValueError(f"Grid::Grid(xmin={_fmt(xmin)}, xmax={_fmt(xmax)}, ncelm={ncelm}) "
                             "invalid argument: ncelm smaller than 1")
        if xmin >= xmax:
            raise ValueError(f"Grid::Grid(xmin={_fmt(xmin)}, xmax={_fmt(xmax)}, ncelm={ncelm}) "
                             "invalid arguments: xmin >= xmax")
        # core.cpp:36-43: cumulative addition (np.add.accumulate is strictly sequential, so the
        # bits match the reference's loop), last boundary pinned to xmax
        xspace = (xmax - xmin) / ncelm
        seed = np.full(ncelm + 1, xspace)
        seed[0] = xmin
        xloc = np.add.accumulate(seed)
        xloc[ncelm] = xmax
        self._init_from_array(xloc)

    def _init_from_array(self, xloc):
        if xloc.size < 2:
            raise ValueError("Grid::init_from_array(xloc) invalid arguments: "
                             f"xloc.size()={xloc.size} smaller than 2")
        bad = np.nonzero(xloc[:-1] >= xloc[1:])[0]
        if bad.size:
            it = int(bad[0])
            raise ValueError("Grid::init_from_array(xloc) invalid arguments: "
                             f"xloc[{it}]={_fmt(xloc[it])} >= xloc[{it + 1}]={_fmt(xloc[it + 1])}")
        n = xloc.size - 1
        self._ncelm = n
        self._xmin = float(xloc[0])
        self._xmax = float(xloc[n])
        nx = 2 * n + 5
        x = _lib.pinned_zeros(nx)
        x[2:2 * n + 3:2] = xloc                                  # core.cpp:77-80
        x[3:2 * n + 2:2] = (x[2:2 * n + 1:2] + x[4:2 * n + 3:2]) / 2  # core.cpp:82-86
        for it in (1, 2):                                        # core.cpp:88-98
            x[2 - it] = x[2] + x[2] - x[2 + it]
            x[nx - 3 + it] = x[nx - 3] + x[nx - 3] - x[nx - 3 - it]
        self._x = SimpleArray(array=x)

    xmin = property(lambda s: s._xmin)
    xmax = property(lambda s: s._xmax)
    ncelm = property(lambda s: s._ncelm)
    nselm = property(lambda s: s._ncelm + 1)
    xcoord = property(lambda s: s._x)
    xsize = property(lambda s: s._x.size)

    def clone(self):
        g = Grid.__new__(Grid)
        g._ncelm, g._xmin, g._xmax = self._ncelm, self._xmin, self._xmax
        g._x = self._x.clone()
        return g

    def __str__(self):
        return f"Grid(xmin={_fmt(self._xmin)}, xmax={_fmt(self._xmax)}, ncelm={self._ncelm})"


class Selm:
    """Solution-element cursor (core.hpp:444-503) over the host arrays."""

    def __init__(self, svr, index, odd_plane):
        self._svr = svr
        self._k = 2 + 2 * int(index) + (1 if odd_plane else 0)

    def duplicate(self):
        s = type(self).__new__(type(self))
        s._svr, s._k = self._svr, self._k
        return s

    dup = property(lambda s: s.duplicate())  # wrap_spacetime.hpp:61-62

    def _x(self, off=0):
        return float(self._svr.grid.xcoord.ndarray[self._k + off])

    x = property(lambda s: s._x())
    xneg = property(lambda s: s._x(-1))
    xpos = property(lambda s: s._x(+1))
    dx = property(lambda s: s._x(+1) - s._x(-1))
    dxneg = property(lambda s: s._x() - s._x(-1))
    dxpos = property(lambda s: s._x(+1) - s._x())
    xctr = property(lambda s: (s._x(-1) + s._x(+1)) / 2)
    index = property(lambda s: (s._k >> 1) - 1)
    on_odd_plane = property(lambda s: bool((s._k - 2) & 1))
    on_even_plane = property(lambda s: not ((s._k - 2) & 1))
    time_increment = property(lambda s: s._svr.time_increment)
    dt = property(lambda s: s._svr.dt)
    hdt = property(lambda s: s._svr.hdt)
    qdt = property(lambda s: s._svr.qdt)
    grid = property(lambda s: s._svr.grid)

    def __str__(self):
        plane = "odd" if self.on_odd_plane else "even"
        return (f"{self._svr._elm_prefix}Selm({plane}, index={self.index}, x={_fmt(self.x)}, "
                f"xneg={_fmt(self.xneg)}, xpos={_fmt(self.xpos)})")

    def move_at(self, offset):
        k = self._k + int(offset)
        hi = self._svr.grid.xsize - 1
        if k < 1 or k >= hi:
            raise IndexError(f"Selm(xindex={self._k})::move_at(offset={int(offset)}): "
                             f"xindex = {k} outside the interval [1, {hi})")
        self._k = k
        return self

    # the Python names bind the range-checked forms (wrap_spacetime.hpp:77-91)
    def move(self, offset): return self.move_at(offset)
    def move_left(self): return self.move_at(-2)
    def move_right(self): return self.move_at(2)
    def move_neg(self): return self.move_at(-1)
    def move_pos(self): return self.move_at(1)
    def move_left_at(self): return self.move_at(-2)
    def move_right_at(self): return self.move_at(2)
    def move_neg_at(self): return self.move_at(-1)
    def move_pos_at(self): return self.move_at(1)

    # py::self <op> py::self (wrap_spacetime.hpp:54-60): defined between cursors of one kind only; cursors of
    # different solvers are unordered and unequal (ElementBase comparisons, core.hpp)
    def _same(self, o): return type(o) is type(self) and self._svr is o._svr
    def __eq__(self, o): return (self._same(o) and self._k == o._k) if isinstance(o, (Selm, Celm)) else NotImplemented
    def __ne__(self, o): return (not (self._same(o) and self._k == o._k)) if isinstance(o, (Selm, Celm)) else NotImplemented
    def __lt__(self, o): return (self._svr is o._svr and self._k < o._k) if type(o) is type(self) else NotImplemented
    def __le__(self, o): return (self._svr is o._svr and self._k <= o._k) if type(o) is type(self) else NotImplemented
    def __gt__(self, o): return (self._svr is o._svr and self._k > o._k) if type(o) is type(self) else NotImplemented
    def __ge__(self, o): return (self._svr is o._svr and self._k >= o._k) if type(o) is type(self) else NotImplemented
    __hash__ = None

    def _iv(self, iv):
        if iv >= self._svr.nvar:
            raise IndexError("out of nvar range")
        return int(iv)

    def get_so0(self, iv): return float(self._svr.so0.ndarray[self._k, self._iv(iv)])
    def get_so1(self, iv): return float(self._svr.so1.ndarray[self._k, self._iv(iv)])
    def get_cfl(self): return float(self._svr.cfl.ndarray[self._k])

    def set_so0(self, iv, v): self._svr.so0.ndarray[self._k, self._iv(iv)] = v
    def set_so1(self, iv, v): self._svr.so1.ndarray[self._k, self._iv(iv)] = v
    def set_cfl(self, v): self._svr.cfl.ndarray[self._k] = v

    # per-element flux accessors (kernel/linear_scalar.hpp:45-87, inviscid_burgers.hpp:51-109);
    # inspection only -- the march computes these on the device.
    def _kernel(self):
        return getattr(self._svr, "kernel", None)   # only the generic Solver carries a Kernel of Python callables

    def xn(self, iv):
        if self._kernel() is not None:
            return self._kernel().xn_calc(self, iv)
        disp = 0.5 * (self.x + self.xneg) - self.xctr
        return self.dxneg * (self.get_so0(iv) + disp * self.get_so1(iv))

    def xp(self, iv):
        if self._kernel() is not None:
            return self._kernel().xp_calc(self, iv)
        disp = 0.5 * (self.x + self.xpos) - self.xctr
        return self.dxpos * (self.get_so0(iv) + disp * self.get_so1(iv))

    def _t(self, iv, sign):
        if self._kernel() is not None:
            return (self._kernel().tn_calc if sign > 0 else self._kernel().tp_calc)(self, iv)
        disp = self.x - self.xctr
        u, ux = self.get_so0(iv), self.get_so1(iv)
        if self._svr._eq == 0:
            ret = u
            ret += disp * ux
            ret += sign * (self.qdt * ux)
        else:
            u2 = u * u
            ret = 0.5 * u2
            ret += disp * u * ux
            ret += sign * (self.qdt * u2 * ux)
        return self.hdt * ret

    def tn(self, iv): return self._t(iv, +1.0)
    def tp(self, iv): return self._t(iv, -1.0)

    def so0p(self, iv):
        if self._kernel() is not None:
            return self._kernel().so0p_calc(self, iv)
        ret = self.get_so0(iv)
        ret += (self.x - self.xctr) * self.get_so1(iv)
        ret -= self.hdt * self.get_so1(iv)
        return ret

    def update_cfl(self):
        if self._kernel() is not None:
            self._kernel().cfl_updater(self)
            return
        hdx = min(self.dxneg, self.dxpos)
        if self._svr._eq == 0:
            self.set_cfl(self.hdt / hdx)
        else:
            self.set_cfl(abs(self.get_so0(0)) * self.hdt / hdx)


class Celm:
    """Conservation-element cursor (core.hpp:527-610)."""

    def __init__(self, svr, index, odd_plane):
        self._svr = svr
        self._k = 3 + 2 * int(index) + (1 if odd_plane else 0)

    def duplicate(self):
        c = type(self).__new__(type(self))
        c._svr, c._k = self._svr, self._k
        return c

    dup = property(lambda s: s.duplicate())

    def _x(self, off=0):
        return float(self._svr.grid.xcoord.ndarray[self._k + off])

    x = property(lambda s: s._x())
    xctr = property(lambda s: s._x())
    xneg = property(lambda s: s._x(-1))
    xpos = property(lambda s: s._x(+1))
    dx = property(lambda s: s._x(+1) - s._x(-1))
    index = property(lambda s: ((s._k - 1) >> 1) - 1)
    on_odd_plane = property(lambda s: bool((s._k - 3) & 1))
    on_even_plane = property(lambda s: not ((s._k - 3) & 1))
    time_increment = property(lambda s: s._svr.time_increment)
    dt = property(lambda s: s._svr.dt)
    hdt = property(lambda s: s._svr.hdt)
    qdt = property(lambda s: s._svr.qdt)
    grid = property(lambda s: s._svr.grid)

    def __str__(self):
        plane = "odd" if self.on_odd_plane else "even"
        return (f"{self._svr._elm_prefix}Celm({plane}, index={self.index}, x={_fmt(self.x)}, "
                f"xneg={_fmt(self.xneg)}, xpos={_fmt(self.xpos)})")

    def move_at(self, offset):
        k = self._k + int(offset)
        hi = self._svr.grid.xsize - 2
        if k < 2 or k >= hi:
            raise IndexError(f"Celm(xindex={self._k})::move_at(offset={int(offset)}): "
                             f"xindex = {k} outside the interval [2, {hi})")
        self._k = k
        return self

    # the Python names bind the range-checked forms (wrap_spacetime.hpp:77-91)
    def move(self, offset): return self.move_at(offset)
    def move_left(self): return self.move_at(-2)
    def move_right(self): return self.move_at(2)
    def move_neg(self): return self.move_at(-1)
    def move_pos(self): return self.move_at(1)
    def move_left_at(self): return self.move_at(-2)
    def move_right_at(self): return self.move_at(2)
    def move_neg_at(self): return self.move_at(-1)
    def move_pos_at(self): return self.move_at(1)

    # py::self <op> py::self (wrap_spacetime.hpp:54-60): defined between cursors of one kind only; cursors of
    # different solvers are unordered and unequal (ElementBase comparisons, core.hpp)
    def _same(self, o): return type(o) is type(self) and self._svr is o._svr
    def __eq__(self, o): return (self._same(o) and self._k == o._k) if isinstance(o, (Selm, Celm)) else NotImplemented
    def __ne__(self, o): return (not (self._same(o) and self._k == o._k)) if isinstance(o, (Selm, Celm)) else NotImplemented
    def __lt__(self, o): return (self._svr is o._svr and self._k < o._k) if type(o) is type(self) else NotImplemented
    def __le__(self, o): return (self._svr is o._svr and self._k <= o._k) if type(o) is type(self) else NotImplemented
    def __gt__(self, o): return (self._svr is o._svr and self._k > o._k) if type(o) is type(self) else NotImplemented
    def __ge__(self, o): return (self._svr is o._svr and self._k >= o._k) if type(o) is type(self) else NotImplemented
    __hash__ = None

    def _selm_at_k(self, k):
        s = Selm.__new__(Selm)
        s._svr, s._k = self._svr, k
        return s

    selm_xn = property(lambda s: s._selm_at_k(s._k - 1))
    selm_xp = property(lambda s: s._selm_at_k(s._k + 1))
    selm_tn = property(lambda s: s._selm_at_k(s._k))
    selm_tp = property(lambda s: s._selm_at_k(s._k))

    def calc_so0(self, iv):
        # Celm::calc_so0, core.hpp:645-654
        sn, sp = self.selm_xn, self.selm_xp
        flux_ll = sn.xp(iv) + sn.tp(iv)
        flux_ur = sp.xn(iv) - sp.tp(iv)
        return (flux_ll + flux_ur) / self.selm_tp.dx

    def _calc_so1(self, iv, alpha):
        # Celm::calc_so1_alpha, core.hpp:656-672
        sn, sp = self.selm_xn, self.selm_xp
        upn, upp = sn.so0p(iv), sp.so0p(iv)
        utp = self.selm_tp.get_so0(iv)
        duxn = (utp - upn) / sn.dxpos
        duxp = (upp - utp) / sp.dxneg
        fan = 1.0 if alpha == 0 else abs(duxn) ** alpha
        fap = 1.0 if alpha == 0 else abs(duxp) ** alpha
        return (fap * duxn + fan * duxp) / (fap + fan + _DBL_MIN)

    def calc_so1_alpha0(self, iv): return self._calc_so1(iv, 0)
    def calc_so1_alpha1(self, iv): return self._calc_so1(iv, 1)
    def calc_so1_alpha2(self, iv): return self._calc_so1(iv, 2)


class SolverElementIterator:
    """``SolverElementIterator`` (wrap_spacetime.hpp:184-263): walks the cursors of one plane."""

    def __init__(self, svr, odd_plane, selm):
        self._svr, self._odd, self._selm, self._cur = svr, bool(odd_plane), bool(selm), 0

    @property
    def nelem(self):
        g = self._svr.grid
        return (g.nselm if self._selm else g.ncelm) - (1 if self._odd else 0)

    def __iter__(self):
        return self

    def __next__(self):
        if self._cur >= self.nelem:
            raise StopIteration
        e = (self._svr.selm if self._selm else self._svr.celm)(self._cur, self._odd)
        self._cur += 1
        return e

    def __str__(self):
        return (f"SolverElementIterator({'selm' if self._selm else 'celm'}, "
                f"{'on_odd_plane' if self._odd else 'on_even_plane'}, current={self._cur}, nelem={self.nelem})")


class _ScalarSolver:
    """Shared body of ``LinearScalarSolver`` / ``InviscidBurgersSolver``
    (SolverBase, core.hpp:678-1036; binding wrap_spacetime.hpp:265-435)."""

    _eq = 0

    def __init__(self, grid, time_increment):
        self._lib = _lib.load()
        self._grid = grid
        n = grid.ncelm
        nx = 2 * n + 5
        self._h = C.c_void_p()
        _lib.check(self._lib.scg_st_create(C.byref(self._h), self._eq, n, float(time_increment)))
        self._dt = float(time_increment)
        self._so0 = SimpleArray((nx, 1), pinned=True)
        self._so1 = SimpleArray((nx, 1), pinned=True)
        self._cfl = SimpleArray((nx,), pinned=True)
        self._resident = False
        self._x_uploaded = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.scg_st_destroy(h)
            self._h = None

    def clone(self, grid=False):
        other = type(self)(self._grid.clone() if grid else self._grid, self._dt)
        if self._resident:
            self.sync_to_host()
        other._so0.ndarray[...] = self._so0.ndarray
        other._so1.ndarray[...] = self._so1.ndarray
        other._cfl.ndarray[...] = self._cfl.ndarray
        return other

    # -- scalars ----------------------------------------------------------------
    grid = property(lambda s: s._grid)
    nvar = property(lambda s: 1)
    dt = property(lambda s: s._dt)
    hdt = property(lambda s: 0.5 * s._dt)   # Field::set_time_increment, core.cpp:111-116
    qdt = property(lambda s: 0.25 * s._dt)

    @property
    def time_increment(self):
        return self._dt

    @time_increment.setter
    def time_increment(self, v):
        self._dt = float(v)
        _lib.check(self._lib.scg_st_set_time_increment(self._h, self._dt))

    # -- arrays -----------------------------------------------------------------
    so0 = property(lambda s: s._so0)
    so1 = property(lambda s: s._so1)
    cfl = property(lambda s: s._cfl)

    def _plane_slice(self, odd_plane):
        n = self._grid.ncelm
        # selm(it, odd) for it in [0, nselm - odd): k = 2 + 2 it + odd
        return slice(3, 2 * n + 2, 2) if odd_plane else slice(2, 2 * n + 3, 2)

    def x(self, odd_plane=False):
        return SimpleArray(array=self._grid.xcoord.ndarray[self._plane_slice(odd_plane)].copy())

    def xctr(self, odd_plane=False):
        xc = self._grid.xcoord.ndarray
        sl = self._plane_slice(odd_plane)
        lo = slice(sl.start - 1, sl.stop - 1, 2)
        hi = slice(sl.start + 1, sl.stop + 1, 2)
        return SimpleArray(array=(xc[lo] + xc[hi]) / 2)

    def _check_iv(self, iv, what):
        if iv >= 1:
            raise IndexError(f"{what}(): out of nvar range")

    def get_so0(self, iv, odd_plane=False):
        self._check_iv(iv, "get_so0")
        return SimpleArray(array=self._so0.ndarray[self._plane_slice(odd_plane), iv].copy())

    def get_so1(self, iv, odd_plane=False):
        self._check_iv(iv, "get_so1")
        return SimpleArray(array=self._so1.ndarray[self._plane_slice(odd_plane), iv].copy())

    def get_cfl(self, odd_plane=False):
        return SimpleArray(array=self._cfl.ndarray[self._plane_slice(odd_plane)].copy())

    def get_so0p(self, iv, odd_plane=False):
        self._check_iv(iv, "get_so0p")
        xc = self._grid.xcoord.ndarray
        sl = self._plane_slice(odd_plane)
        lo = slice(sl.start - 1, sl.stop - 1, 2)
        hi = slice(sl.start + 1, sl.stop + 1, 2)
        ret = self._so0.ndarray[sl, iv].copy()
        ret += (xc[sl] - (xc[lo] + xc[hi]) / 2) * self._so1.ndarray[sl, iv]
        ret -= self.hdt * self._so1.ndarray[sl, iv]
        return SimpleArray(array=ret)

    def _set(self, target, name, arr, odd_plane):
        a = np.asarray(arr.ndarray if isinstance(arr, SimpleArray) else arr, dtype=np.float64)
        if a.ndim != 1:
            raise IndexError(f"{name}(): input not 1D")
        nselm = self._grid.nselm - (1 if odd_plane else 0)
        if a.size != nselm:
            raise IndexError(f"{name}(): arr size {a.size} != nselm {nselm}")
        target[self._plane_slice(odd_plane)] = a

    def set_so0(self, iv, arr, odd_plane=False):
        if iv >= 1:
            raise IndexError(f"set_so0(): iv {iv} >= nvar 1")
        self._set(self._so0.ndarray[:, iv], "set_so0", arr, odd_plane)

    def set_so1(self, iv, arr, odd_plane=False):
        self._check_iv(iv, "set_so1")
        self._set(self._so1.ndarray[:, iv], "set_so1", arr, odd_plane)

    def set_cfl(self, arr, odd_plane=False):
        self._set(self._cfl.ndarray, "set_so1", arr, odd_plane)

    # -- cursors ------------------------------------------------------------------
    def selm(self, ielm, odd_plane=False):
        s = Selm(self, ielm, odd_plane)
        hi = self._grid.xsize - 1
        if s._k < 1 or s._k >= hi:
            raise IndexError(f"Field::selm_at(ielm={int(ielm)}, odd_plane={int(bool(odd_plane))}): "
                             f"xindex = {s._k} outside the interval [1, {hi})")
        return s

    def celm(self, ielm, odd_plane=False):
        c = Celm(self, ielm, odd_plane)
        hi = self._grid.xsize - 2
        if c._k < 2 or c._k >= hi:
            raise IndexError(f"Field::celm_at(ielm={int(ielm)}, odd_plane={int(bool(odd_plane))}): "
                             f"xindex = {c._k} outside the interval [2, {hi})")
        return c

    def selms(self, odd_plane=False):
        return SolverElementIterator(self, odd_plane, True)

    def celms(self, odd_plane=False):
        return SolverElementIterator(self, odd_plane, False)

    _elm_prefix = ""

    def __str__(self):
        # operator<<, spacetime/io.cpp:28-32,48-52,68-72
        return f"{type(self).__name__}(grid={self._grid})"

    # -- coherence ------------------------------------------------------------------
    def resident(self, flag=True):
        if flag and not self._resident:
            self._push(True)
        if not flag and self._resident:
            self.sync_to_host()
        self._resident = bool(flag)
        return self

    def sync_to_host(self):
        _lib.check(self._lib.scg_st_download(self._h, None, _lib.ptr(self._so0.ndarray),
                                             _lib.ptr(self._so1.ndarray), _lib.ptr(self._cfl.ndarray)))

    def _push(self, force=False):
        if self._resident and not force:
            return
        _lib.check(self._lib.scg_st_upload(self._h, _lib.ptr(self._grid.xcoord.ndarray),
                                           _lib.ptr(self._so0.ndarray), _lib.ptr(self._so1.ndarray),
                                           _lib.ptr(self._cfl.ndarray)))

    def _pull(self):
        if not self._resident:
            self.sync_to_host()

    def _phase(self, phase, odd_plane=False, alpha=2):
        self._push()
        _lib.check(self._lib.scg_st_phase(self._h, phase, int(bool(odd_plane)), alpha))
        self._pull()

    # -- phases (wrap_spacetime.hpp:380-424) -----------------------------------------------
    def update_cfl(self, odd_plane): self._phase(PH_CFL, odd_plane)
    def march_half_so0(self, odd_plane): self._phase(PH_SO0, odd_plane)
    def treat_boundary_so0(self): self._phase(PH_BC0)
    def treat_boundary_so1(self): self._phase(PH_BC1)
    def setup_march(self): self.update_cfl(False)

    def _half1(self, alpha):
        # march_half1_alpha, core.hpp:1006-1015
        self._push()
        for ph, odd in ((PH_SO0, 0), (PH_BC0, 0), (PH_CFL, 1), (PH_SO1, 0), (PH_BC1, 0)):
            _lib.check(self._lib.scg_st_phase(self._h, ph, odd, alpha))
        self._pull()

    def _half2(self, alpha):
        # march_half2_alpha, core.hpp:1017-1025
        self._push()
        for ph, odd in ((PH_SO0, 1), (PH_CFL, 0), (PH_SO1, 1)):
            _lib.check(self._lib.scg_st_phase(self._h, ph, odd, alpha))
        self._pull()

    def _march(self, alpha, steps):
        self._push()
        _lib.check(self._lib.scg_st_march(self._h, alpha, int(steps)))
        self._pull()

    def reduce(self):
        """(cfl_min, cfl_max, sum of so0 over the even plane) computed on the device."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        if not self._resident:
            self._push()
        _lib.check(self._lib.scg_st_reduce(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


def _add_alpha_methods(cls):
    for alpha in (0, 1, 2):
        def so1(self, odd_plane, _a=alpha): self._phase(PH_SO1, odd_plane, _a)
        def h1(self, _a=alpha): self._half1(_a)
        def h2(self, _a=alpha): self._half2(_a)
        def m(self, steps, _a=alpha): self._march(_a, steps)
        setattr(cls, f"march_half_so1_alpha{alpha}", so1)
        setattr(cls, f"march_half1_alpha{alpha}", h1)
        setattr(cls, f"march_half2_alpha{alpha}", h2)
        setattr(cls, f"march_alpha{alpha}", m)


_add_alpha_methods(_ScalarSolver)


class LinearScalarSolver(_ScalarSolver):
    """``spacetime::LinearScalarSolver`` (kernel/linear_scalar.hpp:29-43) on the device."""
    _eq = 0
    _elm_prefix = "LinearScalar"

    def __init__(self, grid, time_increment):
        super().__init__(grid, time_increment)


class InviscidBurgersSolver(_ScalarSolver):
    """``spacetime::InviscidBurgersSolver`` (kernel/inviscid_burgers.hpp:33-47) on the device."""
    _eq = 1
    _elm_prefix = "InviscidBurgers"

    def __init__(self, grid, time_increment):
        super().__init__(grid, time_increment)


class Kernel:
    """``spacetime::Kernel`` (core.hpp:223-273, 505-519): six callables ``xn_calc, xp_calc, tn_calc, tp_calc,
    so0p_calc(se, iv) -> float`` and ``cfl_updater(se)``; ``reset()`` restores the defaults (zeros, identity, cfl = 0)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.xn_calc = lambda se, iv: 0.0
        self.xp_calc = lambda se, iv: 0.0
        self.tn_calc = lambda se, iv: 0.0
        self.tp_calc = lambda se, iv: 0.0
        self.so0p_calc = lambda se, iv: se.get_so0(iv)
        self.cfl_updater = lambda se: se.set_cfl(0.0)


class Solver(_ScalarSolver):
    """The generic ``spacetime::Solver`` (core.hpp:1038-1053; binding wrap_spacetime.cpp:130-164) whose physics is a
    :class:`Kernel` of PYTHON CALLABLES.  It is host-only BY CONSTRUCTION, here as in the reference: the march calls back
    into Python once per solution element, which a device cannot do.  It is part of the drop-in surface (the reference's
    tests drive it, tests/test_spacetime_solution.py) and is NOT a fallback of the device solvers --
    ``LinearScalarSolver`` / ``InviscidBurgersSolver`` never come here.  The loops are those of ``SolverBase``
    (core.hpp:946-1036) over the cursor classes above."""

    _eq = -1

    def __init__(self, grid, time_increment, nvar=1):
        self._grid = grid
        self._nvar = int(nvar)
        nx = 2 * grid.ncelm + 5
        self._dt = float(time_increment)
        self._so0 = SimpleArray((nx, self._nvar))
        self._so1 = SimpleArray((nx, self._nvar))
        self._cfl = SimpleArray((nx,))
        self.kernel = Kernel()
        self._h = None
        self._resident = False

    def __del__(self):
        pass

    nvar = property(lambda s: s._nvar)

    @property
    def time_increment(self):
        return self._dt

    @time_increment.setter
    def time_increment(self, v):
        self._dt = float(v)

    def clone(self, grid=False):
        other = Solver(self._grid.clone() if grid else self._grid, self._dt, self._nvar)
        for n in ("_so0", "_so1", "_cfl"):
            getattr(other, n).ndarray[...] = getattr(self, n).ndarray
        k, o = self.kernel, other.kernel
        o.xn_calc, o.xp_calc, o.tn_calc, o.tp_calc, o.so0p_calc, o.cfl_updater = (
            k.xn_calc, k.xp_calc, k.tn_calc, k.tp_calc, k.so0p_calc, k.cfl_updater)
        return other

    def _check_iv(self, iv, what):
        if iv >= self._nvar:
            raise IndexError(f"{what}(): out of nvar range")

    def set_so0(self, iv, arr, odd_plane=False):
        if iv >= self._nvar:
            raise IndexError(f"set_so0(): iv {iv} >= nvar {self._nvar}")
        self._set(self._so0.ndarray[:, iv], "set_so0", arr, odd_plane)

    # -- coherence: nothing to do, the arrays ARE the state ----------------------------------------------
    def resident(self, flag=True): return self
    def sync_to_host(self): pass
    def _push(self, force=False): pass
    def _pull(self): pass

    def reduce(self):
        cfl = self.get_cfl().ndarray
        return float(cfl.min()), float(cfl.max()), float(self.get_so0(0).ndarray.sum())

    # -- SolverBase loops, core.hpp:946-1004 ---------------------------------------------------------------
    def march_half_so0(self, odd_plane):
        start = -1 if odd_plane else 0
        for ic in range(start, self._grid.ncelm):
            ce = Celm(self, ic, odd_plane)
            for iv in range(self._nvar):
                ce.selm_tp.set_so0(iv, ce.calc_so0(iv))

    def update_cfl(self, odd_plane):
        start = -1 if odd_plane else 0
        for ic in range(start, self._grid.nselm):
            Selm(self, ic, odd_plane).update_cfl()

    def _march_so1(self, odd_plane, alpha):
        start = -1 if odd_plane else 0
        for ic in range(start, self._grid.ncelm):
            ce = Celm(self, ic, odd_plane)
            for iv in range(self._nvar):
                ce.selm_tp.set_so1(iv, ce._calc_so1(iv, alpha))

    def _treat(self, arr):
        n = self._grid.ncelm
        left, right = Selm(self, -1, True), Selm(self, n, True)          # the two odd-plane ghosts
        for iv in range(self._nvar):
            arr[left._k, iv] = arr[Selm(self, n - 1, True)._k, iv]
            arr[right._k, iv] = arr[Selm(self, 0, True)._k, iv]

    def treat_boundary_so0(self): self._treat(self._so0.ndarray)
    def treat_boundary_so1(self): self._treat(self._so1.ndarray)

    def _half1(self, alpha):
        self.march_half_so0(False)
        self.treat_boundary_so0()
        self.update_cfl(True)
        self._march_so1(False, alpha)
        self.treat_boundary_so1()

    def _half2(self, alpha):
        self.march_half_so0(True)
        self.update_cfl(False)
        self._march_so1(True, alpha)

    def _march(self, alpha, steps):
        for _ in range(int(steps)):
            self._half1(alpha)
            self._half2(alpha)

    def _phase(self, phase, odd_plane=False, alpha=2):
        if phase == PH_SO1:
            self._march_so1(odd_plane, alpha)
        else:
            raise NotImplementedError


class SolverProxy:
    """``solvcon.spacetime.SolverProxy`` (solvcon/spacetime.py:42-72): owns a generic :class:`Solver` whose kernel hooks are
    this object's ``_xn_calc / _xp_calc / _tn_calc / _tp_calc / _so0p_calc / _cfl_updater`` methods -- subclass and
    override them; everything else is forwarded to the solver."""

    _HOOKS = ("xn_calc", "xp_calc", "tn_calc", "tp_calc", "so0p_calc", "cfl_updater")

    def __init__(self, *args, **kw):
        self.svr = Solver(*args, **kw)
        for name in self._HOOKS:
            setattr(self.svr.kernel, name, getattr(self, "_" + name))

    def _xn_calc(self, se, iv): return 1.0
    def _xp_calc(self, se, iv): return 1.0
    def _tn_calc(self, se, iv): return 1.0
    def _tp_calc(self, se, iv): return 1.0
    def _so0p_calc(self, se, iv): return se.get_so0(iv)
    def _cfl_updater(self, se): se.set_cfl(1.0)

    def __getattr__(self, name):
        return getattr(self.svr, name)


LinearScalarSelm = InviscidBurgersSelm = Selm
LinearScalarCelm = InviscidBurgersCelm = Celm
