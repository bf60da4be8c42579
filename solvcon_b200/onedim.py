"""Host-side mirror of ``_solvcon.onedim`` over the device library.

``Euler1DCore`` keeps the reference's binding surface
(``cpp/solvcon/onedim/pymod/wrap_onedim.cpp:38-150``): same constructor, same
ndarray properties (zero-copy views of host arrays the caller may write
through, exactly like the reference's ``to_ndarray`` views of ``SimpleArray``
memory), same phase/march method names, same error text.

Coherence protocol: the host arrays stay the API-visible
storage.  Every compute method uploads the arrays it reads, runs on the device,
and downloads the arrays it writes.  ``resident(True)`` suspends the transfers
(one upload on entry, explicit ``sync_to_host()``) for benchmarking and long
marches.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

PH_SO0, PH_CFL, PH_SO1, PH_BC0, PH_BC1 = range(5)


class Euler1DCore:
    """Device twin of ``solvcon::onedim::Euler1DCore`` (Euler1DCore.hpp:41-145)."""

    nvar = 3
    BOUND_COUNT = 2
    TINY = 1.e-100
    R = 8.31446261815324

    def __init__(self, ncoord, time_increment):
        ncoord = int(ncoord)
        if ncoord % 2 == 0:
            raise ValueError("ncoord cannot be even")  # Euler1DCore.cpp:25-28
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.scg_e1d_create(C.byref(self._h), ncoord, float(time_increment)))
        self._ncoord = ncoord
        self._dt = float(time_increment)
        # host storage, reference layout (Euler1DCore.cpp:29-34)
        # page-locked host twins (plain memory without a device): transfers run at the PCIe rate
        self._coord = _lib.pinned_zeros(ncoord)
        self._cfl = _lib.pinned_zeros(ncoord)
        self._so0 = _lib.pinned_zeros((ncoord, 3))
        self._so1 = _lib.pinned_zeros((ncoord, 3))
        self._gamma = np.full(ncoord, 1.4)
        self._resident = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.scg_e1d_destroy(h)
            self._h = None

    def __str__(self):
        # operator<<, Euler1DCore.cpp:16-20
        return f"Euler1DCore(ncoord={self._ncoord}, time_increment={self._dt:g})"

    def clone(self):
        other = Euler1DCore(self._ncoord, self._dt)
        if self._resident:
            self.sync_to_host()
        for n in ("_coord", "_cfl", "_so0", "_so1", "_gamma"):
            getattr(other, n)[...] = getattr(self, n)
        return other

    # -- scalars -----------------------------------------------------------
    time_increment = property(lambda self: self._dt)
    ncoord = property(lambda self: self._ncoord)

    # -- arrays (views, writable through) ------------------------------------
    coord = property(lambda self: self._coord)
    cfl = property(lambda self: self._cfl)
    so0 = property(lambda self: self._so0)
    so1 = property(lambda self: self._so1)
    gamma = property(lambda self: self._gamma)

    # -- derived fields (Euler1DCore.hpp:106-168; plain numpy on the host copy) ----
    @property
    def density(self):
        return self._so0[:, 0].copy()

    @property
    def velocity(self):
        return self._so0[:, 1] / (self._so0[:, 0] + self.TINY)

    @property
    def pressure(self):
        ret = self._so0[:, 1].copy()
        ret *= ret
        ret /= 2.0 * self._so0[:, 0] + self.TINY
        ret = self._so0[:, 2] - ret
        ret *= self._gamma - 1.0
        return ret

    @property
    def temperature(self):
        ret = (self._gamma - 1.0) / self.R
        with np.errstate(divide="ignore", invalid="ignore"):
            ret = ret * (self._so0[:, 2] / self._so0[:, 0]
                         - 0.5 * self._so0[:, 1] ** 2 / self._so0[:, 0] ** 2)
        return ret

    @property
    def internal_energy(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            ret = self._so0[:, 2] / self._so0[:, 0]
            ret = ret - 0.5 * (self._so0[:, 1] ** 2 / self._so0[:, 0] ** 2)
        return ret

    @property
    def entropy(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            return self.pressure / np.power(self.density, self._gamma)

    # -- coherence -----------------------------------------------------------
    def resident(self, flag=True):
        """Keep the field on the device between calls (uploads once now)."""
        if flag and not self._resident:
            self._push(True)
        if not flag and self._resident:
            self.sync_to_host()
        self._resident = bool(flag)
        return self

    def sync_to_host(self):
        _lib.check(self._lib.scg_e1d_download(self._h, _lib.ptr(self._so0), _lib.ptr(self._so1),
                                              _lib.ptr(self._cfl)))

    def _push(self, force=False):
        if self._resident and not force:
            return
        _lib.check(self._lib.scg_e1d_upload(self._h, _lib.ptr(self._coord), _lib.ptr(self._gamma),
                                            _lib.ptr(self._so0), _lib.ptr(self._so1), _lib.ptr(self._cfl)))

    def _pull(self):
        if not self._resident:
            self.sync_to_host()

    def _phase(self, phase, odd_plane=False, alpha=2):
        self._push()
        _lib.check(self._lib.scg_e1d_phase(self._h, phase, int(bool(odd_plane)), alpha))
        self._pull()

    # -- phases (wrap_onedim.cpp:107-150) ----------------------------------------
    def update_cfl(self, odd_plane):
        self._phase(PH_CFL, odd_plane)

    def march_half_so0(self, odd_plane):
        self._phase(PH_SO0, odd_plane)

    def treat_boundary_so0(self):
        self._phase(PH_BC0)

    def treat_boundary_so1(self):
        self._phase(PH_BC1)

    def setup_march(self):
        self.update_cfl(False)  # Euler1DCore.hpp:126

    def _half(self, alpha, odd):
        # march_half1_alpha / march_half2_alpha, Euler1DCore.hpp:322-337
        self._push()
        for ph in (PH_SO0, PH_CFL, PH_SO1):
            _lib.check(self._lib.scg_e1d_phase(self._h, ph, int(odd), alpha))
        self._pull()

    def _march(self, alpha, steps):
        self._push()
        _lib.check(self._lib.scg_e1d_march(self._h, alpha, int(steps)))
        self._pull()

    def march_half_so1_alpha1(self, odd_plane):
        self._so1_phase(1, odd_plane)

    def march_half_so1_alpha2(self, odd_plane):
        self._so1_phase(2, odd_plane)

    def _so1_phase(self, alpha, odd_plane):
        self._phase(PH_SO1, odd_plane, alpha)

    def march_half1_alpha1(self):
        self._half(1, False)

    def march_half1_alpha2(self):
        self._half(2, False)

    def march_half2_alpha1(self):
        self._half(1, True)

    def march_half2_alpha2(self):
        self._half(2, True)

    def march_alpha1(self, steps):
        self._march(1, steps)

    def march_alpha2(self, steps):
        self._march(2, steps)

    # -- additive API: device-side reductions of the callers' numpy min/max/sum ------
    def reduce(self):
        """(cfl_min, cfl_max, sum of density) computed on the device
        (callers: solvcon/pilot/onedim/_euler1d.py:146-147)."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        if not self._resident:
            self._push()
        _lib.check(self._lib.scg_e1d_reduce(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value
