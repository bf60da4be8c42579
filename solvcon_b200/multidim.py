"""Host-side mirror of ``_solvcon.EulerCore`` over the device library.

Keeps the reference's binding surface (``cpp/solvcon/multidim/pymod/wrap_multidim.cpp:247-408``):
``EulerCore(mesh, time_increment)``; ``ndim, ncell, ngstcell, neq, time_increment``; ``sigma0, taumin,
tauscale`` (rw); ``prepare_ce(), init_solution(gamma, rho, v, p), march(steps), march_substep(), update(),
calc_cfl(), calc_solt(), calc_soln(), calc_dsoln(), add_nonrefl(faces), add_slipwall(faces),
add_inlet(faces, value), clear_bc(), get_normal_matrix(ifc), bc_soln(), bc_dsoln()``; SimpleArray
properties ``cevol, cecnd, sfcnd, sfnml, so0c, so0n, so0t, so1c, so1n, stm, cflo, cflc, gamma`` whose
``.ndarray`` views alias solver memory and may be written through (``tests/multidim/euler/test_field.py:50-83``).

Coherence protocol: the host arrays stay the API-visible storage.  Every compute method
uploads the state (and the CE geometry if the caller changed it), runs on the device, and downloads
the state.  ``resident(True)`` suspends the transfers for long marches and benchmarking;
``sync_to_host()`` brings the field back.  There is no CPU path: without the CUDA library every method
raises.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib
from .simplearray import SimpleArray

(MD_UPDATE, MD_SOLT, MD_SOLN, MD_BC_SOLN, MD_CFL, MD_DSOLN, MD_BC_DSOLN, MD_PREPARE_CE, MD_SUBSTEP,
 MD_SUBSTEP_IFACE, MD_SUBSTEP_REST) = range(11)
BC_NONREFL, BC_SLIPWALL, BC_INLET = 0, 1, 2
ORDER_NONE, ORDER_MORTON, ORDER_RCM = 0, 1, 2
_ORDERS = {"none": 0, "morton": 1, "rcm": 2, 0: 0, 1: 1, 2: 2}

ARRAY_IDS = {"cevol": 0, "cecnd": 1, "sfcnd": 2, "sfnml": 3, "so0c": 4, "so0n": 5, "so0t": 6, "so1c": 7,
             "so1n": 8, "cflo": 9, "cflc": 10, "gamma": 11}
_GEOMETRY = ("cevol", "cecnd", "sfcnd", "sfnml")
_STATE = ("so0c", "so0n", "so0t", "so1c", "so1n", "cflo", "cflc", "gamma")
CLMFC, FCMND = 6, 4


def _refcount_baseline():
    """Reference counts of a SimpleArray held only by a dict, and of its ndarray, measured the way
    ``EulerCore._held`` reads them."""
    d = {"x": SimpleArray((1,))}
    sa = d.get("x")
    return sys.getrefcount(sa), sys.getrefcount(sa._a)


_REF_SA, _REF_ND = _refcount_baseline()


def _raise(exc=None):
    msg = _lib.load().scg_last_error().decode()
    if exc is None:
        exc = ValueError if msg.startswith("EulerCore::") else _lib.ScgError
    raise exc(msg)


def _table(mesh, name):
    """Full stored table (ghost rows first) of a mesh object as a C-contiguous ndarray."""
    t = getattr(mesh, name)
    a = t.ndarray if hasattr(t, "ndarray") else np.asarray(t)
    return np.ascontiguousarray(a)


class DeviceMesh:
    """Device copy of the StaticMesh tables the march reads, in the reordered device layout
    (``scg_mesh_create``).  ``order``: ``"none" | "morton" | "rcm"``."""

    def __init__(self, mesh, order="morton", owned=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.order = _ORDERS[order]
        own = None if owned is None else np.ascontiguousarray(owned, dtype=np.int8)
        self.ndim = int(mesh.ndim)
        self.ncell, self.nface, self.nnode = int(mesh.ncell), int(mesh.nface), int(mesh.nnode)
        self.ngstcell, self.ngstface, self.ngstnode = int(mesh.ngstcell), int(mesh.ngstface), int(mesh.ngstnode)
        if self.ngstcell == 0 and getattr(mesh, "nbound", 1) != 0:
            raise ValueError("EulerCore: the mesh has no ghost cells; call build_boundary() and build_ghost() first")
        if getattr(mesh, "on_device", 0) == 2:
            # the tables were built by the device builder (csrc/scg_hmesh_dev.cu) and are still there: no upload
            if self._lib.scg_mesh_create_from_hmesh(C.byref(self._h), mesh._h, self.order, _lib.ptr(own)):
                _raise(ValueError)
            return
        t = {n: _table(mesh, n) for n in ("cltpn", "clfcs", "fccls", "fcnds", "ndcrd", "fccnd", "fcnml", "clcnd")}
        for n in ("cltpn", "clfcs", "fccls", "fcnds"):
            if t[n].dtype != np.int32:
                t[n] = t[n].astype(np.int32)
        rc = self._lib.scg_mesh_create_part(
            C.byref(self._h), self.ndim, self.nnode, self.nface, self.ncell, self.ngstnode, self.ngstface,
            self.ngstcell, _lib.ptr(t["cltpn"]), _lib.ptr(t["clfcs"]), _lib.ptr(t["fccls"]), _lib.ptr(t["fcnds"]),
            _lib.ptr(t["ndcrd"]), _lib.ptr(t["fccnd"]), _lib.ptr(t["fcnml"]), _lib.ptr(t["clcnd"]), self.order,
            _lib.ptr(own))
        if rc:
            _raise(ValueError)

    @property
    def perm(self):
        """``perm[k]`` = reference cell id stored at device slot ``k``."""
        out = np.zeros(self.ncell, dtype=np.int32)
        _lib.check(self._lib.scg_mesh_perm(self._h, _lib.ptr(out)))
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.scg_mesh_destroy(h)
            self._h = None


class EulerCore:
    """Device twin of ``solvcon::EulerCore`` (``cpp/solvcon/multidim/euler.hpp:56-211``)."""

    def __init__(self, mesh, time_increment, order="morton", owned=None):
        self._lib = _lib.load()
        self._mesh = mesh
        self._dt = float(time_increment)
        self._h = C.c_void_p()
        if not isinstance(mesh, DeviceMesh) and int(mesh.ncell) == 0:
            # an empty mesh constructs (tests/test_euler.py:12-13 of the reference): zero-row arrays, nothing to march
            self._dmesh = None
            self._ndim, self._ncell, self._ngst = int(mesh.ndim), 0, 0
            self._neq = self._ndim + 2
        else:
            self._dmesh = mesh if isinstance(mesh, DeviceMesh) else DeviceMesh(mesh, order, owned)
            dm = self._dmesh
            self._ndim, self._ncell, self._ngst = dm.ndim, dm.ncell, dm.ngstcell
            self._neq = self._ndim + 2
            if self._lib.scg_md_create(C.byref(self._h), dm._h, self._dt):
                _raise(_lib.ScgError)
        T, nd, neq, nc, ng = self._ncell + self._ngst, self._ndim, self._neq, self._ncell, self._ngst
        # allocation shapes: euler_ce.cpp:31-60, euler_solution.cpp:19-41.  Host arrays are materialised
        # on first access (a 10^8-cell mesh never needs the padded host geometry unless the caller asks).
        self._shapes = {
            "cevol": ((T, CLMFC + 1), ng), "cecnd": ((T, (CLMFC + 1) * nd), ng),
            "sfcnd": ((nc, CLMFC * FCMND, nd), 0), "sfnml": ((nc, CLMFC * FCMND, nd), 0),
            "so0c": ((T, neq), ng), "so0n": ((T, neq), ng), "so0t": ((T, neq), ng),
            "so1c": ((T, neq, nd), ng), "so1n": ((T, neq, nd), ng), "stm": ((T, neq), ng),
            "cflo": ((T,), ng), "cflc": ((T,), ng), "gamma": ((T,), ng),
        }
        self._a = {}
        self._sigma0, self._taumin, self._tauscale = 3.0, 0.0, 1.0  # euler.hpp:184-188
        self._resident = False
        self._geo_seen = {}
        self._stale, self._exposed = set(), set()   # see "coherence" below
        self.h2d_bytes = self.d2h_bytes = 0

    def _get(self, name, expose=True):
        """The host SimpleArray of ``name``; on first access it is allocated and filled from the device, later it is
        refreshed when the device holds newer data.  ``expose``: the caller receives it and may write into it."""
        a = self._a.get(name)
        if a is not None and name in self._stale:
            self._download(name)
        if expose and name in ARRAY_IDS:
            self._exposed.add(name)
        if a is None:
            shape, ng = self._shapes[name]
            a = SimpleArray(shape, 0.0, nghost=ng, pinned=name in ARRAY_IDS and self._h.value is not None)
            self._a[name] = a
            if name in ARRAY_IDS and self._h.value is not None:
                _lib.check(self._lib.scg_md_download(self._h, ARRAY_IDS[name], _lib.ptr(a.ndarray)))
                self.d2h_bytes += a.ndarray.nbytes
            if name in _GEOMETRY:
                self._geo_seen[name] = a.ndarray.copy()
        return a

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.scg_md_destroy(h)
            self._h = None

    # -- scalars ---------------------------------------------------------------------------------
    ndim = property(lambda self: self._ndim)
    ncell = property(lambda self: self._ncell)
    ngstcell = property(lambda self: self._ngst)
    neq = property(lambda self: self._neq)
    time_increment = property(lambda self: self._dt)
    mesh = property(lambda self: self._mesh)

    def _set_param(self, name, v):
        setattr(self, name, float(v))
        _lib.check(self._lib.scg_md_set_params(self._h, self._sigma0, self._taumin, self._tauscale))

    sigma0 = property(lambda self: self._sigma0, lambda self, v: self._set_param("_sigma0", v))
    taumin = property(lambda self: self._taumin, lambda self, v: self._set_param("_taumin", v))
    tauscale = property(lambda self: self._tauscale, lambda self, v: self._set_param("_tauscale", v))

    # -- arrays ------------------------------------------------------------------------------------
    cevol = property(lambda self: self._get("cevol"))
    cecnd = property(lambda self: self._get("cecnd"))
    sfcnd = property(lambda self: self._get("sfcnd"))
    sfnml = property(lambda self: self._get("sfnml"))
    so0c = property(lambda self: self._get("so0c"))
    so0n = property(lambda self: self._get("so0n"))
    so0t = property(lambda self: self._get("so0t"))
    so1c = property(lambda self: self._get("so1c"))
    so1n = property(lambda self: self._get("so1n"))
    stm = property(lambda self: self._get("stm"))
    cflo = property(lambda self: self._get("cflo"))
    cflc = property(lambda self: self._get("cflc"))
    gamma = property(lambda self: self._get("gamma"))

    # -- coherence ---------------------------------------------------------------------------------
    # Host arrays are the API-visible storage, but only bytes somebody can observe travel:
    #   * a host array is UPLOADED before a compute call iff the caller may have written it since the last upload:
    #     it was handed out by its property since then (``_exposed``) or the caller still holds it / a view of it
    #     (reference count of the SimpleArray or of its ndarray, ``_held``);
    #   * after a compute call every array the device wrote is marked STALE; it is DOWNLOADED at once iff the caller
    #     holds it or a view of it (the reference's views alias live solver memory: held views stay current), else
    #     at the next access through its property.
    # So ``march(); svr.so0n`` moves so0n and nothing else, and a front end that holds views of everything gets
    # everything, exactly as it would read it from the reference's memory.  ``h2d_bytes`` / ``d2h_bytes`` count the
    # bytes actually copied.  ``resident(True)`` suspends all implicit transfers (the caller promises not to write the
    # host arrays); arrays are still refreshed when accessed through their property.
    def _held(self, name):
        sa = self._a.get(name)
        if sa is None:
            return False
        return sys.getrefcount(sa) > _REF_SA or sys.getrefcount(sa._a) > _REF_ND

    def _upload(self, name, rows=None):
        a = self._a[name].ndarray
        if rows is None:
            _lib.check(self._lib.scg_md_upload(self._h, ARRAY_IDS[name], _lib.ptr(a)))
            self.h2d_bytes += a.nbytes
        elif rows[1] > rows[0]:
            _lib.check(self._lib.scg_md_upload_rows(self._h, ARRAY_IDS[name], _lib.ptr(a), rows[0], rows[1]))
            self.h2d_bytes += (rows[1] - rows[0]) * (a.nbytes // max(a.shape[0], 1))

    def _download(self, name):
        a = self._a[name].ndarray
        _lib.check(self._lib.scg_md_download(self._h, ARRAY_IDS[name], _lib.ptr(a)))
        self.d2h_bytes += a.nbytes
        self._stale.discard(name)

    def _dirty(self, name):
        return name in self._a and name not in self._stale and (name in self._exposed or self._held(name))

    def _push(self, names=_STATE, force=False, ghost_only=()):
        """Before a compute call: upload what the caller may have written (``force``: every host copy that is not
        older than the device's).  ``ghost_only``: arrays whose body rows the call overwrites before reading them."""
        if self._resident and not force:
            return
        for n in names:
            if n in self._a and n not in self._stale and (force or self._dirty(n)):
                self._upload(n, (0, self._ngst) if n in ghost_only else None)
                self._exposed.discard(n)
        # the CE geometry only travels when the caller wrote into it
        for n in _GEOMETRY:
            if n not in self._a:
                continue
            cur = self._a[n].ndarray
            if not np.array_equal(cur, self._geo_seen[n]):
                _lib.check(self._lib.scg_md_upload(self._h, ARRAY_IDS[n], _lib.ptr(cur)))
                self.h2d_bytes += cur.nbytes
                self._geo_seen[n] = cur.copy()

    def _pull(self, names=_STATE, force=False):
        """After a compute call that wrote ``names`` on the device: held arrays are refreshed now, the others when
        they are next accessed.  ``force``: refresh every materialised one now."""
        for n in names:
            if n not in self._a:
                continue
            self._stale.add(n)
            self._exposed.discard(n)
            if force or (not self._resident and self._held(n)):
                self._download(n)

    def upload(self, names=_STATE):
        """Additive API: push the named host arrays to the device now (resident mode: after host-side edits)."""
        for n in names:
            if n in self._a and n not in self._stale:
                self._upload(n)
                self._exposed.discard(n)

    def download(self, names=_STATE):
        """Additive API: bring the named arrays to the host now."""
        for n in names:
            self._get(n)
            if n in self._stale:
                self._download(n)

    def resident(self, flag=True):
        """Keep the field on the device between calls: no implicit transfers until ``resident(False)`` /
        ``sync_to_host()`` (uploads the host copies once now)."""
        if flag and not self._resident:
            self._push(force=True)
        self._resident = bool(flag)
        if not flag:
            self.sync_to_host()
        return self

    def sync_to_host(self):
        for n in list(self._stale):
            if n in self._a:
                self._download(n)

    def _swap_host(self):
        a = self._a
        for c, n in (("so0c", "so0n"), ("so1c", "so1n")):
            a[c], a[n] = a[n], a[c]
            for flags in (self._stale, self._exposed):
                hc, hn = c in flags, n in flags
                flags.discard(c)
                flags.discard(n)
                if hc:
                    flags.add(n)
                if hn:
                    flags.add(c)

    def materialise(self, names=_STATE):
        """Allocate (and fill from the device) the named host arrays now."""
        for n in names:
            self._get(n, expose=False)
        return self

    def _phase(self, phase):
        self._push()
        if self._lib.scg_md_phase(self._h, phase):
            _raise()
        self._pull()

    # -- preparation (wrap_multidim.cpp:275-307) ----------------------------------------------------------
    def prepare_ce(self):
        self._push()
        if self._lib.scg_md_phase(self._h, MD_PREPARE_CE):
            _raise()
        for n in _GEOMETRY:
            if n in self._a:
                a = self._a[n].ndarray
                _lib.check(self._lib.scg_md_download(self._h, ARRAY_IDS[n], _lib.ptr(a)))
                self.d2h_bytes += a.nbytes
                self._geo_seen[n] = a.copy()

    def init_solution(self, gamma, rho, v, p):
        v = list(v)
        if len(v) != self._ndim:
            raise ValueError(f"EulerCore::init_solution: velocity size {len(v)} must equal ndim {self._ndim}")
        vv = (C.c_double * 3)(*(list(map(float, v)) + [0.0] * (3 - len(v))))
        self._push()
        if self._lib.scg_md_init_solution(self._h, float(gamma), float(rho), vv, float(p)):
            _raise(ValueError)
        self._pull()

    # -- march (wrap_multidim.cpp:309-327) -----------------------------------------------------------------
    def march(self, steps):
        if int(steps) < 1:
            self._push()
        else:
            # march(steps >= 1) reads so0n / so1n (they become so0c / so1c at the first update()) and gamma; every
            # body row of so0c / so1c / cflo / cflc is overwritten before it is read (their ghost rows are not all
            # rewritten: unregistered ghosts), and so0t on all rows
            self._push(names=("so0n", "so1n", "gamma", "so0c", "so1c", "cflo", "cflc"),
                       ghost_only=("so0c", "so1c", "cflo", "cflc"))
            self._exposed.discard("so0t")
        if self._lib.scg_md_march(self._h, int(steps)):
            _raise()
        if int(steps) >= 1:
            self._pull(tuple(n for n in _STATE if n != "gamma"))

    def march_substep(self):
        self.materialise(("so0c", "so0n", "so1c", "so1n"))
        self._push()
        if self._lib.scg_md_phase(self._h, MD_SUBSTEP):
            _raise()
        self._swap_host()
        self._pull()

    def update(self):
        # pointer swap on both sides (euler.hpp:123-127): views taken earlier alias the other buffer
        self.materialise(("so0c", "so0n", "so1c", "so1n"))
        self._push(names=("so0c", "so0n", "so1c", "so1n"))
        _lib.check(self._lib.scg_md_phase(self._h, MD_UPDATE))
        self._swap_host()

    def calc_cfl(self):
        self._phase(MD_CFL)

    def calc_solt(self):
        self._phase(MD_SOLT)

    def calc_soln(self):
        self._phase(MD_SOLN)

    def calc_dsoln(self):
        self._phase(MD_DSOLN)

    def bc_soln(self):
        self._phase(MD_BC_SOLN)

    def bc_dsoln(self):
        self._phase(MD_BC_DSOLN)

    # -- boundaries (wrap_multidim.cpp:329-383) ----------------------------------------------------------------
    def _add_bc(self, kind, faces, value=()):
        f = np.ascontiguousarray(faces, dtype=np.int32)
        val = np.ascontiguousarray(value, dtype=np.float64)
        rc = self._lib.scg_md_add_bc(self._h, kind, _lib.ptr(f) if len(f) else None, len(f),
                                     _lib.ptr(val) if len(val) else None, len(val))
        if rc:
            _raise(ValueError)

    def add_nonrefl(self, faces):
        self._add_bc(BC_NONREFL, faces)

    def add_slipwall(self, faces):
        self._add_bc(BC_SLIPWALL, faces)

    def add_inlet(self, faces, value):
        self._add_bc(BC_INLET, faces, value)

    def clear_bc(self):
        _lib.check(self._lib.scg_md_clear_bc(self._h))

    def get_normal_matrix(self, ifc):
        ifc = int(ifc)
        nface = self._dmesh.nface
        if ifc < 0 or ifc >= nface:
            raise IndexError(f"EulerCore: face {ifc} out of range [0, {nface})")
        out = np.zeros((self._ndim, self._ndim))
        _lib.check(self._lib.scg_md_normal_matrix(self._h, ifc, _lib.ptr(out)))
        return out.tolist()

    # -- additive API -----------------------------------------------------------------------------------
    def set_fused(self, mode):
        """march() kernels: 3 (default) cp.async-staged soln+cfl kernel + equation-outer dsoln kernel,
        2 soln+cfl kernel + cooperative dsoln kernel, 1 one fused thread-per-cell kernel, 0 one kernel per
        reference phase (bit-identical results)."""
        _lib.check(self._lib.scg_md_set_fused(self._h, int(mode)))

    def reduce(self, clvol=None):
        """(cflo.min, cflo.max, sum(density*clvol)) over body cells on the device -- the callers' numpy
        reductions of ``solvcon/multidim/euler/_field.py:138-148``.  ``clvol``: the mesh's full table
        (ghost rows first); cached on the device after the first call."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        if not self._resident:
            self._push()
        vol = None if clvol is None else np.ascontiguousarray(clvol, dtype=np.float64)
        _lib.check(self._lib.scg_md_reduce(self._h, _lib.ptr(vol), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    _FIELD_IDS = {"density": 0, "velx": 1, "vely": 2, "velz": 3, "speed": 4, "pressure": 5, "mach": 6,
                  "total_energy": 7, "cfl": 8}

    def derived(self, name):
        """One derived body-cell field (``solvcon/multidim/euler/_field.py:56-126``) evaluated on the device from
        the newest solution, ``[ncell]`` in reference cell order -- 8 B per cell cross the bus instead of the state."""
        if name not in self._FIELD_IDS:
            raise ValueError(f"unknown field '{name}'")
        if not self._resident:
            self._push()
        out = np.empty(self._ncell, dtype=np.float64)
        _lib.check(self._lib.scg_md_derived(self._h, self._FIELD_IDS[name], _lib.ptr(out)))
        return out

    # -- halo exchange of a partition (see solvcon_b200/decomp.py) ------------------------------------
    def halo_set(self, send_cells, recv_cells):
        snd = np.ascontiguousarray(send_cells, dtype=np.int32)
        rcv = np.ascontiguousarray(recv_cells, dtype=np.int32)
        self._nsend, self._nrecv = len(snd), len(rcv)
        _lib.check(self._lib.scg_md_halo_set(self._h, _lib.ptr(snd) if len(snd) else None, len(snd),
                                             _lib.ptr(rcv) if len(rcv) else None, len(rcv)))

    def halo_width(self, what=0):
        return self._neq * (1 + self._ndim) if what == 0 else self._ndim + 1

    def halo_pack(self, what=0):
        """Pack the send cells on the device; returns (device pointer, number of doubles)."""
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self._lib.scg_md_halo_pack(self._h, int(what), C.byref(p), C.byref(n)))
        return p.value, n.value

    def halo_unpack(self, dev_ptr, what=0):
        _lib.check(self._lib.scg_md_halo_unpack(self._h, int(what), C.c_void_p(dev_ptr)))
        self._device_wrote()

    def _device_wrote(self):
        """A resident-mode call changed the state on the device: every materialised host copy is now older."""
        for n in _STATE:
            if n in self._a:
                self._stale.add(n)
                self._exposed.discard(n)

    def substep_local(self):
        """One march_substep of this partition without any host transfer (resident mode only)."""
        if self._lib.scg_md_phase(self._h, MD_SUBSTEP):
            _raise()
        self._device_wrote()

    def substep_interface(self):
        """First stage of a split substep: update(), calc_solt, then the owned INTERFACE cells (``owned == 2``);
        their state may be packed and sent as soon as this returns (stream order)."""
        if self._lib.scg_md_phase(self._h, MD_SUBSTEP_IFACE):
            _raise()
        self._device_wrote()

    def substep_rest(self):
        """Second stage of a split substep: the interior cells and the boundary kernels."""
        if self._lib.scg_md_phase(self._h, MD_SUBSTEP_REST):
            _raise()
        self._device_wrote()

    def algorithmic_bytes(self):
        """Algorithmic bytes one substep moves on this mesh."""
        v = C.c_int64()
        _lib.check(self._lib.scg_md_algorithmic_bytes(self._h, C.byref(v)))
        return v.value
