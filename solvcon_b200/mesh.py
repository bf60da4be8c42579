"""Host-side mesh objects for the multidim path.

``TableMesh`` wraps finished StaticMesh tables (ghost rows first, as the reference's SimpleArrays
store them, ``cpp/solvcon/mesh/StaticMesh_boundary.cpp:219-234``) so that any producer of the tables
-- the reference's own ``StaticMesh``, :class:`StaticMesh` below, or a synthetic generator -- can feed
:class:`solvcon_b200.multidim.EulerCore`.
"""
from __future__ import annotations

import numpy as np

from .simplearray import SimpleArray

TABLES = ("ndcrd", "fccnd", "fcnml", "fcara", "clcnd", "clvol", "fctpn", "cltpn", "clgrp",
          "fcnds", "fccls", "clnds", "clfcs", "ednds", "bndfcs")


class TableMesh:
    """A mesh given directly by its tables: ``tables[name] = (array incl. ghost rows, nghost)``."""

    def __init__(self, ndim, tables):
        self.ndim = int(ndim)
        self._t = {}
        for name, (a, g) in tables.items():
            self._t[name] = SimpleArray(array=np.ascontiguousarray(a), nghost=g)
        self.ngstnode = self._t["ndcrd"].nghost
        self.ngstface = self._t["fccls"].nghost
        self.ngstcell = self._t["cltpn"].nghost
        self.nnode = self._t["ndcrd"].nbody
        self.nface = self._t["fccls"].nbody
        self.ncell = self._t["cltpn"].nbody
        self.nbound = len(self._t["bndfcs"].ndarray) if "bndfcs" in self._t else self.ngstcell

    def __getattr__(self, name):
        t = self.__dict__.get("_t", {})
        if name in t:
            return t[name]
        raise AttributeError(name)


# --------------------------------------------------------------------------------------------------
# StaticMesh: native table builder (csrc/scg_hmesh.cpp) behind the reference's binding surface
# --------------------------------------------------------------------------------------------------
import ctypes as _C

from . import _lib as _libmod

(_C_NDIM, _C_NNODE, _C_NFACE, _C_NCELL, _C_NBOUND, _C_NGSTNODE, _C_NGSTFACE, _C_NGSTCELL, _C_NEDGE, _C_NBCS) = range(10)


class StaticMeshBc:
    """Mirror of ``StaticMeshBc`` (``StaticMesh.hpp:146-246``, binding ``wrap_StaticMesh.cpp:33-43``): ``name`` (rw),
    ``nbound``, ``facn`` [n,3]; one Python object per group, like the reference's ``shared_ptr`` holder."""

    NONAME = "<NONAME>"

    def __init__(self, name, facn, mesh=None, index=-1):
        from .simplearray import SimpleArray
        self._name = name
        self._mesh, self._index = mesh, index
        # a SimpleArray like in the reference binding, so that ``bc.facn.ndarray[:, 0]`` (obsrefl _driver.py:489-491)
        # and ``bc.facn[i, 0]`` both work
        self.facn = facn if isinstance(facn, SimpleArray) else SimpleArray(array=np.ascontiguousarray(facn))

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        if self._mesh is not None:
            _libmod.check(self._mesh._lib.scg_hmesh_bc_rename(self._mesh._h, self._index, str(value).encode()))
        self._name = str(value)

    def __len__(self):
        return self.nbound

    @property
    def nbound(self):
        return self.facn.ndarray.shape[0]


class StaticMesh:
    """Host mesh with the reference's surface (``mesh/pymod/wrap_StaticMesh.cpp:61-175``):
    ``StaticMesh(ndim, nnode, nface=0, ncell=0)``; fill ``ndcrd``, ``cltpn``, ``clnds`` through the array
    views; ``build_interior(do_metric, build_edge)``, ``add_bc(name, faces)``, ``build_boundary()``,
    ``build_ghost()``; the 15 tables as SimpleArrays (zero-copy views of the native storage)."""

    FCMND, CLMND, CLMFC, FCREL, BFREL = 4, 8, 6, 4, 3
    NONCELLTYPE, POINT, LINE, QUADRILATERAL, TRIANGLE, HEXAHEDRON, TETRAHEDRON, PRISM, PYRAMID = range(9)

    def __init__(self, ndim, nnode, nface=0, ncell=0):
        self._lib = _libmod.load()
        self._h = _C.c_void_p()
        if self._lib.scg_hmesh_create(_C.byref(self._h), int(ndim), int(nnode), int(nface), int(ncell)):
            raise ValueError(self._lib.scg_last_error().decode())
        self._cache = {}
        self._bcobj = []

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.scg_hmesh_destroy(h)
            self._h = None

    def _count(self, which):
        v = _C.c_int64()
        _libmod.check(self._lib.scg_hmesh_count(self._h, which, _C.byref(v)))
        return int(v.value)

    ndim = property(lambda s: s._count(_C_NDIM))
    nnode = property(lambda s: s._count(_C_NNODE))
    nface = property(lambda s: s._count(_C_NFACE))
    ncell = property(lambda s: s._count(_C_NCELL))
    nbound = property(lambda s: s._count(_C_NBOUND))
    ngstnode = property(lambda s: s._count(_C_NGSTNODE))
    ngstface = property(lambda s: s._count(_C_NGSTFACE))
    ngstcell = property(lambda s: s._count(_C_NGSTCELL))
    nedge = property(lambda s: s._count(_C_NEDGE))
    nbcs = property(lambda s: s._count(_C_NBCS))

    def _table(self, name):
        if name in self._cache:
            return self._cache[name]
        ptr = _C.c_void_p()
        rows, cols, ngh = _C.c_int64(), _C.c_int64(), _C.c_int64()
        real, one_d = _C.c_int(), _C.c_int()
        _libmod.check(self._lib.scg_hmesh_table(self._h, name.encode(), _C.byref(ptr), _C.byref(rows), _C.byref(cols),
                                                _C.byref(ngh), _C.byref(real), _C.byref(one_d)))
        dtype = np.float64 if real.value else np.int32
        shape = (rows.value,) if one_d.value else (rows.value, cols.value)
        n = rows.value * (1 if one_d.value else cols.value)
        if n == 0 or not ptr.value:
            a = np.zeros(shape, dtype=dtype)
        else:
            ctype = _C.c_double if real.value else _C.c_int32
            a = np.ctypeslib.as_array(_C.cast(ptr, _C.POINTER(ctype)), shape=(n,)).reshape(shape)
        sa = SimpleArray(array=a, nghost=int(ngh.value))
        sa._owner = self  # keep the native storage alive as long as the view is
        self._cache[name] = sa
        return sa

    def __getattr__(self, name):
        if name in TABLES:
            return self._table(name)
        raise AttributeError(name)

    def _check(self, rc, exc=ValueError):
        if rc:
            msg = self._lib.scg_last_error().decode()
            if "already built" in msg:
                exc = RuntimeError
            raise exc(msg)
        self._cache = {}  # the native arrays may have been re-allocated

    def set_device(self, mode):
        """Additive API: where the build_* calls run -- ``True`` GPU kernels (``csrc/scg_hmesh_dev.cu``; tables stay on
        the device and come to the host one by one when they are accessed), ``False`` host threads, ``None`` automatic
        (device when one is present and the mesh has >= 2e5 cells).  Same tables either way."""
        _libmod.check(self._lib.scg_hmesh_set_device(self._h, -1 if mode is None else int(bool(mode))))
        return self

    _REAL = ("ndcrd", "fccnd", "fcnml", "fcara", "clcnd", "clvol")
    _COLS = {"fcara": 1, "clvol": 1, "fctpn": 1, "cltpn": 1, "clgrp": 1, "fcnds": 5, "fccls": 4, "clnds": 9, "clfcs": 7}

    def gather(self, name, idx):
        """Additive API: rows ``idx`` (entity indices, negative = ghost) of table ``name`` as an array ``[len(idx), cols]``
        -- gathered on the device when the table was built there, so a few boundary rows of a 1e8-face table do not
        bring the whole table to the host."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        cols = self._COLS.get(name, self.ndim)
        out = np.empty((len(idx), cols), dtype=np.float64 if name in self._REAL else np.int32)
        if self._lib.scg_hmesh_gather(self._h, name.encode(), _libmod.ptr(idx) if len(idx) else None, len(idx), _libmod.ptr(out)):
            raise IndexError(self._lib.scg_last_error().decode())
        return out

    @property
    def on_device(self):
        """0 = host tables, 1 = interior tables live on the device, 2 = + ghost entities and no table has been handed out
        as a (writable) host view since: a complete, current device mesh."""
        v = _C.c_int()
        _libmod.check(self._lib.scg_hmesh_on_device(self._h, _C.byref(v)))
        return int(v.value)

    def build_interior(self, do_metric=True, build_edge=True, _do_metric=None):
        self._check(self._lib.scg_hmesh_build_interior(self._h, int(bool(do_metric)), int(bool(build_edge))))

    def build_edge(self):
        self._check(self._lib.scg_hmesh_build_edge(self._h))

    def build_boundary(self):
        self._check(self._lib.scg_hmesh_build_boundary(self._h))

    def build_ghost(self):
        self._check(self._lib.scg_hmesh_build_ghost(self._h))

    def add_bc(self, name, faces):
        f = np.ascontiguousarray(faces, dtype=np.int32)
        self._check(self._lib.scg_hmesh_add_bc(self._h, str(name).encode(), _libmod.ptr(f) if len(f) else None, len(f)))
        return self.bc(self.nbcs - 1)

    def bc(self, key):
        if isinstance(key, str):
            for i in range(self.nbcs):
                b = self.bc(i)
                if b.name == key:
                    return b
            raise KeyError(key)
        key = int(key)
        name, facn, n = _C.c_char_p(), _C.c_void_p(), _C.c_int64()
        if self._lib.scg_hmesh_bc(self._h, key, _C.byref(name), _C.byref(facn), _C.byref(n)):
            raise IndexError(self._lib.scg_last_error().decode())
        # one Python object per group (``assertIs(bottom, mh.bc(0))``, tests/test_mesh.py:247-253); the facn rows are a
        # view of the group's native storage, which is never re-made
        while len(self._bcobj) <= key:
            self._bcobj.append(None)
        if self._bcobj[key] is None:
            if n.value:
                a = np.ctypeslib.as_array(_C.cast(facn, _C.POINTER(_C.c_int32)), shape=(n.value * 3,)).reshape(n.value, 3)
            else:
                a = np.zeros((0, 3), dtype=np.int32)
            b = StaticMeshBc(name.value.decode(), a, self, key)
            b.facn._owner = self
            self._bcobj[key] = b
        return self._bcobj[key]

    @property
    def bcs(self):
        return [self.bc(i) for i in range(self.nbcs)]

    def tables(self):
        """Deep copies of all tables: ``{name: (array incl. ghost rows, nghost)}``."""
        return {n: (self._table(n).ndarray.copy(), self._table(n).nghost) for n in TABLES}


def make_mesh(ndim, ndcrd, cltpn, clnds, bcs=(), build_edge=False, device=None):
    """Nodes + cells -> a fully built StaticMesh (interior, boundary, ghost).  ``device``: see ``StaticMesh.set_device``."""
    ndcrd = np.asarray(ndcrd, dtype=np.float64)
    m = StaticMesh(ndim, len(ndcrd), 0, len(cltpn))
    m.set_device(device)
    m.ndcrd.ndarray[...] = ndcrd
    m.cltpn.ndarray[...] = cltpn
    m.clnds.ndarray[...] = clnds
    m.build_interior(True, build_edge)
    for name, faces in bcs:
        m.add_bc(name, faces)
    m.build_boundary()
    m.build_ghost()
    # no view of an earlier table generation has left this function: free the generations retired by the builds
    _libmod.check(m._lib.scg_hmesh_drop_retired(m._h))
    return m
