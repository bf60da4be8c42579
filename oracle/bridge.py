"""ctypes bridge to the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``ref``  : oracle/_ref/libsolvcon_ref.so -- the UNMODIFIED reference sources
             compiled in place (oracle/Makefile, oracle/ref_shim.cpp).
* ``port`` : oracle/liboracle_cese.so -- our plain-C restatement
             (oracle/cese_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``solvcon_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(HERE, "_ref", "libsolvcon_ref.so")
PORT_PATH = os.path.join(HERE, "liboracle_cese.so")

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int32)

_ref = None
_port = None


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def have_port() -> bool:
    return os.path.exists(PORT_PATH)


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    if a.dtype == np.float64:
        return a.ctypes.data_as(DP)
    if a.dtype == np.int32:
        return a.ctypes.data_as(IP)
    raise TypeError(a.dtype)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_PATH)
        lib.ref_last_error.restype = C.c_char_p
        for f in ("ref_e1d_coord", "ref_e1d_cfl", "ref_e1d_so0", "ref_e1d_so1", "ref_e1d_gamma",
                  "ref_st_x", "ref_st_so0", "ref_st_so1", "ref_st_cfl"):
            getattr(lib, f).restype = DP
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.ref_e1d_create.restype = C.c_void_p
        lib.ref_e1d_create.argtypes = [C.c_int64, C.c_double]
        lib.ref_e1d_destroy.argtypes = [C.c_void_p]
        lib.ref_e1d_phase.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.ref_e1d_march.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        lib.ref_e1d_derived.argtypes = [C.c_void_p, C.c_int, DP]
        lib.ref_st_create.restype = C.c_void_p
        lib.ref_st_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int64, C.c_double]
        lib.ref_st_create_xloc.restype = C.c_void_p
        lib.ref_st_create_xloc.argtypes = [C.c_int, DP, C.c_int64, C.c_double]
        lib.ref_st_destroy.argtypes = [C.c_void_p]
        lib.ref_st_nx.restype = C.c_int64
        lib.ref_st_nx.argtypes = [C.c_void_p]
        lib.ref_st_phase.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.ref_st_march.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        lib.ref_mesh_create.restype = C.c_void_p
        lib.ref_mesh_create.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64]
        lib.ref_mesh_destroy.argtypes = [C.c_void_p]
        lib.ref_mesh_build_interior.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ref_mesh_add_bc.argtypes = [C.c_void_p, C.c_char_p, IP, C.c_int64]
        lib.ref_mesh_build_boundary.argtypes = [C.c_void_p]
        lib.ref_mesh_build_ghost.argtypes = [C.c_void_p]
        lib.ref_mesh_count.restype = C.c_int64
        lib.ref_mesh_count.argtypes = [C.c_void_p, C.c_int]
        lib.ref_mesh_table.restype = C.c_void_p
        lib.ref_mesh_table.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        lib.ref_md_create.restype = C.c_void_p
        lib.ref_md_create.argtypes = [C.c_void_p, C.c_double]
        lib.ref_md_destroy.argtypes = [C.c_void_p]
        lib.ref_md_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        lib.ref_md_init_solution.argtypes = [C.c_void_p, C.c_double, C.c_double, DP, C.c_double]
        lib.ref_md_add_bc.argtypes = [C.c_void_p, C.c_int, IP, C.c_int64, DP, C.c_int]
        lib.ref_md_clear_bc.argtypes = [C.c_void_p]
        lib.ref_md_phase.argtypes = [C.c_void_p, C.c_int]
        lib.ref_md_march.argtypes = [C.c_void_p, C.c_int64]
        lib.ref_md_array.restype = DP
        lib.ref_md_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.ref_md_normal_matrix.argtypes = [C.c_void_p, C.c_int32, DP]
        _ref = lib
    return _ref


class OrcMeshStruct(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("ncell", C.c_int32), ("ngstcell", C.c_int32), ("nface", C.c_int32),
                ("cltpn", C.c_void_p), ("clfcs", C.c_void_p), ("fccls", C.c_void_p), ("fcnds", C.c_void_p),
                ("ndcrd", C.c_void_p), ("fccnd", C.c_void_p), ("fcnml", C.c_void_p), ("clcnd", C.c_void_p)]


class OrcMdStruct(C.Structure):
    _fields_ = [("mesh", C.POINTER(OrcMeshStruct)), ("dt", C.c_double),
                ("sigma0", C.c_double), ("taumin", C.c_double), ("tauscale", C.c_double),
                ("cevol", C.c_void_p), ("cecnd", C.c_void_p), ("sfcnd", C.c_void_p), ("sfnml", C.c_void_p),
                ("so0c", C.c_void_p), ("so0n", C.c_void_p), ("so0t", C.c_void_p),
                ("so1c", C.c_void_p), ("so1n", C.c_void_p),
                ("cflo", C.c_void_p), ("cflc", C.c_void_p), ("gamma", C.c_void_p),
                ("nbnd", C.c_int32), ("bnd_kind", C.c_void_p), ("bnd_nface", C.c_void_p),
                ("bnd_faces", C.c_void_p), ("bnd_value", C.c_void_p)]


def port_lib():
    global _port
    if _port is None:
        lib = C.CDLL(PORT_PATH)
        lib.orc_grid_fill.argtypes = [C.c_double, C.c_double, C.c_int64, DP]
        lib.orc_grid_fill_xloc.argtypes = [DP, C.c_int64, DP]
        lib.orc_st_phase.argtypes = [C.c_int, C.c_int64, DP, DP, DP, DP, C.c_double, C.c_int, C.c_int, C.c_int]
        lib.orc_st_march.argtypes = [C.c_int, C.c_int64, DP, DP, DP, DP, C.c_double, C.c_int, C.c_int64]
        lib.orc_e1d_phase.argtypes = [C.c_int64, DP, DP, DP, DP, DP, C.c_double, C.c_int, C.c_int, C.c_int]
        lib.orc_e1d_march.argtypes = [C.c_int64, DP, DP, DP, DP, DP, C.c_double, C.c_int, C.c_int64]
        lib.orc_md_phase.argtypes = [C.POINTER(OrcMdStruct), C.c_int]
        lib.orc_md_march.argtypes = [C.POINTER(OrcMdStruct), C.c_int64]
        lib.orc_normal_matrix.argtypes = [C.c_int, DP, DP]
        _port = lib
    return _port


def _check(rc, lib):
    if rc != 0:
        raise RuntimeError(lib.ref_last_error().decode())


# --------------------------------------------------------------------- 1D ---

PH_SO0, PH_CFL, PH_SO1, PH_BC0, PH_BC1 = range(5)


class RefEuler1D:
    """Reference ``onedim.Euler1DCore`` (wrap_onedim.cpp:38-150) through the shim."""

    def __init__(self, ncoord, dt):
        self.lib = ref_lib()
        self.h = self.lib.ref_e1d_create(ncoord, dt)
        if not self.h:
            raise ValueError(self.lib.ref_last_error().decode())
        self.ncoord, self.dt = ncoord, dt
        as_array = np.ctypeslib.as_array
        self.coord = as_array(self.lib.ref_e1d_coord(self.h), shape=(ncoord,))
        self.cfl = as_array(self.lib.ref_e1d_cfl(self.h), shape=(ncoord,))
        self.gamma = as_array(self.lib.ref_e1d_gamma(self.h), shape=(ncoord,))
        self.so0 = as_array(self.lib.ref_e1d_so0(self.h), shape=(ncoord, 3))
        self.so1 = as_array(self.lib.ref_e1d_so1(self.h), shape=(ncoord, 3))

    def phase(self, phase, odd=False, alpha=2):
        _check(self.lib.ref_e1d_phase(self.h, phase, int(odd), alpha), self.lib)

    def march(self, alpha, steps):
        _check(self.lib.ref_e1d_march(self.h, alpha, steps), self.lib)

    E1D_FIELDS = ("density", "velocity", "pressure", "temperature", "internal_energy", "entropy")

    def derived(self, name):
        """The reference's own derived-field array (Euler1DCore.cpp:37-205)."""
        out = np.empty(self.ncoord)
        _check(self.lib.ref_e1d_derived(self.h, self.E1D_FIELDS.index(name), _p(out)), self.lib)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_e1d_destroy(self.h)
            self.h = None


class PortEuler1D:
    """Our C restatement with the same surface as :class:`RefEuler1D`."""

    def __init__(self, ncoord, dt):
        if ncoord % 2 == 0:
            raise ValueError("ncoord cannot be even")
        self.lib = port_lib()
        self.ncoord, self.dt = ncoord, dt
        self.coord = np.zeros(ncoord)
        self.cfl = np.zeros(ncoord)
        self.gamma = np.full(ncoord, 1.4)
        self.so0 = np.zeros((ncoord, 3))
        self.so1 = np.zeros((ncoord, 3))

    def _args(self):
        return (self.ncoord, _p(self.coord), _p(self.gamma), _p(self.so0), _p(self.so1), _p(self.cfl), self.dt)

    def phase(self, phase, odd=False, alpha=2):
        assert 0 == self.lib.orc_e1d_phase(*self._args(), phase, int(odd), alpha)

    def march(self, alpha, steps):
        assert 0 == self.lib.orc_e1d_march(*self._args(), alpha, steps)


class RefScalar1D:
    """Reference ``spacetime.{LinearScalar,InviscidBurgers}Solver`` through the shim."""

    def __init__(self, eq, xmin=None, xmax=None, ncelm=None, dt=0.0, xloc=None):
        self.lib = ref_lib()
        if xloc is not None:
            xloc = np.ascontiguousarray(xloc, dtype=np.float64)
            self.h = self.lib.ref_st_create_xloc(eq, _p(xloc), len(xloc), dt)
        else:
            self.h = self.lib.ref_st_create(eq, xmin, xmax, ncelm, dt)
        if not self.h:
            raise ValueError(self.lib.ref_last_error().decode())
        self.eq, self.dt = eq, dt
        nx = self.lib.ref_st_nx(self.h)
        self.nx = nx
        self.ncelm = (nx - 5) // 2
        as_array = np.ctypeslib.as_array
        self.x = as_array(self.lib.ref_st_x(self.h), shape=(nx,))
        self.so0 = as_array(self.lib.ref_st_so0(self.h), shape=(nx,))
        self.so1 = as_array(self.lib.ref_st_so1(self.h), shape=(nx,))
        self.cfl = as_array(self.lib.ref_st_cfl(self.h), shape=(nx,))

    def phase(self, phase, odd=False, alpha=2):
        _check(self.lib.ref_st_phase(self.h, phase, int(odd), alpha), self.lib)

    def march(self, alpha, steps):
        _check(self.lib.ref_st_march(self.h, alpha, steps), self.lib)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_st_destroy(self.h)
            self.h = None


class PortScalar1D:
    def __init__(self, eq, xmin=None, xmax=None, ncelm=None, dt=0.0, xloc=None):
        self.lib = port_lib()
        self.eq, self.dt = eq, dt
        if xloc is not None:
            xloc = np.ascontiguousarray(xloc, dtype=np.float64)
            ncelm = len(xloc) - 1
            self.x = np.zeros(2 * ncelm + 5)
            if 0 != self.lib.orc_grid_fill_xloc(_p(xloc), len(xloc), _p(self.x)):
                raise ValueError("bad xloc")
        else:
            self.x = np.zeros(2 * ncelm + 5)
            if 0 != self.lib.orc_grid_fill(xmin, xmax, ncelm, _p(self.x)):
                raise ValueError("bad grid")
        self.ncelm = ncelm
        self.nx = 2 * ncelm + 5
        self.so0 = np.zeros(self.nx)
        self.so1 = np.zeros(self.nx)
        self.cfl = np.zeros(self.nx)

    def _args(self):
        return (self.eq, self.ncelm, _p(self.x), _p(self.so0), _p(self.so1), _p(self.cfl), self.dt)

    def phase(self, phase, odd=False, alpha=2):
        assert 0 == self.lib.orc_st_phase(*self._args(), phase, int(odd), alpha)

    def march(self, alpha, steps):
        assert 0 == self.lib.orc_st_march(*self._args(), alpha, steps)


# ------------------------------------------------------------------- mesh ---

MESH_TABLES = ("ndcrd", "fccnd", "fcnml", "fcara", "clcnd", "clvol", "fctpn", "cltpn", "clgrp",
               "fcnds", "fccls", "clnds", "clfcs", "ednds", "bndfcs")
(C_NDIM, C_NNODE, C_NFACE, C_NCELL, C_NBOUND, C_NGSTNODE, C_NGSTFACE, C_NGSTCELL, C_NEDGE) = range(9)


class RefMesh:
    """Reference ``StaticMesh`` (wrap_StaticMesh.cpp:61-175) through the shim."""

    def __init__(self, ndim, nnode, nface=0, ncell=0):
        self.lib = ref_lib()
        self.h = self.lib.ref_mesh_create(ndim, nnode, nface, ncell)
        if not self.h:
            raise ValueError(self.lib.ref_last_error().decode())
        self.ndim = ndim

    def count(self, which):
        return int(self.lib.ref_mesh_count(self.h, which))

    ncell = property(lambda s: s.count(C_NCELL))
    nface = property(lambda s: s.count(C_NFACE))
    nnode = property(lambda s: s.count(C_NNODE))
    nbound = property(lambda s: s.count(C_NBOUND))
    ngstcell = property(lambda s: s.count(C_NGSTCELL))
    ngstface = property(lambda s: s.count(C_NGSTFACE))
    ngstnode = property(lambda s: s.count(C_NGSTNODE))

    def table(self, name):
        """Zero-copy view of the full table (ghost rows first); returns (array, nghost)."""
        rows, cols, ngh, isreal = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        ptr = self.lib.ref_mesh_table(self.h, name.encode(), C.byref(rows), C.byref(cols), C.byref(ngh),
                                      C.byref(isreal))
        if rows.value == 0:
            dt = np.float64 if isreal.value else np.int32
            return np.zeros((0, cols.value), dtype=dt), 0
        if not ptr:
            raise KeyError(name)
        ctype = C.c_double if isreal.value else C.c_int32
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(rows.value * cols.value,))
        one_d = name in ("fcara", "clvol", "fctpn", "cltpn", "clgrp")
        a = a.reshape((rows.value,) if one_d else (rows.value, cols.value))
        return a, int(ngh.value)

    def body(self, name):
        a, ngh = self.table(name)
        return a[ngh:]

    def build_interior(self, do_metric=True, do_edge=True):
        _check(self.lib.ref_mesh_build_interior(self.h, int(do_metric), int(do_edge)), self.lib)

    def add_bc(self, name, faces):
        f = np.ascontiguousarray(faces, dtype=np.int32)
        _check(self.lib.ref_mesh_add_bc(self.h, name.encode(), _p(f), len(f)), self.lib)

    def build_boundary(self):
        _check(self.lib.ref_mesh_build_boundary(self.h), self.lib)

    def build_ghost(self):
        _check(self.lib.ref_mesh_build_ghost(self.h), self.lib)

    def tables(self):
        """Deep copies of all tables: {name: (array incl. ghost rows, nghost)}."""
        out = {}
        for n in MESH_TABLES:
            a, g = self.table(n)
            out[n] = (a.copy(), g)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_mesh_destroy(self.h)
            self.h = None


def ref_gmsh_mesh(data: bytes):
    """``Gmsh(data).to_block()`` of the reference (cpp/solvcon/inout/gmsh.cpp) as a RefMesh."""
    lib = ref_lib()
    lib.ref_gmsh_to_block.restype = C.c_void_p
    lib.ref_gmsh_to_block.argtypes = [C.c_char_p, C.c_int64]
    h = lib.ref_gmsh_to_block(data, len(data))
    if not h:
        raise ValueError(lib.ref_last_error().decode())
    m = RefMesh.__new__(RefMesh)
    m.lib, m.h = lib, h
    m.ndim = int(lib.ref_mesh_count(h, C_NDIM))
    return m


def make_ref_mesh(ndim, ndcrd, cltpn, clnds, bcs=()):
    """Build a reference StaticMesh from nodes+cells the way the reference's tests do
    (tests/test_multidim.py:152-281): fill ndcrd/cltpn/clnds, build_interior,
    add_bc*, build_boundary, build_ghost."""
    ndcrd = np.asarray(ndcrd, dtype=np.float64)
    cltpn = np.asarray(cltpn, dtype=np.int32)
    clnds = np.asarray(clnds, dtype=np.int32)
    m = RefMesh(ndim, len(ndcrd), 0, len(cltpn))
    m.body("ndcrd")[...] = ndcrd
    m.body("cltpn")[...] = cltpn
    m.body("clnds")[...] = clnds
    m.build_interior(True, True)
    for name, faces in bcs:
        m.add_bc(name, faces)
    m.build_boundary()
    m.build_ghost()
    return m


# --------------------------------------------------------------- multidim ---

(MD_UPDATE, MD_SOLT, MD_SOLN, MD_BC_SOLN, MD_CFL, MD_DSOLN, MD_BC_DSOLN, MD_PREPARE_CE, MD_SUBSTEP) = range(9)
MD_ARRAYS = ("cevol", "cecnd", "sfcnd", "sfnml", "so0c", "so0n", "so0t", "so1c", "so1n", "cflo", "cflc", "gamma")
BC_NONREFL, BC_SLIPWALL, BC_INLET = 0, 1, 2


def md_shapes(ndim, ncell, ngst):
    T = ncell + ngst
    neq = ndim + 2
    return {
        "cevol": ((T, 7), ngst), "cecnd": ((T, 7 * ndim), ngst),
        "sfcnd": ((ncell, 24, ndim), 0), "sfnml": ((ncell, 24, ndim), 0),
        "so0c": ((T, neq), ngst), "so0n": ((T, neq), ngst), "so0t": ((T, neq), ngst),
        "so1c": ((T, neq, ndim), ngst), "so1n": ((T, neq, ndim), ngst), "stm": ((T, neq), ngst),
        "cflo": ((T,), ngst), "cflc": ((T,), ngst), "gamma": ((T,), ngst),
    }


class RefEuler:
    """Reference ``EulerCore`` (wrap_multidim.cpp:247-408) through the shim."""

    def __init__(self, mesh: RefMesh, dt):
        self.lib = ref_lib()
        self.mesh = mesh
        self.h = self.lib.ref_md_create(mesh.h, dt)
        if not self.h:
            raise ValueError(self.lib.ref_last_error().decode())
        self.ndim, self.ncell, self.ngst = mesh.ndim, mesh.ncell, mesh.ngstcell
        self.neq = self.ndim + 2
        self.shapes = md_shapes(self.ndim, self.ncell, self.ngst)

    def arr(self, name):
        """Zero-copy view incl. ghost rows (re-query after update(): buffers swap)."""
        size, ngh = C.c_int64(), C.c_int64()
        ptr = self.lib.ref_md_array(self.h, name.encode(), C.byref(size), C.byref(ngh))
        if not ptr:
            raise KeyError(name)
        shape, g = self.shapes[name]
        a = np.ctypeslib.as_array(ptr, shape=(size.value,)).reshape(shape)
        assert g == ngh.value
        return a

    def body(self, name):
        return self.arr(name)[self.shapes[name][1]:]

    def set_params(self, sigma0, taumin, tauscale):
        self.lib.ref_md_set_params(self.h, sigma0, taumin, tauscale)

    def init_solution(self, gamma, rho, v, p):
        vv = np.zeros(3)
        vv[:len(v)] = v
        rc = self.lib.ref_md_init_solution(self.h, gamma, rho, _p(vv), p)
        if rc:
            raise ValueError(self.lib.ref_last_error().decode())

    def add_bc(self, kind, faces, value=()):
        f = np.ascontiguousarray(faces, dtype=np.int32)
        v = np.ascontiguousarray(value, dtype=np.float64)
        rc = self.lib.ref_md_add_bc(self.h, kind, _p(f), len(f), _p(v) if len(v) else None, len(v))
        if rc:
            raise ValueError(self.lib.ref_last_error().decode())

    def clear_bc(self):
        self.lib.ref_md_clear_bc(self.h)

    def phase(self, phase):
        _check(self.lib.ref_md_phase(self.h, phase), self.lib)

    def march(self, steps):
        _check(self.lib.ref_md_march(self.h, steps), self.lib)

    def normal_matrix(self, ifc):
        out = np.zeros((self.ndim, self.ndim))
        _check(self.lib.ref_md_normal_matrix(self.h, ifc, _p(out)), self.lib)
        return out

    def gradient_element(self, icl, tau, udf):
        """The reference's GradientElement(mesh, cecnd, icl, tau) flattened: dict of meta / rcl / idis / jdis / faces
        (1-based) / dmat / grads (= solve_gradient(ifge, udf)) / nfge_inverse."""
        lib = self.lib
        lib.ref_md_gradient_element.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        meta, rcl = np.zeros(4, np.int32), np.zeros(6, np.int32)
        idis, jdis = np.zeros(19), np.zeros(18)
        faces, dmat, grads = np.zeros(24, np.int32), np.zeros(72), np.zeros(24)
        u = np.zeros(3)
        u[:len(udf)] = udf
        _check(lib.ref_md_gradient_element(self.h, self.mesh.h, icl, tau, _p(u), _p(meta), _p(rcl), _p(idis), _p(jdis),
                                           _p(faces), _p(dmat), _p(grads)), lib)
        return {"meta": meta, "rcl": rcl, "idis": idis[:18].reshape(6, 3), "jdis": jdis.reshape(6, 3),
                "faces": faces.reshape(8, 3), "dmat": dmat.reshape(8, 3, 3), "grads": grads.reshape(8, 3),
                "nfge_inverse": float(idis[18])}

    def snapshot(self):
        return {n: self.arr(n).copy() for n in MD_ARRAYS}

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_md_destroy(self.h)
            self.h = None


class PortEuler:
    """Our C restatement of EulerCore over plain numpy tables (``tables`` as returned by
    :meth:`RefMesh.tables` or by the product's own mesh builder)."""

    def __init__(self, tables, ndim, dt):
        self.lib = port_lib()
        self.ndim, self.dt = ndim, dt
        self.neq = ndim + 2
        self.t = {k: (np.ascontiguousarray(a), g) for k, (a, g) in tables.items()}
        clt, g = self.t["cltpn"]
        self.ngst = g
        self.ncell = len(clt) - g
        self.nface = len(self.t["fccls"][0]) - self.t["fccls"][1]
        self.shapes = md_shapes(ndim, self.ncell, self.ngst)
        self.a = {n: np.zeros(s) for n, (s, _) in self.shapes.items()}
        self.sigma0, self.taumin, self.tauscale = 3.0, 0.0, 1.0
        self.bcs = []
        self.cur = ("so0c", "so0n", "so1c", "so1n")  # names after swaps

    def _body_ptr(self, a, nghost):
        return a.ctypes.data + nghost * (a.strides[0])

    def arr(self, name):
        return self.a[name]

    def body(self, name):
        return self.a[name][self.shapes[name][1]:]

    def set_params(self, sigma0, taumin, tauscale):
        self.sigma0, self.taumin, self.tauscale = sigma0, taumin, tauscale

    def init_solution(self, gamma, rho, v, p):
        # EulerCore::init_solution, euler_solution.cpp:54-99
        vsq = 0.0
        for d in range(self.ndim):
            vsq += v[d] * v[d]
        energy = p / (gamma - 1.0) + 0.5 * rho * vsq
        self.a["gamma"][...] = gamma
        b = self.body("so0n")
        b[:, 0] = rho
        for d in range(self.ndim):
            b[:, 1 + d] = rho * v[d]
        b[:, self.neq - 1] = energy

    def add_bc(self, kind, faces, value=()):
        self.bcs.append((kind, np.ascontiguousarray(faces, dtype=np.int32),
                         np.ascontiguousarray(value, dtype=np.float64)))

    def clear_bc(self):
        self.bcs = []

    def _structs(self):
        ms = OrcMeshStruct()
        ms.ndim, ms.ncell, ms.ngstcell, ms.nface = self.ndim, self.ncell, self.ngst, self.nface
        for n in ("cltpn", "clfcs", "fccls", "fcnds", "ndcrd", "fccnd", "fcnml", "clcnd"):
            a, g = self.t[n]
            setattr(ms, n, self._body_ptr(a, g))
        md = OrcMdStruct()
        md.mesh = C.pointer(ms)
        md.dt, md.sigma0, md.taumin, md.tauscale = self.dt, self.sigma0, self.taumin, self.tauscale
        for n in MD_ARRAYS:
            setattr(md, n, self._body_ptr(self.a[n], self.shapes[n][1]))
        nb = len(self.bcs)
        kinds = np.array([b[0] for b in self.bcs], dtype=np.int32)
        nfaces = np.array([len(b[1]) for b in self.bcs], dtype=np.int32)
        fptrs = (C.c_void_p * max(nb, 1))(*[b[1].ctypes.data for b in self.bcs])
        vptrs = (C.c_void_p * max(nb, 1))(*[b[2].ctypes.data if len(b[2]) else None for b in self.bcs])
        md.nbnd = nb
        md.bnd_kind = kinds.ctypes.data
        md.bnd_nface = nfaces.ctypes.data
        md.bnd_faces = C.cast(fptrs, C.c_void_p)
        md.bnd_value = C.cast(vptrs, C.c_void_p)
        keep = (ms, kinds, nfaces, fptrs, vptrs)
        return md, keep

    def _after(self, md):
        # the C side swaps pointers on update(); mirror that in the name->array map
        c0 = self._body_ptr(self.a["so0c"], self.ngst)
        if md.so0c != c0:
            self.a["so0c"], self.a["so0n"] = self.a["so0n"], self.a["so0c"]
            self.a["so1c"], self.a["so1n"] = self.a["so1n"], self.a["so1c"]

    def phase(self, phase):
        md, keep = self._structs()
        assert 0 == self.lib.orc_md_phase(C.byref(md), phase)
        self._after(md)

    def march(self, steps):
        md, keep = self._structs()
        assert 0 == self.lib.orc_md_march(C.byref(md), steps)
        self._after(md)

    def snapshot(self):
        return {n: self.a[n].copy() for n in MD_ARRAYS}
