"""ctypes binding of ``libscg_cuda.so`` -- the C-ABI declared in ``include/scg.h``.

This module is the only place the package touches the shared library.  There is
no CPU fallback: if the library is missing or no CUDA device is present, every
compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int32)
I64P = C.POINTER(C.c_int64)


def lib_path(strict: bool | None = None) -> str:
    """Path of the device library.  ``SCG_STRICT_FP=0`` selects the FMA-contracted speed
    variant (``libscg_cuda_fma.so``); the default is the bit-parity build."""
    if os.environ.get("SCG_LIB"):
        return os.path.join(HERE, os.environ["SCG_LIB"])
    if strict is None:
        strict = os.environ.get("SCG_STRICT_FP", "1") != "0"
    return os.path.join(HERE, "libscg_cuda.so" if strict else "libscg_cuda_fma.so")


class ScgError(RuntimeError):
    pass


_SIGNATURES = {
    # runtime
    "scg_last_error": (C.c_char_p, []),
    "scg_build_is_strict": (C.c_int, []),
    "scg_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "scg_init": (C.c_int, [C.c_int]),
    "scg_shutdown": (C.c_int, []),
    "scg_sync": (C.c_int, []),
    "scg_mem_info": (C.c_int, [I64P, I64P]),
    "scg_launch_count": (C.c_int64, [C.c_int]),
    "scg_timing_dump": (C.c_int, []),
    "scg_stream_handle": (C.c_void_p, []),
    "scg_timer_start": (C.c_int, []),
    "scg_timer_stop": (C.c_int, [DP]),
    "scg_flush_l2": (C.c_int, [C.c_int64]),
    "scg_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "scg_host_free": (C.c_int, [C.c_void_p]),
    "scg_dev_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "scg_dev_free": (C.c_int, [C.c_void_p]),
    "scg_dev_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    # 1D scalar
    "scg_st_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_double]),
    "scg_st_destroy": (C.c_int, [C.c_void_p]),
    "scg_st_set_time_increment": (C.c_int, [C.c_void_p, C.c_double]),
    "scg_st_grid_uniform": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "scg_st_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scg_st_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scg_st_phase": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "scg_st_march": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "scg_st_reduce": (C.c_int, [C.c_void_p, DP, DP, DP]),
    "scg_st_set_halo_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "scg_st_march_half": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "scg_st_halo_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "scg_st_halo_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    # 1D Euler
    "scg_e1d_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_double]),
    "scg_e1d_destroy": (C.c_int, [C.c_void_p]),
    "scg_e1d_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scg_e1d_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scg_e1d_phase": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "scg_e1d_march": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "scg_e1d_reduce": (C.c_int, [C.c_void_p, DP, DP, DP]),
    # host mesh
    "scg_meshgen_mixed3d_count": (C.c_int, [C.c_int64] * 11 + [I64P, I64P]),
    "scg_meshgen_mixed3d_cells": (C.c_int, [C.c_int64] * 11 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, I64P]),
    "scg_hmesh_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.c_int64]),
    "scg_hmesh_destroy": (C.c_int, [C.c_void_p]),
    "scg_hmesh_count": (C.c_int, [C.c_void_p, C.c_int, I64P]),
    "scg_hmesh_table": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), I64P, I64P, I64P,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scg_hmesh_build_interior": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "scg_hmesh_build_edge": (C.c_int, [C.c_void_p]),
    "scg_hmesh_add_bc": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "scg_hmesh_bc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), I64P]),
    "scg_st_march_ring": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int]),
    "scg_comm_unique_id": (C.c_int, [C.c_void_p, C.c_int64, C.c_char_p]),
    "scg_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_char_p]),
    "scg_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scg_comm_finalize": (C.c_int, []),
    "scg_md_halo_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scg_md_halo_exchange": (C.c_int, [C.c_void_p, C.c_int]),
    "scg_md_march_dist": (C.c_int, [C.c_void_p, C.c_int64]),
    "scg_hmesh_bc_rename": (C.c_int, [C.c_void_p, C.c_int64, C.c_char_p]),
    "scg_hmesh_drop_retired": (C.c_int, [C.c_void_p]),
    "scg_hmesh_build_boundary": (C.c_int, [C.c_void_p]),
    "scg_hmesh_build_ghost": (C.c_int, [C.c_void_p]),
    "scg_hmesh_set_device": (C.c_int, [C.c_void_p, C.c_int]),
    "scg_hmesh_gather": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scg_hmesh_on_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "scg_mesh_create_from_hmesh": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p]),
    # multidim
    "scg_mesh_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "scg_mesh_create_part": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "scg_mesh_destroy": (C.c_int, [C.c_void_p]),
    "scg_mesh_perm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "scg_md_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_double]),
    "scg_md_destroy": (C.c_int, [C.c_void_p]),
    "scg_md_set_params": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "scg_md_set_time_increment": (C.c_int, [C.c_void_p, C.c_double]),
    "scg_md_set_fused": (C.c_int, [C.c_void_p, C.c_int]),
    "scg_md_derived": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "scg_md_init_solution": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_double]),
    "scg_md_add_bc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "scg_md_clear_bc": (C.c_int, [C.c_void_p]),
    "scg_md_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "scg_md_upload_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64]),
    "scg_md_download": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "scg_md_phase": (C.c_int, [C.c_void_p, C.c_int]),
    "scg_md_march": (C.c_int, [C.c_void_p, C.c_int64]),
    "scg_md_reduce": (C.c_int, [C.c_void_p, C.c_void_p, DP, DP, DP]),
    "scg_md_algorithmic_bytes": (C.c_int, [C.c_void_p, I64P]),
    "scg_md_normal_matrix": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "scg_selftest_div": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "scg_selftest_rsq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "scg_md_halo_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "scg_md_halo_pack": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), I64P]),
    "scg_md_halo_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
}

_lib = None


def load(path: str | None = None):
    """Load (once) and return the ctypes handle with all signatures set."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or lib_path()
    if not os.path.exists(path):
        raise ScgError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  solvcon_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_names():
    return sorted(_SIGNATURES)


def check(rc: int):
    if rc != 0:
        raise ScgError(load().scg_last_error().decode())


def ptr(a):
    """Raw data pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data


def device_count() -> int:
    n = C.c_int(0)
    rc = load().scg_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def init(device: int = 0):
    check(load().scg_init(device))


def sync():
    check(load().scg_sync())


def launch_count(reset: bool = False) -> int:
    return int(load().scg_launch_count(int(reset)))


class Timer:
    """CUDA-event timer on the library's work stream."""

    def __enter__(self):
        check(load().scg_timer_start())
        self.ms = None
        return self

    def __exit__(self, *exc):
        v = C.c_double(0)
        check(load().scg_timer_stop(C.byref(v)))
        self.ms = v.value
        return False


def pinned_zeros(shape, dtype=np.float64):
    """A zero-filled numpy array in page-locked host memory (``scg_host_alloc``), freed with the array; falls back
    to ordinary memory where no CUDA device is available (CPU-only test runs).  Pinned memory makes the
    upload/download of the API-visible arrays run at the PCIe rate."""
    import weakref
    shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(x) for x in shape)
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    if nbytes < (1 << 16) or os.environ.get("SCG_NO_PINNED"):
        return np.zeros(shape, dtype=dt)
    lib = load()
    p = C.c_void_p()
    if lib.scg_host_alloc(C.byref(p), nbytes) != 0 or not p.value:
        return np.zeros(shape, dtype=dt)
    buf = (C.c_char * nbytes).from_address(p.value)
    a = np.frombuffer(buf, dtype=dt).reshape(shape)
    a[...] = 0
    weakref.finalize(buf, lib.scg_host_free, C.c_void_p(p.value))
    return a
