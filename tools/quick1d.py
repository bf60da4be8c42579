"""Quick device-resident throughput probe of the 1D kernels (not the bench contract)."""
import ctypes as C
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solvcon_b200 import _lib
from solvcon_b200 import spacetime as st
from solvcon_b200.onedim import Euler1DCore

_lib.init(0)
lib = _lib.load()
peak = 6549.8e9
for eq, name in ((0, "linear"), (1, "burgers")):
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    h = C.c_void_p()
    _lib.check(lib.scg_st_create(C.byref(h), eq, n, 0.5 * 8 * np.pi / n))
    _lib.check(lib.scg_st_grid_uniform(h, 0.0, 8 * np.pi))
    nx = 2 * n + 5
    x = np.zeros(nx); so0 = np.zeros(nx); so1 = np.zeros(nx)
    _lib.check(lib.scg_st_download(h, _lib.ptr(x), None, None, None))
    so0[2::2] = 0.5 * np.sin(x[2::2]); so1[2::2] = 0.5 * np.cos(x[2::2])
    _lib.check(lib.scg_st_upload(h, None, _lib.ptr(so0), _lib.ptr(so1), None))
    _lib.check(lib.scg_st_march(h, 2, 3))
    _lib.sync()
    steps = 10
    with _lib.Timer() as t:
        _lib.check(lib.scg_st_march(h, 2, steps))
    upd = (2 * n + 1) * steps
    print(f"1D {name}: n={n} {upd / t.ms * 1e3 / 1e9:.2f} G upd/s, {upd * 56 / t.ms * 1e3 / 1e9:.0f} GB/s alg, frac {upd * 56 / t.ms * 1e3 / peak:.3f}")
    lib.scg_st_destroy(h)

N = (int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000) | 1
h = C.c_void_p()
_lib.check(lib.scg_e1d_create(C.byref(h), N, 0.8 * 2.0 / N))
coord = np.linspace(-1, 1, N); so0 = np.zeros((N, 3)); so1 = np.zeros((N, 3)); ga = np.full(N, 1.4)
so0[:, 0] = np.where(coord < 0, 1.0, 0.125); so0[:, 2] = np.where(coord < 0, 2.5, 0.25)
_lib.check(lib.scg_e1d_upload(h, _lib.ptr(coord), _lib.ptr(ga), _lib.ptr(so0), _lib.ptr(so1), None))
_lib.check(lib.scg_e1d_march(h, 2, 3)); _lib.sync()
steps = 10
with _lib.Timer() as t:
    _lib.check(lib.scg_e1d_march(h, 2, steps))
upd = (N - 3) * steps
print(f"1D euler: N={N} {upd / t.ms * 1e3 / 1e9:.2f} G upd/s, {upd * 128 / t.ms * 1e3 / 1e9:.0f} GB/s alg, frac {upd * 128 / t.ms * 1e3 / peak:.3f}")
