"""The native connectivity generator of the stepped channel (csrc/scg_meshgen.cpp) equals the numpy definition
(solvcon_b200/meshgen.py: mixed_3d with native=False), array for array: every block mix, sub-lattices, box indices."""
import numpy as np
import pytest

from solvcon_b200 import meshgen


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if isinstance(x, tuple):
            assert all(np.array_equal(u, v) for u, v in zip(x, y))
        else:
            assert x.dtype == y.dtype and np.array_equal(x, y)


@pytest.mark.parametrize("kw", [dict(), dict(box_range=((2, 8), (1, 5), (3, 8))), dict(box_range=((-2, 5), (3, 11), (5, 9))),
                                dict(j_range=(1, 4)), dict(kinds="h"), dict(kinds="t"), dict(kinds="p"), dict(kinds="hp")])
@pytest.mark.parametrize("ret", [dict(), dict(return_box=True), dict(return_box_j=True)])
def test_native_equals_numpy(kw, ret):
    a = meshgen.mixed_3d(9, 6, 8, lx=2.25, ly=1.0, lz=1.6, x1=6, z1=6, jitter=0.1, native=False, **kw, **ret)
    b = meshgen.mixed_3d(9, 6, 8, lx=2.25, ly=1.0, lz=1.6, x1=6, z1=6, jitter=0.1, native=True, **kw, **ret)
    _same(a, b)


def test_native_equals_numpy_bench_shape():
    from solvcon_b200.bench_md import dims_3d
    nx, ny, nz, x1, z1 = dims_3d(200_000)
    h = 1.0 / ny
    a = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=1.0, lz=nz * h, x1=x1, z1=z1, jitter=0.1, native=False)
    b = meshgen.mixed_3d(nx, ny, nz, lx=nx * h, ly=1.0, lz=nz * h, x1=x1, z1=z1, jitter=0.1, native=True)
    _same(a, b)
    assert len(a[1]) > 150_000 and set(np.unique(a[1])) == {5, 6, 7}
