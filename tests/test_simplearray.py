"""``SimpleArray`` mirror (solvcon_b200/simplearray.py): the subset of the reference contract the CESE front ends and
the reference's own tests use (wrap_SimpleArray.hpp:65-141,193-214, array_common.hpp:171-229,
SimpleArray.hpp:2241-2270,2480-2512; used at _field.py:47-54, test_multidim.py:489-505, test_euler_multidim.py:15-19).
CPU only."""
import numpy as np
import pytest

from solvcon_b200.simplearray import SimpleArray, SimpleArrayFloat64, SimpleArrayInt32


def test_construction_views_and_metadata():
    a = SimpleArrayFloat64((5, 3), 1.5)
    assert a.shape == (5, 3) and a.stride == (3, 1) and a.size == 15 and a.nbytes == 120 and a.itemsize == 8
    assert a.ndarray.base is None or a.ndarray.base is not None        # a view of the live buffer, not a copy:
    a.ndarray[2, 1] = -4.0
    assert a[2, 1] == -4.0
    src = np.arange(6, dtype=np.int32).reshape(3, 2)
    b = SimpleArrayInt32(array=src)
    assert b.ndarray is src and b.shape == (3, 2)                        # wraps without copying
    np.testing.assert_almost_equal(a, a.ndarray)                          # accepted by numpy.testing directly
    c = a.clone()
    c.fill(0.0)
    assert a[0, 0] == 1.5 and c[0, 0] == 0.0


def test_ghost_shifted_indexing():
    a = SimpleArray(array=np.arange(12.0).reshape(6, 2))
    a.nghost = 2
    assert a.has_ghost and a.nghost == 2 and a.nbody == 4
    # row -1 is the first ghost row, i.e. the stored row nghost - 1; row 0 the first body row
    assert a[-1, 0] == a.ndarray[1, 0] and a[-2, 1] == a.ndarray[0, 1] and a[0, 0] == a.ndarray[2, 0]
    assert a[3, 1] == a.ndarray[5, 1]
    a[-1, 1] = 99.0
    assert a.ndarray[1, 1] == 99.0
    for bad in (-3, 4):
        with pytest.raises(IndexError, match="out of bounds"):
            a[bad, 0]
        with pytest.raises(IndexError):
            a[bad, 0] = 1.0
    assert np.array_equal(a.body, a.ndarray[2:]) and np.array_equal(a.ghost[0], a.ndarray[1])
    v = SimpleArray(array=np.arange(5.0))
    v.nghost = 1
    assert v[-1] == 0.0 and v[0] == 1.0 and v[3] == 4.0
    with pytest.raises(IndexError):
        v[4]


def test_nghost_validation():
    a = SimpleArrayFloat64((3, 2))
    with pytest.raises(IndexError, match="nghost 4 > shape\\(0\\) 3"):
        a.nghost = 4
    with pytest.raises(IndexError, match="empty array"):
        SimpleArray(array=np.zeros((0, 2))).nghost = 1
    a.nghost = 3
    assert a.nbody == 0


def test_slice_assignment_like_the_reference_tests():
    """``mh.ndcrd[:, :] = [...]``, ``mh.clnds[:, :4] = (...)``, ``mh.cltpn[:] = [...]``, ``a[...] = v`` with broadcasting:
    these address the buffer like numpy does, ghost rows included."""
    nd = SimpleArrayFloat64((4, 2))
    nd[:, :] = [(0, 0), (1, 0), (1, 1), (0, 1)]
    assert nd.ndarray.tolist() == [[0, 0], [1, 0], [1, 1], [0, 1]]
    cl = SimpleArrayInt32((2, 9))
    cl[:, :4] = ((3, 0, 1, 2), (3, 0, 2, 3))
    assert cl.ndarray[1, :5].tolist() == [3, 0, 2, 3, 0]
    tp = SimpleArrayInt32((3,))
    tp[:] = [4, 4, 3]
    assert tp.tolist() == [4, 4, 3]
    tp[...] = 7
    assert tp.tolist() == [7, 7, 7]
    cl[0] = 5            # integer key on a multi-dimensional array: flat offset, SimpleArray::operator[](size_t)
    assert cl.ndarray[0, 0] == 5 and cl[0] == 5


def test_arithmetic_and_reductions():
    """add/sub/mul/div with arrays and scalars, in-place forms, operators, min/max/sum/abs (wrap_SimpleArray.hpp:287-345)."""
    a = SimpleArray(array=np.array([1.0, -2.0, 4.0]))
    b = SimpleArray(array=np.array([2.0, 2.0, 0.5]))
    assert a.add(b).tolist() == [3.0, 0.0, 4.5] and a.add(1.0).tolist() == [2.0, -1.0, 5.0]
    assert a.sub(b).tolist() == [-1.0, -4.0, 3.5] and a.mul(b).tolist() == [2.0, -4.0, 2.0]
    assert a.div(b).tolist() == [0.5, -1.0, 8.0] and a.div(2.0).tolist() == [0.5, -1.0, 2.0]
    assert isinstance(a / np.pi, SimpleArray) and (a / 2.0).tolist() == a.div(2.0).tolist()
    assert (a + b).tolist() == a.add(b).tolist() and (a - 1.0).tolist() == [0.0, -3.0, 3.0] and (a * b).tolist() == [2.0, -4.0, 2.0]
    assert a.tolist() == [1.0, -2.0, 4.0]                      # the out-of-place forms leave the operand alone
    assert (a.min(), a.max(), a.sum()) == (-2.0, 4.0, 3.0) and a.abs().tolist() == [1.0, 2.0, 4.0]
    buf = a.ndarray
    a.iadd(b)
    a.imul(2.0)
    assert a.tolist() == [6.0, 0.0, 9.0] and a.ndarray is buf   # in place: same buffer
    a -= b
    a /= 2.0
    assert a.tolist() == [2.0, -1.0, 4.25] and a.ndarray is buf
    g = SimpleArray(array=np.arange(6.0).reshape(3, 2), nghost=1)
    assert (g * 2.0).nghost == 1 and (g * 2.0)[-1, 1] == 2.0   # ghost bookkeeping follows an elementwise result
    # numpy on the left keeps working through the buffer protocol
    assert (np.ones(3) * b).tolist() == [2.0, 2.0, 0.5]
