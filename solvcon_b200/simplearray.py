"""Minimal host array with the reference's ``SimpleArray`` contract.

Only the subset the CESE front ends and tests use is mirrored
(``cpp/solvcon/buffer/pymod/wrap_SimpleArray.hpp:65-141,193-214``,
``cpp/solvcon/buffer/SimpleArray.hpp:2241-2270,2480-2512``):

* ``ndarray``      zero-copy numpy view of the whole buffer, ghost rows first
* ``nghost`` (rw), ``nbody``, ``has_ghost``, ``shape``, ``stride``, ``size``, ``nbytes``
* ``fill``, ``clone``
* ``a[i]`` / ``a[i, j, ...]`` with ghost-shifted first index: row ``-1`` is the first ghost
  row, ``IndexError`` outside ``[-nghost, nbody)``
* ``a[...] = value`` for scalars, tuples, slices / tuples of slices / ellipsis with
  broadcasting (these address the buffer like numpy does, ghost rows included)
* the buffer protocol via ``__array__`` so ``np.testing.assert_*`` accept it directly
* ``add/sub/mul/div`` (+ in-place forms and the Python operators), ``min/max/sum/abs`` -- what front ends apply to
  solver output (``svr.xctr() / np.pi``)

Everything else of the reference class (sort, matmul, SIMD, plex, ...) is off the march path.
"""
from __future__ import annotations

import numpy as np


class SimpleArray:
    __slots__ = ("_a", "_nghost", "_owner")

    def __init__(self, shape=None, value=None, dtype=np.float64, array=None, nghost=0, pinned=False):
        if array is not None:
            self._a = array if isinstance(array, np.ndarray) else np.asarray(array)
        else:
            if isinstance(shape, (int, np.integer)):
                shape = (int(shape),)
            if pinned:
                # page-locked memory for the host twins of device arrays (additive: ``pinned`` is not a reference
                # argument); plain memory where no device is available
                from . import _lib
                self._a = _lib.pinned_zeros(tuple(shape), dtype)
            else:
                self._a = np.zeros(tuple(shape), dtype=dtype)
            if value is not None:
                self._a.fill(value)
        self._nghost = 0
        if nghost:
            self.nghost = nghost

    # -- views and metadata -----------------------------------------------------
    @property
    def ndarray(self):
        return self._a

    def __array__(self, dtype=None, copy=None):
        if dtype is not None and np.dtype(dtype) != self._a.dtype:
            return self._a.astype(dtype)
        return self._a

    @property
    def shape(self):
        return tuple(self._a.shape)

    @property
    def stride(self):
        return tuple(s // self._a.itemsize for s in self._a.strides)

    @property
    def size(self):
        return int(self._a.size)

    @property
    def nbytes(self):
        return int(self._a.nbytes)

    @property
    def itemsize(self):
        return int(self._a.itemsize)

    def __len__(self):
        return int(self._a.size)

    @property
    def nghost(self):
        return self._nghost

    @nghost.setter
    def nghost(self, n):
        # SimpleArray::set_nghost, SimpleArray.hpp:2244-2270
        n = int(n)
        if n != 0:
            if self._a.ndim == 0 or self._a.size == 0:
                raise IndexError(f"SimpleArray: cannot set nghost {n} > 0 to an empty array")
            if n > self._a.shape[0]:
                raise IndexError(f"SimpleArray: cannot set nghost {n} > shape(0) {self._a.shape[0]}")
        self._nghost = n

    @property
    def nbody(self):
        return (self._a.shape[0] - self._nghost) if self._a.ndim else 0

    @property
    def has_ghost(self):
        return self._nghost != 0

    @property
    def body(self):
        """numpy view of the body rows (additive convenience)."""
        return self._a[self._nghost:]

    @property
    def ghost(self):
        """numpy view of the ghost rows in *reference index order*: row 0 is index -1."""
        return self._a[:self._nghost][::-1]

    def fill(self, value):
        self._a.fill(value)

    def clone(self):
        out = SimpleArray(array=self._a.copy())
        out._nghost = self._nghost
        return out

    def tolist(self):
        return self._a.tolist()

    # -- elementwise arithmetic and reductions the front ends use on solver output ---------------------------------
    # (wrap_SimpleArray.hpp:287-345,470-500: add/sub/mul/div with an array or a scalar, their in-place forms,
    #  min/max/sum/abs; the Python operators map onto the same calls)
    @staticmethod
    def _operand(other):
        return other._a if isinstance(other, SimpleArray) else other

    def _wrap(self, a):
        out = SimpleArray(array=a)
        out._nghost = self._nghost if a.shape[:1] == self._a.shape[:1] else 0
        return out

    def add(self, other): return self._wrap(self._a + self._operand(other))
    def sub(self, other): return self._wrap(self._a - self._operand(other))
    def mul(self, other): return self._wrap(self._a * self._operand(other))
    def div(self, other): return self._wrap(self._a / self._operand(other))

    def iadd(self, other): self._a += self._operand(other)
    def isub(self, other): self._a -= self._operand(other)
    def imul(self, other): self._a *= self._operand(other)
    def idiv(self, other): self._a /= self._operand(other)

    __add__, __sub__, __mul__, __truediv__ = add, sub, mul, div

    def __iadd__(self, other):
        self.iadd(other)
        return self

    def __isub__(self, other):
        self.isub(other)
        return self

    def __imul__(self, other):
        self.imul(other)
        return self

    def __itruediv__(self, other):
        self.idiv(other)
        return self

    def min(self): return self._a.min().item()
    def max(self): return self._a.max().item()
    def sum(self): return self._a.sum().item()
    def abs(self): return self._wrap(np.abs(self._a))

    # -- element access ------------------------------------------------------------
    def _shift(self, i):
        i = int(i)
        if i < -self._nghost or i >= self.nbody:
            raise IndexError(
                f"SimpleArray: index {i} is out of bounds with body size {self.nbody} and ghost size {self._nghost}")
        return i + self._nghost

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            if self._a.ndim == 1:
                return self._a[self._shift(key)]
            # flat offset into the buffer, like SimpleArray::operator[](size_t)
            return self._a.reshape(-1)[int(key)]
        if isinstance(key, tuple) and key and all(isinstance(k, (int, np.integer)) for k in key):
            return self._a[(self._shift(key[0]),) + tuple(int(k) for k in key[1:])]
        return self._a[key]

    def __setitem__(self, key, value):
        if isinstance(key, (int, np.integer)):
            if self._a.ndim == 1:
                self._a[self._shift(key)] = value
            else:
                self._a.reshape(-1)[int(key)] = value
            return
        if isinstance(key, tuple) and key and all(isinstance(k, (int, np.integer)) for k in key):
            self._a[(self._shift(key[0]),) + tuple(int(k) for k in key[1:])] = value
            return
        self._a[key] = value

    def __repr__(self):
        return f"SimpleArray(shape={self.shape}, nghost={self._nghost}, dtype={self._a.dtype})"


def SimpleArrayFloat64(shape=None, value=None, array=None):
    return SimpleArray(shape, value, np.float64, array)


def SimpleArrayInt32(shape=None, value=None, array=None):
    return SimpleArray(shape, value, np.int32, array)
