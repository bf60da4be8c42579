"""Point / segment containers the oblique-shock application stores its node cloud and shock path in
(``solvcon/pilot/apps/obsrefl/_driver.py:84-115,140-258``, ``_analytic.py:168-180``).  Containers only: append, index,
``pack_array()``.  The geometry engine of ``cpp/solvcon/universe`` (R-tree, clipping, Bezier ...) is out of scope."""
from __future__ import annotations

import numpy as np

from solvcon_b200.simplearray import SimpleArray


class PointPadFp64:
    def __init__(self, ndim, nelem=0):
        if ndim not in (2, 3):
            raise ValueError(f"PointPad::PointPad: ndim = {ndim} must be 2 or 3")
        self._ndim = int(ndim)
        self._c = [[0.0] * int(nelem) for _ in range(self._ndim)]

    ndim = property(lambda self: self._ndim)

    def __len__(self):
        return len(self._c[0])

    def append(self, *xyz):
        if len(xyz) != self._ndim:
            raise ValueError(f"PointPad::append: {len(xyz)} coordinates given, ndim = {self._ndim}")
        for col, v in zip(self._c, xyz):
            col.append(float(v))

    def _at(self, d, i):
        if i < 0 or i >= len(self):
            raise IndexError(f"PointPad: index {i} is out of bounds with size {len(self)}")
        return self._c[d][i]

    def x_at(self, i): return self._at(0, i)
    def y_at(self, i): return self._at(1, i)
    def z_at(self, i): return self._at(2, i)

    def pack_array(self):
        """``[npoint, ndim]`` SimpleArray of the coordinates."""
        return SimpleArray(array=np.array(self._c, dtype=np.float64).T.copy().reshape(len(self), self._ndim))


class _Segment:
    __slots__ = ("x0", "y0", "z0", "x1", "y1", "z1")

    def __init__(self, p0, p1):
        self.x0, self.y0, self.z0 = p0
        self.x1, self.y1, self.z1 = p1


class SegmentPadFp64:
    def __init__(self, ndim, nelem=0):
        if ndim not in (2, 3):
            raise ValueError(f"SegmentPad::SegmentPad: ndim = {ndim} must be 2 or 3")
        self._ndim = int(ndim)
        self._s = []

    ndim = property(lambda self: self._ndim)

    def __len__(self):
        return len(self._s)

    def append(self, *c):
        if len(c) != 2 * self._ndim:
            raise ValueError(f"SegmentPad::append: {len(c)} coordinates given, ndim = {self._ndim}")
        c = [float(v) for v in c]
        p0, p1 = c[:self._ndim], c[self._ndim:]
        if self._ndim == 2:
            p0, p1 = p0 + [0.0], p1 + [0.0]
        self._s.append(_Segment(p0, p1))

    def __getitem__(self, i):
        if i < 0 or i >= len(self._s):
            raise IndexError(f"SegmentPad: index {i} is out of bounds with size {len(self._s)}")
        return self._s[i]

    def pack_array(self):
        """``[nsegment, 2*ndim]`` SimpleArray: (x0, y0[, z0], x1, y1[, z1]) per row."""
        names = ("x0", "y0", "z0", "x1", "y1", "z1") if self._ndim == 3 else ("x0", "y0", "x1", "y1")
        a = np.array([[getattr(s, n) for n in names] for s in self._s], dtype=np.float64).reshape(len(self._s), len(names))
        return SimpleArray(array=a)
