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
