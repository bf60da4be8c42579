"""``GradientElement`` of the reference as a host-side probe (``cpp/solvcon/multidim/GradientElement.{hpp,cpp}``, binding
``multidim/pymod/wrap_multidim.cpp:79-208``): the gradient evaluation points of ONE cell, their displacements from the
centroid of the general gradient element, the fundamental-gradient-element tables and the per-element gradient solve.

The march itself never builds these objects -- ``md_dsoln_eqo_kernel`` expands the same arithmetic per cell on the
device (``scg_md_kernels.cuh``).  This class exists because it is part of the drop-in surface: the
reference's tests and callers probe single cells with it.  One cell at a time, so it is plain Python floats: the same
IEEE operations in the same order as the reference (Python never contracts to FMA), bit for bit
(``tests/test_gradient_element.py`` against the compiled reference).
"""
from __future__ import annotations

import numpy as np

__all__ = ["GradientElement"]

# GradientElementTypeGroup, GradientElement.cpp:16-79: cell type -> tuple of FGEs, each a tuple of 1-based local faces
_FGE = {
    3: ((1, 2), (2, 3), (3, 4), (4, 1)),
    4: ((1, 2), (2, 3), (3, 1)),
    5: ((2, 3, 5), (6, 3, 2), (4, 3, 6), (5, 3, 4), (5, 1, 2), (2, 1, 6), (6, 1, 4), (4, 1, 5)),
    6: ((3, 1, 2), (2, 1, 4), (4, 1, 3), (2, 4, 3)),
    7: ((5, 2, 4), (3, 2, 5), (4, 2, 3), (4, 1, 5), (5, 1, 3), (3, 1, 4)),
    8: ((1, 5, 2), (2, 5, 3), (3, 5, 4), (4, 5, 1), (1, 3, 4), (3, 1, 2)),
}


def _body(a):
    """(ndarray incl. ghost rows, nghost) of a SimpleArray-like or ndarray."""
    nd = getattr(a, "ndarray", a)
    return np.asarray(nd), int(getattr(a, "nghost", 0))


def _det(a, nd):
    if nd == 2:
        return a[0][0] * a[1][1] - a[0][1] * a[1][0]
    return (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1])
            - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
            + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))


def _adj(a, nd):
    if nd == 2:
        return [[a[1][1], -a[0][1]], [-a[1][0], a[0][0]]]
    return [[a[1][1] * a[2][2] - a[1][2] * a[2][1], a[0][2] * a[2][1] - a[0][1] * a[2][2], a[0][1] * a[1][2] - a[0][2] * a[1][1]],
            [a[1][2] * a[2][0] - a[1][0] * a[2][2], a[0][0] * a[2][2] - a[0][2] * a[2][0], a[0][2] * a[1][0] - a[0][0] * a[1][2]],
            [a[1][0] * a[2][1] - a[1][1] * a[2][0], a[0][1] * a[2][0] - a[0][0] * a[2][1], a[0][0] * a[1][1] - a[0][1] * a[1][0]]]


class GradientElement:
    def __init__(self, mesh, cecnd, icl, tau):
        nd = int(mesh.ndim)
        if nd < 2 or nd > 3:
            raise ValueError("GradientElement: ndim must be 2 or 3")
        ce, ngc = _body(cecnd)
        clfcs, ngf = _body(mesh.clfcs)
        fccls, ngfc = _body(mesh.fccls)
        clcnd, ngcl = _body(mesh.clcnd)
        cltpn, ngt = _body(mesh.cltpn)
        icl, tau = int(icl), float(tau)
        self._icl, self._ndim = icl, nd
        self._clnfc = nfc = int(clfcs[ngf + icl, 0])
        self._fge = _FGE[int(cltpn[ngt + icl])]
        icecnd = [float(ce[ngc + icl, d]) for d in range(nd)]
        self._rcls, gp, jd = [], [], []
        for ifl in range(nfc):
            ifc = int(clfcs[ngf + icl, ifl + 1])
            jcl = int(fccls[ngfc + ifc, 0]) + int(fccls[ngfc + ifc, 1]) - icl      # StaticMesh::fcrcl, StaticMesh.hpp:355-358
            self._rcls.append(jcl)
            g, j = [], []
            for d in range(nd):
                mid = float(ce[ngc + icl, (ifl + 1) * nd + d])
                jce = float(ce[ngc + jcl, d]) if jcl >= 0 else float(clcnd[ngcl + jcl, d])
                p = mid + tau * (jce - mid)
                g.append(p)
                j.append(p - jce)
            gp.append(g)
            jd.append(j)
        avg = [0.0] * nd
        for ifl in range(nfc):
            for d in range(nd):
                avg[d] += gp[ifl][d]
        for d in range(nd):
            avg[d] /= nfc
        # calc_gge_centroid, GradientElement.cpp:81-133
        cnd, voc = [0.0] * nd, 0.0
        for tface in self._fge:
            sub = list(avg)
            dst = []
            for ivx in range(nd):
                fl = tface[ivx] - 1
                row = []
                for d in range(nd):
                    sub[d] += gp[fl][d]
                    row.append(gp[fl][d] - avg[d])
                dst.append(row)
            for d in range(nd):
                sub[d] /= float(nd + 1)
            vob = _det(dst, nd)
            voc += vob
            for d in range(nd):
                cnd[d] += sub[d] * vob
        for d in range(nd):
            cnd[d] /= voc
        self._idis = [[gp[f][d] - cnd[d] for d in range(nd)] for f in range(nfc)]
        self._jdis = [[jd[f][d] + icecnd[d] - cnd[d] for d in range(nd)] for f in range(nfc)]

    # -- wrap_management ------------------------------------------------------------------------------
    icl = property(lambda self: self._icl)
    ndim = property(lambda self: self._ndim)
    clnfc = property(lambda self: self._clnfc)
    nfge = property(lambda self: len(self._fge))
    nfge_inverse = property(lambda self: 1.0 / len(self._fge))

    # -- bounds, wrap_multidim.cpp:42-66 ----------------------------------------------------------------
    def _check_ifl(self, ifl):
        if ifl < 0 or ifl >= self._clnfc:
            raise IndexError(f"GradientElement: ifl {ifl} out of range [0, {self._clnfc})")

    def _check_d(self, d):
        if d < 0 or d >= self._ndim:
            raise IndexError(f"GradientElement: d {d} out of range [0, {self._ndim})")

    def _check_ifge(self, ifge):
        if ifge < 0 or ifge >= len(self._fge):
            raise IndexError(f"GradientElement: ifge {ifge} out of range [0, {len(self._fge)})")

    # -- wrap_gradient ----------------------------------------------------------------------------------
    def rcl(self, ifl):
        self._check_ifl(ifl)
        return self._rcls[ifl]

    def idis(self, ifl, d):
        self._check_ifl(ifl)
        self._check_d(d)
        return self._idis[ifl][d]

    def jdis(self, ifl, d):
        self._check_ifl(ifl)
        self._check_d(d)
        return self._jdis[ifl][d]

    def faces(self, ifge):
        self._check_ifge(ifge)
        return list(self._fge[ifge][:self._ndim])

    def displacement_matrix(self, ifge):
        self._check_ifge(ifge)
        return [list(self._idis[fl - 1]) for fl in self._fge[ifge][:self._ndim]]

    def solve_gradient(self, ifge, udf):
        self._check_ifge(ifge)
        udf = list(udf)
        nd = self._ndim
        if len(udf) != nd:
            raise ValueError(f"GradientElement: udf size {len(udf)} must equal ndim {nd}")
        dst = self.displacement_matrix(ifge)
        adj = _adj(dst, nd)
        r = [0.0] * nd
        for i in range(nd):
            for j in range(nd):
                r[i] += adj[i][j] * float(udf[j])
        det = _det(dst, nd)
        return [r[d] / det for d in range(nd)]
