"""Oblique-shock reflection on the device ``EulerCore`` -- the end-to-end acceptance run of SURVEY.md 8(f) row f4.

Mirrors the driver layer of the reference's one real application over ``EulerCore``
(``solvcon/pilot/apps/obsrefl/_driver.py``: ``ObliqueShockMesher`` :21-320, ``ObliqueShockRelation`` :286-406,
``ObliqueShock`` :408-551) with the same class and method names, arguments and defaults, so a script written against
the reference runs here unchanged; GUI, session and panels of that package are out of scope.  A Mach-3 stream enters
from the left, the post-shock state of a 10 degree deflection is imposed along the top, the bottom is a slip wall and
the right an outflow: the incident shock reflects off the wall and the steady field has three uniform zones whose
states the oblique-shock relations give in closed form -- the physics-level check the unit tests lack.

The mesher builds the reference's three mesh flavours node for node (``tests/golden/make_obsrefl_golden.py`` runs the
reference's own mesher source against these classes' output): 'quad', 'triangle' (boxes cut along the ll-ur
diagonal, flipped at two corners) and 'unstructured' (Bowyer-Watson Delaunay triangulation of the boundary nodes
plus RNG-free jittered interior points, with Steiner points until no cell has two boundary faces).
"""
from __future__ import annotations

import math

import numpy as np

from ..mesh import StaticMesh
from ..multidim import EulerCore

__all__ = ["ObliqueShock", "ObliqueShockMesher", "ObliqueShockRelation", "measure_zones"]

QUADRILATERAL, TRIANGLE = 3, 4


class ObliqueShockMesher:
    BOUNDARY_NAMES = ("left", "top", "bottom", "right")

    def __init__(self, nx=64, ny=16, ll=(0.0, 0.0), ur=(4.0, 1.0)):
        self.nx, self.ny = nx, ny
        (self.x0, self.y0), (self.x1, self.y1) = ll, ur

    @property
    def cell_extent(self):
        return ((self.x1 - self.x0) / self.nx, (self.y1 - self.y0) / self.ny)

    def _node(self, it, jt):
        return (self.x0 + it * (self.x1 - self.x0) / self.nx, self.y0 + jt * (self.y1 - self.y0) / self.ny)

    def _nid(self, it, jt):
        return jt * (self.nx + 1) + it

    def _box(self, it, jt):
        return (self._nid(it, jt), self._nid(it + 1, jt), self._nid(it + 1, jt + 1), self._nid(it, jt + 1))

    # -- nodes and cells of the three flavours -------------------------------------------------------------
    def nodes_and_cells(self, cell_type="unstructured"):
        nx, ny = self.nx, self.ny
        if cell_type in ("quad", "triangle"):
            nodes = [self._node(it, jt) for jt in range(ny + 1) for it in range(nx + 1)]
            if cell_type == "quad":
                return nodes, QUADRILATERAL, [(4,) + self._box(it, jt) for jt in range(ny) for it in range(nx)]
            cells = []
            for jt in range(ny):
                for it in range(nx):
                    ll, lr, ur, ul = self._box(it, jt)
                    # the two corner boxes are cut the other way: their corner cell would carry both boundary edges
                    if (it, jt) in ((0, ny - 1), (nx - 1, 0)):
                        cells += [(3, ll, lr, ul), (3, lr, ur, ul)]
                    else:
                        cells += [(3, ll, lr, ur), (3, ll, ur, ul)]
            return nodes, TRIANGLE, cells
        if cell_type == "unstructured":
            nodes = self._jitter_points()
            return nodes, TRIANGLE, [(3,) + tri for tri in self._split_double_boundary(nodes)]
        raise ValueError(f"unknown cell_type '{cell_type}'")

    def make_mesh(self, cell_type="unstructured"):
        nodes, tpn, cells = self.nodes_and_cells(cell_type)
        mh = StaticMesh(ndim=2, nnode=len(nodes), nface=0, ncell=len(cells))
        mh.ndcrd.ndarray[...] = np.asarray(nodes, dtype=np.float64)
        mh.cltpn.ndarray[...] = tpn
        mh.clnds.ndarray[:, :len(cells[0])] = np.asarray(cells, dtype=np.int32)
        mh.build_interior(do_metric=True)
        for name, faces in zip(self.BOUNDARY_NAMES, self.classify_boundary(mh)):
            mh.add_bc(name, faces)
        mh.build_boundary()
        mh.build_ghost()
        return mh

    def _jitter_points(self):
        """Counter-clockwise outline walk of the structured boundary nodes, then the interior nodes displaced in
        logical space by an RNG-free phase bounded by 0.3 boxes per axis (_driver.py:129-153)."""
        nx, ny = self.nx, self.ny
        pts = [self._node(it, 0) for it in range(nx + 1)]
        pts += [self._node(nx, jt) for jt in range(1, ny + 1)]
        pts += [self._node(it, ny) for it in range(nx - 1, -1, -1)]
        pts += [self._node(0, jt) for jt in range(ny - 1, 0, -1)]
        for jt in range(1, ny):
            for it in range(1, nx):
                pts.append(self._node(it + 0.3 * math.sin(12.9898 * it + 78.233 * jt),
                                      jt + 0.3 * math.cos(26.651 * it + 41.347 * jt)))
        return pts

    @staticmethod
    def _circumcircle(pts, ia, ib, ic):
        (ax, ay), (bx, by), (cx, cy) = pts[ia], pts[ib], pts[ic]
        dd = 2.0 * (ax * (by - cy) + bx * (cy - ay) + cx * (ay - by))
        a2, b2, c2 = ax * ax + ay * ay, bx * bx + by * by, cx * cx + cy * cy
        ux = (a2 * (by - cy) + b2 * (cy - ay) + c2 * (ay - by)) / dd
        uy = (a2 * (cx - bx) + b2 * (ax - cx) + c2 * (bx - ax)) / dd
        return ux, uy, (ax - ux) ** 2 + (ay - uy) ** 2

    @classmethod
    def _triangulate(cls, points):
        """Bowyer-Watson: insert the points one at a time into a far super-triangle, carving the triangles whose
        circumcircle strictly contains the point (on-circle counts as outside) and fanning the cavity rim to it
        (_driver.py:169-228; insertion order and tie rule decide the result, so they are kept)."""
        npt = len(points)
        pts = list(points)
        xs, ys = [p[0] for p in pts], [p[1] for p in pts]
        xmid, ymid = (min(xs) + max(xs)) / 2.0, (min(ys) + max(ys)) / 2.0
        span = max(max(xs) - min(xs), max(ys) - min(ys), 1.0)
        pts += [(xmid - 64.0 * span, ymid - 32.0 * span), (xmid + 64.0 * span, ymid - 32.0 * span), (xmid, ymid + 64.0 * span)]
        cc = {(npt, npt + 1, npt + 2): cls._circumcircle(pts, npt, npt + 1, npt + 2)}
        for ip in range(npt):
            px, py = pts[ip]
            bad = [tri for tri, (ux, uy, rr) in cc.items() if (px - ux) ** 2 + (py - uy) ** 2 < rr * (1.0 - 1e-12)]
            dedges = set()
            for ta, tb, tc in bad:
                dedges.update(((ta, tb), (tb, tc), (tc, ta)))
                del cc[(ta, tb, tc)]
            for ea, eb in dedges:
                if (eb, ea) not in dedges:
                    cc[(ea, eb, ip)] = cls._circumcircle(pts, ea, eb, ip)
        tris = []
        for tri in cc:
            if max(tri) >= npt:
                continue
            (ax, ay), (bx, by), (cx, cy) = (pts[iv] for iv in tri)
            if (bx - ax) * (cy - ay) - (cx - ax) * (by - ay) < 0.0:
                tri = (tri[0], tri[2], tri[1])
            tris.append(tri)
        return tris

    @classmethod
    def _split_double_boundary(cls, points):
        """Triangulate, adding a Steiner point at the centroid of every cell that has two boundary edges, until none
        is left (_driver.py:230-268).  ``points`` grows in place."""
        for _ in range(16):
            tris = cls._triangulate(points)
            shared = {}
            for tri in tris:
                for it in range(3):
                    edge = frozenset((tri[it], tri[(it + 1) % 3]))
                    shared[edge] = shared.get(edge, 0) + 1
            extra = []
            for tri in tris:
                nb = sum(1 for it in range(3) if shared[frozenset((tri[it], tri[(it + 1) % 3]))] == 1)
                if nb >= 2:
                    extra.append((sum(points[iv][0] for iv in tri) / 3.0, sum(points[iv][1] for iv in tri) / 3.0))
            if not extra:
                return tris
            points.extend(extra)
        raise RuntimeError("boundary-cell refinement did not converge")

    @staticmethod
    def classify_boundary(mh, tol=1e-9):
        """(left, top, bottom, right) sorted boundary-face lists by face centre and outward normal (_driver.py:270-318)."""
        nface = mh.nface
        fccls, fccnd, fcnml = mh.fccls.body[:nface], mh.fccnd.body[:nface], mh.fcnml.body[:nface]
        b = np.nonzero(fccls[:, 1] < 0)[0]
        xc, yc, nx, ny = fccnd[b, 0], fccnd[b, 1], fcnml[b, 0], fcnml[b, 1]
        xmin, xmax, ymax = xc.min(), xc.max(), yc.max()
        left = (np.abs(xc - xmin) <= tol) & (nx < 0.0)
        right = ~left & (np.abs(xc - xmax) <= tol) & (nx > 0.0)
        top = ~left & ~right & (np.abs(yc - ymax) <= tol) & (ny > 0.0)
        bottom = ~left & ~right & ~top
        return tuple(sorted(b[m].tolist()) for m in (left, top, bottom, right))


class ObliqueShockRelation:
    """Flow-property jumps across an oblique shock (Anderson, Modern Compressible Flow, ch. 4); ``beta`` = shock angle,
    ``theta`` = deflection, radians (_driver.py:286-406)."""

    def __init__(self, gamma):
        self.gamma = gamma

    def calc_density_ratio(self, mach1, beta):
        g, mn1sq = self.gamma, (mach1 * math.sin(beta)) ** 2
        return (g + 1) * mn1sq / ((g - 1) * mn1sq + 2)

    def calc_pressure_ratio(self, mach1, beta):
        g, mn1sq = self.gamma, (mach1 * math.sin(beta)) ** 2
        return 1 + 2 * g / (g + 1) * (mn1sq - 1)

    def calc_temperature_ratio(self, mach1, beta):
        return self.calc_pressure_ratio(mach1, beta) / self.calc_density_ratio(mach1, beta)

    def calc_normal_dmach(self, mach_n1):
        g, mn1sq = self.gamma, mach_n1 * mach_n1
        return math.sqrt(((g - 1) * mn1sq + 2) / (2 * g * mn1sq - (g - 1)))

    def calc_dmach(self, mach1, beta=None, theta=None, delta=1):
        if (beta is None) == (theta is None):
            raise ValueError(f"got (beta={beta}, theta={theta}), but I need to take either beta or theta")
        if beta is None:
            beta = self.calc_shock_angle(mach1, theta, delta=delta)
        if theta is None:
            theta = self.calc_flow_angle(mach1, beta)
        return self.calc_normal_dmach(mach1 * math.sin(beta)) / math.sin(beta - theta)

    def calc_flow_angle(self, mach1, beta):
        return math.atan(self.calc_flow_tangent(mach1, beta))

    def calc_flow_tangent(self, mach1, beta):
        g, m1sq = self.gamma, mach1 * mach1
        return 2 / math.tan(beta) * (m1sq * math.sin(beta) ** 2 - 1) / (m1sq * (g + math.cos(2 * beta)) + 2)

    def calc_shock_angle(self, mach1, theta, delta=1):
        return math.atan(self.calc_shock_tangent(mach1, theta, delta))

    def calc_shock_tangent(self, mach1, theta, delta):
        g, m1sq = self.gamma, mach1 * mach1
        lmbd, chi = self.calc_shock_tangent_aux(mach1, theta)
        num = m1sq - 1 + 2 * lmbd * math.cos((4 * math.pi * delta + math.acos(chi)) / 3)
        return num / (3 * (1 + (g - 1) / 2 * m1sq) * math.tan(theta))

    def calc_shock_tangent_aux(self, mach1, theta):
        g, m1sq, tansq = self.gamma, mach1 * mach1, math.tan(theta) ** 2
        disc = (m1sq - 1) ** 2 - 3 * (1 + (g - 1) / 2 * m1sq) * (1 + (g + 1) / 2 * m1sq) * tansq
        if disc <= 0.0:
            raise ValueError(f"no attached shock for mach1={mach1:g} and theta={math.degrees(theta):g} deg")
        lmbd = math.sqrt(disc)
        chi = ((m1sq - 1) ** 3 - 9 * (1 + (g - 1) / 2 * m1sq)
               * (1 + (g - 1) / 2 * m1sq + (g + 1) / 4 * m1sq * m1sq) * tansq) / lmbd ** 3
        return lmbd, chi


class ObliqueShock:
    """Drive the device CESE Euler core over the oblique-shock reflection (_driver.py:408-551)."""

    def __init__(self):
        for k in ("gamma", "density", "pressure", "mach", "speedofsound", "velocity", "relation", "theta", "shock_angle",
                  "density2", "pressure2", "mach2", "velocity2", "shock_angle2", "density3", "pressure3", "mach3",
                  "velocity3", "mesher", "mesh", "svr"):
            setattr(self, k, None)

    def build_constant(self, gamma=1.4, density=1.0, pressure=1.0, mach=3.0, angle=10.0):
        self.gamma, self.density, self.pressure, self.mach = gamma, density, pressure, mach
        self.speedofsound = math.sqrt(gamma * pressure / density)
        self.velocity = mach * self.speedofsound
        self.relation = rel = ObliqueShockRelation(gamma=gamma)
        self.theta = theta = math.radians(angle)
        self.shock_angle = beta = rel.calc_shock_angle(mach, theta)
        self.density2 = density * rel.calc_density_ratio(mach, beta)
        self.pressure2 = pressure * rel.calc_pressure_ratio(mach, beta)
        self.mach2 = mach2 = rel.calc_dmach(mach, beta=beta)
        speed2 = mach2 * math.sqrt(gamma * self.pressure2 / self.density2)
        self.velocity2 = (speed2 * math.cos(theta), -speed2 * math.sin(theta))
        self.shock_angle2 = beta2 = rel.calc_shock_angle(mach2, theta)
        self.density3 = self.density2 * rel.calc_density_ratio(mach2, beta2)
        self.pressure3 = self.pressure2 * rel.calc_pressure_ratio(mach2, beta2)
        self.mach3 = rel.calc_dmach(mach2, beta=beta2)
        self.velocity3 = (self.mach3 * math.sqrt(gamma * self.pressure3 / self.density3), 0.0)

    def build_numerical(self, cell_type="unstructured", time_increment=2.e-3, sigma0=3.0, taumin=0.0, tauscale=1.0,
                        order="morton", **mesher_kw):
        if self.gamma is None:
            raise ValueError("constants are not set; call build_constant()")
        self.mesher = ObliqueShockMesher(**mesher_kw)
        self.mesh = self.mesher.make_mesh(cell_type=cell_type)
        svr = EulerCore(self.mesh, time_increment, order=order)
        svr.sigma0, svr.taumin, svr.tauscale = sigma0, taumin, tauscale
        svr.init_solution(gamma=self.gamma, rho=self.density, v=[self.velocity, 0.0], p=self.pressure)
        left, top, bottom, right = (self.mesh.bc(name).facn.ndarray[:, 0].tolist() for name in self.mesher.BOUNDARY_NAMES)
        svr.add_inlet(left, value=[self.density, self.velocity, 0.0, self.pressure, self.gamma])
        svr.add_inlet(top, value=[self.density2, self.velocity2[0], self.velocity2[1], self.pressure2, self.gamma])
        svr.add_slipwall(bottom)
        svr.add_nonrefl(right)
        svr.bc_soln()    # prime the ghost rows from the initial interior state
        svr.bc_dsoln()
        self.svr = svr

    def march(self, steps):
        self.svr.march(steps=steps)

    def zone_states(self):
        return [(self.density, self.velocity, 0.0, self.pressure),
                (self.density2, self.velocity2[0], self.velocity2[1], self.pressure2),
                (self.density3, self.velocity3[0], self.velocity3[1], self.pressure3)]

    def shock_path(self):
        if self.mesher is None:
            raise ValueError("mesh is not built; call build_numerical()")
        msh = self.mesher
        xhit = msh.x0 + (msh.y1 - msh.y0) / math.tan(self.shock_angle)
        if xhit >= msh.x1:
            return [(msh.x0, msh.y1), (msh.x1, msh.y1 - (msh.x1 - msh.x0) * math.tan(self.shock_angle))]
        slope = math.tan(self.shock_angle2 - self.theta)
        xend = min(msh.x1, xhit + (msh.y1 - msh.y0) / slope)
        return [(msh.x0, msh.y1), (xhit, msh.y0), (xend, msh.y0 + (xend - xhit) * slope)]


def measure_zones(shock, margin=0.12):
    """Mean (rho, vx, vy, p) of the computed field over the cells well inside each of the three analytic zones (at
    least ``margin`` away, in y, from both shocks and from the outline) -- the acceptance measurement: a converged
    run has to hold ``shock.zone_states()`` there.  Uses the device-side derived fields (EulerField, row f2)."""
    from ..euler_field import EulerField
    fld = EulerField(shock.svr, shock.mesh)
    c = fld.centroid
    x, y = c[:, 0], c[:, 1]
    msh = shock.mesher
    (xa, ya), (xh, yh), (xe, ye) = shock.shock_path()
    inc = ya + (x - xa) * (yh - ya) / (xh - xa)          # incident shock height at x
    ref = yh + (x - xh) * (ye - yh) / (xe - xh)          # reflected shock height at x
    inside = (x > msh.x0 + margin) & (x < msh.x1 - margin) & (y > msh.y0 + 0.5 * margin) & (y < msh.y1 - 0.5 * margin)
    zones = [inside & (y < inc - margin) & (x < xh - margin),
             inside & (y > inc + margin) & ((y > ref + margin) | (x < xh)),
             inside & (x > xh + margin) & (y < ref - margin)]
    rho, vx, vy, p = fld.density, fld.velx, fld.vely, fld.pressure
    out = []
    for z in zones:
        n = int(z.sum())
        out.append((n,) + tuple(float(a[z].mean()) if n else float("nan") for a in (rho, vx, vy, p)))
    return out
