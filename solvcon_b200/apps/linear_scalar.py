"""The 1D linear-scalar demonstration solver over the device march (``solvcon/onedim/linear_scalar.py:17-47``).

Same class name, constructor and ``init_solver`` contract as the reference front end: a uniform grid of ``ncelm``
conservation elements on ``[xmin, xmax]``, ``dt = cfl * dx``, and one period of a sine wave between ``2 pi`` and ``4 pi``
(zero elsewhere) as the initial condition of the even plane.  The reference fills it through the ``Selm`` cursors one
element at a time; here the same values go into the solver with two whole-array calls, which is what makes the front
end usable at the 10^9 points of BASELINE configs[1].  Everything else is delegated to the device solver.
"""
from __future__ import annotations

import numpy as np

__all__ = ["LinearScalarSolver"]


class LinearScalarSolver:

    def __init__(self, xmin, xmax, ncelm, cfl=1):
        self._core = self.init_solver(xmin, xmax, ncelm, cfl)

    def __getattr__(self, name):
        return getattr(self._core, name)

    @staticmethod
    def init_solver(xmin, xmax, ncelm, cfl=1):
        from .. import spacetime as st
        grid = st.Grid(xmin, xmax, ncelm)
        dx = (grid.xmax - grid.xmin) / grid.ncelm
        svr = st.LinearScalarSolver(grid=grid, time_increment=dx * cfl)
        xctr = np.asarray(svr.xctr(odd_plane=False))
        inside = np.logical_not((xctr < 2 * np.pi) | (xctr > 2 * 2 * np.pi))
        svr.set_so0(0, np.where(inside, np.sin(xctr), 0.0))
        svr.set_so1(0, np.where(inside, np.cos(xctr), 0.0))
        return svr
