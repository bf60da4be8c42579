"""``EulerField`` of the reference (``solvcon/multidim/euler/_field.py``) over the device ``EulerCore``.

Same names and meaning; the scalar fields and the two reductions are evaluated ON THE DEVICE
(``scg_md_derived``, ``scg_md_reduce``) with numpy's operation order, so a caller that asks for ``mach`` after every
chunk moves 8 B per cell instead of downloading the whole state.  ``conserved`` / ``gamma`` stay host views of the
solver arrays like in the reference (writing into ``conserved`` seeds the field)."""
from __future__ import annotations

import numpy as np

__all__ = ["EulerField"]


class EulerField:
    FIELDS = ("density", "velx", "vely", "speed", "pressure", "mach", "total_energy", "cfl")   # _field.py:29-30
    MOMENTA = ("momx", "momy", "momz")

    def __init__(self, svr, mesh):
        self.svr = svr
        self.mesh = mesh

    @property
    def conserveds(self):
        return ("density",) + self.MOMENTA[:self.svr.ndim] + ("total_energy",)

    @property
    def conserved(self):
        arr = self.svr.so0n
        return arr.ndarray[arr.nghost:]

    @property
    def gamma(self):
        arr = self.svr.gamma
        return arr.ndarray[arr.nghost:]

    @property
    def centroid(self):
        arr = self.mesh.clcnd
        return arr.ndarray[arr.nghost:]

    @property
    def volume(self):
        arr = self.mesh.clvol
        return arr.ndarray[arr.nghost:]

    @property
    def vel(self):
        names = ("velx", "vely", "velz")[:self.svr.ndim]
        return np.stack([self.svr.derived(n) for n in names], axis=1)

    density = property(lambda self: self.svr.derived("density"))
    velx = property(lambda self: self.svr.derived("velx"))
    vely = property(lambda self: self.svr.derived("vely"))
    speed = property(lambda self: self.svr.derived("speed"))
    total_energy = property(lambda self: self.svr.derived("total_energy"))
    pressure = property(lambda self: self.svr.derived("pressure"))
    mach = property(lambda self: self.svr.derived("mach"))
    cfl = property(lambda self: self.svr.derived("cfl"))

    def field(self, name):
        if name not in self.FIELDS:
            raise ValueError(f"unknown field '{name}'")
        return getattr(self, name)

    def calc_overall_mass(self):
        """Density integrated over the whole domain (device reduction; _field.py:138-140)."""
        return float(self.svr.reduce(self.mesh.clvol.ndarray)[2])

    def calc_cfl_range(self):
        """Smallest and largest CFL number in the domain (device reduction; _field.py:142-148)."""
        lo, hi, _ = self.svr.reduce()
        return float(lo), float(hi)
