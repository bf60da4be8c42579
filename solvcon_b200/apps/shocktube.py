"""Sod shock tube over the device ``Euler1DCore`` -- the caller of BASELINE.json configs[0].

Mirrors the front end of the reference (``solvcon/onedim/euler1d.py``: ``Euler1DSolver`` :21-61, ``ShockTube`` :64-392)
with the same class, method and attribute names: ``build_constant`` fixes the five-zone analytic solution (Anderson,
Modern Compressible Flow, 3/e, ch. 7: left state 1, expansion fan 2, post-fan state 3, contact surface, post-shock
state 4, right state 5), ``build_numerical`` builds the CESE solver with the diaphragm initial condition, ``build_field``
evaluates the analytic fields at a time ``t``.  The analytic part is plain numpy and needs no GPU; the numerical part
runs on the device core.  ``tests/test_shocktube.py`` pins the analytic constants with the reference's own known answers
(tests/test_onedim_euler.py:113-141); ``tests/test_gpu_1d.py`` holds the acceptance run numerical-vs-analytic.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Euler1DSolver", "ShockTube"]


class Euler1DSolver:
    """The CESE solver of the 1D Euler equations: a thin holder of ``Euler1DCore`` with a uniform grid (euler1d.py:21-61)."""

    def __init__(self, xmin, xmax, ncoord, time_increment=0.05, keep_edge=False):
        self._core = self.init_solver(xmin, xmax, ncoord, time_increment, gamma=1.4)
        last = self.ncoord - 1
        start, stop = (0, last) if keep_edge else (2, last - 2)
        self.xindices = np.linspace(start, stop, (stop - start) // 2 + 1, dtype="int32")   # the even-plane points

    def __getattr__(self, name):
        return getattr(self._core, name)

    @staticmethod
    def init_solver(xmin, xmax, ncoord, time_increment, gamma):
        from ..onedim import Euler1DCore
        svr = Euler1DCore(ncoord=ncoord, time_increment=time_increment)
        svr.coord[...] = np.linspace(xmin, xmax, num=ncoord)
        svr.cfl.fill(0)
        svr.gamma.fill(gamma)
        svr.so0[...] = 0.0
        svr.so1[...] = 0.0
        return svr

    @staticmethod
    def calc_u2(gamma, density, velocity, pressure):
        return density * (1.0 / (gamma - 1) * pressure / density + velocity * velocity / 2)


class ShockTube:
    R = 8.31446261815324
    _ZONE_FIELDS = ("velocity", "pressure", "density", "temperature", "internal_energy", "entropy", "speedofsound")
    _FIELDS = ("density", "velocity", "pressure", "temperature", "internal_energy", "entropy")

    def __init__(self):
        self.gamma = None
        for z in (1, 5, 4, 3):
            for f in self._ZONE_FIELDS:
                setattr(self, f"{f}{z}", None)
        self.velocity_shock = self.velocity_con = None
        self.coord_field = None
        for f in self._FIELDS:
            setattr(self, f + "_field", None)
        self.svr = None

    # -- thermodynamic helpers ------------------------------------------------------------------------------
    def calc_speedofsound(self, pressure, density):
        return np.sqrt(self.gamma * pressure / density)

    def calc_internal_energy(self, pressure, density):
        return pressure / (density * (self.gamma - 1))

    def calc_entropy(self, pressure, density):
        return pressure / (density ** self.gamma)

    def _fill_zone(self, z, velocity, pressure, density, temperature):
        setattr(self, f"velocity{z}", velocity)
        setattr(self, f"pressure{z}", pressure)
        setattr(self, f"density{z}", density)
        setattr(self, f"temperature{z}", temperature)
        setattr(self, f"speedofsound{z}", self.calc_speedofsound(pressure=pressure, density=density))
        setattr(self, f"internal_energy{z}", self.calc_internal_energy(pressure, density))
        setattr(self, f"entropy{z}", self.calc_entropy(pressure, density))

    # -- the analytic solution ----------------------------------------------------------------------------------
    def build_constant(self, gamma, pressure1, density1, pressure5, density5):
        self.gamma = gamma
        self._fill_zone(1, 0.0, pressure1, density1, pressure1 / (density1 * self.R))
        self._fill_zone(5, 0.0, pressure5, density5, pressure5 / (density5 * self.R))
        p45 = self.calc_pressure45()                                  # shock strength
        v4 = self.calc_velocity4(pressure45=p45)                      # normal-shock relations into zone 4
        self._fill_zone(4, v4, p45 * pressure5, self.calc_density45(pressure45=p45) * density5,
                        self.calc_temperature45(p45) * self.temperature5)
        self.velocity_shock = self.speedofsound5 * np.sqrt((gamma + 1) / (2 * gamma) * (p45 - 1) + 1)
        self.velocity_con = v4
        # across the contact surface velocity and pressure are continuous; the fan gives the density of zone 3
        self._fill_zone(3, v4, self.pressure4, self.calc_density2(x=0, t=0, velocity2=v4),
                        self.calc_temperature2(x=0, t=0, velocity2=v4))

    def calc_pressure45(self, tolerance=1.e-10, maxiter=50):
        """p4 / p5 by the secant method on the shock-tube relation (Anderson eq. 7.94)."""
        g = self.gamma
        p15 = self.pressure1 / self.pressure5
        a51 = self.speedofsound5 / self.speedofsound1

        def residue(p45):
            w = p45 - 1
            frac = (g - 1) * a51 * w / np.sqrt(2 * g * (2 * g + (g + 1) * w))
            return p15 - p45 * (1 - frac) ** (-2 * g / (g - 1))

        xa, fa = p15, residue(p15)
        xb = xa / 2
        fb = residue(xb)
        for _ in range(maxiter):
            if np.abs(fb) <= tolerance:
                break
            slope = (fb - fa) / (xb - xa)
            xa, fa = xb, fb
            xb = xb - fb / slope
            fb = residue(xb)
        return xb

    def calc_density45(self, pressure45):
        k = (self.gamma + 1) / (self.gamma - 1)
        return (1 + k * pressure45) / (k + pressure45)

    def calc_temperature45(self, pressure45):
        k = (self.gamma + 1) / (self.gamma - 1)
        return pressure45 * ((k + pressure45) / (1 + k * pressure45))

    def calc_velocity4(self, pressure45):
        g = self.gamma
        return self.speedofsound5 / g * (pressure45 - 1) * np.sqrt((2 * g / (g + 1)) / (pressure45 + (g - 1) / (g + 1)))

    def calc_velocity2(self, x, t):
        return 2 / (self.gamma + 1) * (self.speedofsound1 + x / t)

    def calc_speedofsound2_ratio(self, x, t, velocity2=None):
        if velocity2 is None:
            velocity2 = self.calc_velocity2(x, t)
        return 1 - (self.gamma - 1) / 2 * (velocity2 / self.speedofsound1)

    def calc_pressure2(self, x, t):
        return self.pressure1 * self.calc_speedofsound2_ratio(x, t) ** (2 * self.gamma / (self.gamma - 1))

    def calc_density2(self, x, t, velocity2=None):
        return self.density1 * self.calc_speedofsound2_ratio(x, t, velocity2=velocity2) ** (2 / (self.gamma - 1))

    def calc_temperature2(self, x, t, velocity2=None):
        return self.temperature1 * self.calc_speedofsound2_ratio(x, t, velocity2=velocity2) ** 2

    def calc_locations(self, t):
        """[x12, x23, x34, x45]: head and tail of the fan, contact surface, shock."""
        return np.array([-self.speedofsound1 * t, (self.velocity3 - self.speedofsound3) * t, self.velocity3 * t,
                         self.velocity_shock * t], dtype="float64")

    def build_field(self, t, coord=None):
        """Analytic fields at time ``t`` on ``coord`` (default: the even-plane points of the numerical solver)."""
        if coord is None:
            coord = self.svr.coord[self.svr.xindices]
        x = self.coord_field = np.array(coord, dtype="float64")
        l12, l23, l34, l45 = self.calc_locations(t=t)
        zone = np.select([x < l12, x < l23, x < l34, x < l45], [1, 2, 3, 4], default=5)
        for f in self._FIELDS:
            out = np.zeros_like(x)
            for z in (1, 3, 4, 5):
                out[zone == z] = getattr(self, f"{f}{z}")
            setattr(self, f + "_field", out)
        fan = zone == 2
        if fan.any():
            xf = x[fan]
            v2 = self.calc_velocity2(xf, t)
            self.velocity_field[fan] = v2
            self.density_field[fan] = self.calc_density2(xf, t)
            self.pressure_field[fan] = self.calc_pressure2(xf, t)
            self.temperature_field[fan] = self.calc_temperature2(xf, t, v2)
            self.internal_energy_field[fan] = self.calc_internal_energy(self.pressure_field[fan], self.density_field[fan])
            self.entropy_field[fan] = self.calc_entropy(self.pressure_field[fan], self.density_field[fan])

    # -- the numerical solver ---------------------------------------------------------------------------------------
    def build_numerical(self, xmin, xmax, ncoord, time_increment=0.05, xdiaphragm=0.0, keep_edge=False):
        if self.gamma is None:
            raise ValueError("constants are not set; call build_constant()")
        svr = self.svr = Euler1DSolver(xmin, xmax, ncoord, time_increment=time_increment, keep_edge=keep_edge)
        svr.gamma.fill(self.gamma)
        left = svr.coord < xdiaphragm
        right = np.logical_not(left)
        svr.so0[left, 0] = self.density1
        svr.so0[right, 0] = self.density5
        svr.so0[:, 1] = 0.0
        svr.so0[left, 2] = svr.calc_u2(self.gamma, self.density1, 0.0, self.pressure1)
        svr.so0[right, 2] = svr.calc_u2(self.gamma, self.density5, 0.0, self.pressure5)
        svr.so1.fill(0)
        svr.setup_march()
