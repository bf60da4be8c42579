"""Shared deterministic test cases (initial conditions) for the parity tests."""
import numpy as np


def sod_ic(core, xmin=-1.0, xmax=1.0, gamma=1.4, p1=1.0, rho1=1.0, p5=0.1, rho5=0.125):
    """Sod shock tube IC, restating solvcon/onedim/euler1d.py:47-53,146-166 (ShockTube.build_numerical):
    coord = linspace, zero field, then left/right constant states with u2 = rho*(ie+ke)."""
    n = core.ncoord
    core.coord[...] = np.linspace(xmin, xmax, num=n)
    core.cfl.fill(0)
    core.gamma.fill(gamma)
    core.so0[...] = 0.0
    core.so1[...] = 0.0
    x = core.coord
    left = x < 0  # euler1d.py splits at the origin
    for mask, rho, p in ((left, rho1, p1), (~left, rho5, p5)):
        ie = 1.0 / (gamma - 1) * p / rho
        core.so0[mask, 0] = rho
        core.so0[mask, 1] = 0.0
        core.so0[mask, 2] = rho * (ie + 0.0)


def smooth_e1d_ic(core, xmin=0.0, xmax=4.0):
    """A smooth non-trivial 1D Euler state with a non-uniform grid and varying gamma."""
    n = core.ncoord
    s = np.linspace(0.0, 1.0, n)
    core.coord[...] = xmin + (xmax - xmin) * (s + 0.05 * np.sin(2 * np.pi * s) / (2 * np.pi))
    x = core.coord
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * x / (xmax - xmin))
    v = 0.3 * np.cos(2 * np.pi * x / (xmax - xmin))
    p = 1.0 + 0.1 * np.sin(4 * np.pi * x / (xmax - xmin))
    core.gamma[...] = 1.4 + 0.05 * np.cos(x)
    ga = core.gamma
    core.so0[:, 0] = rho
    core.so0[:, 1] = rho * v
    core.so0[:, 2] = p / (ga - 1.0) + 0.5 * rho * v * v
    core.so1[...] = 0.0
    core.so1[:, 0] = 0.2 * 2 * np.pi / (xmax - xmin) * np.cos(2 * np.pi * x / (xmax - xmin))
    core.cfl.fill(0)


def scalar_ic(x, so0, so1, cfl=None):
    """so0 = 1/2 sin x, so1 = 1/2 cos x on the even plane.  The reference
    leaves its arrays uninitialised, so every slot is written here (cfl too when given)."""
    so0[...] = 0.0
    so1[...] = 0.0
    if cfl is not None:
        cfl[...] = 0.0
    so0[2::2] = 0.5 * np.sin(x[2::2])
    so1[2::2] = 0.5 * np.cos(x[2::2])


# ------------------------------------------------------------------ multidim cases ---

def md_mesh_case(name):
    """Named small meshes (nodes + cells) for the multidim parity tests."""
    from solvcon_b200 import meshgen
    if name == "quad2d":
        return (2,) + meshgen.mixed_2d(12, 6, lx=3.0, ly=1.0, tri_every=0)
    if name == "tri2d":
        return (2,) + meshgen.mixed_2d(10, 5, lx=3.0, ly=1.0, tri_every=1, jitter=0.2)
    if name == "mixed2d":
        return (2,) + meshgen.mixed_2d(16, 8, lx=3.0, ly=1.0, tri_every=4, jitter=0.25)
    if name == "step2d":
        return (2,) + meshgen.mixed_2d(30, 10, lx=3.0, ly=1.0, tri_every=4, jitter=0.2, step=(0.6, 0.2))
    if name == "hex3d":
        return (3,) + meshgen.mixed_3d(5, 4, 3, kinds="h", jitter=0.15)
    if name == "tet3d":
        return (3,) + meshgen.mixed_3d(3, 3, 2, kinds="t", jitter=0.15)
    if name == "prism3d":
        return (3,) + meshgen.mixed_3d(4, 3, 3, kinds="p", jitter=0.15)
    if name == "mixed3d":
        return (3,) + meshgen.mixed_3d(6, 4, 4, kinds="hpt", jitter=0.15)
    # smallest meshes: every neighbour of every cell is a ghost, groups far smaller than a thread block
    if name == "one_quad":
        return (2,) + meshgen.mixed_2d(1, 1, lx=1.0, ly=1.0, tri_every=0)
    if name == "two_tri":
        return (2,) + meshgen.mixed_2d(1, 1, lx=1.0, ly=1.0, tri_every=1)
    if name == "one_hex":
        return (3,) + meshgen.mixed_3d(1, 1, 1, kinds="h")
    if name == "six_tet":
        return (3,) + meshgen.mixed_3d(1, 1, 1, kinds="t")
    if name == "two_prism":
        return (3,) + meshgen.mixed_3d(1, 1, 1, kinds="p")
    if name == "six_pyramid":
        return (3,) + meshgen.pyramid_3d(1, 1, 1)
    # several waves of CTAs per cell-type group (148 SMs x 4 CTAs x 64 cells = 38k cells per wave)
    if name == "mixed3d_150k":
        return (3,) + meshgen.mixed_3d(48, 44, 48, kinds="hpt", jitter=0.15)
    if name == "mixed2d_200k":
        return (2,) + meshgen.mixed_2d(600, 280, lx=3.0, ly=1.0, tri_every=4, jitter=0.25, step=(0.6, 0.2))
    if name == "pyramid3d":
        return (3,) + meshgen.pyramid_3d(4, 3, 3, jitter=0.12)
    raise KeyError(name)


def md_bump_ic(core_body, clcnd_body, ndim, gamma=1.4, mach=2.0, amp=0.3):
    """Supersonic stream + Gaussian bump written into body rows of so0n (``core_body(name)`` returns
    the writable body view); gradients start at zero like init_solution leaves them."""
    c = clcnd_body
    ctr = c.mean(axis=0)
    span = c.max(axis=0) - c.min(axis=0)
    span = np.where(span > 0, span, 1.0)   # one-cell meshes
    r2 = (((c - ctr) / (0.25 * span)) ** 2).sum(axis=1)
    rho = 1.4 * (1.0 + amp * np.exp(-r2))
    p = 1.0 * (1.0 + amp * np.exp(-r2))
    v = np.zeros((len(c), ndim))
    v[:, 0] = mach * np.sqrt(gamma * 1.0 / 1.4)
    v[:, 1] = 0.1 * np.sin(3.0 * c[:, 0])
    so0 = core_body("so0n")
    so0[:, 0] = rho
    for d in range(ndim):
        so0[:, 1 + d] = rho * v[:, d]
    so0[:, ndim + 1] = p / (gamma - 1.0) + 0.5 * rho * (v * v).sum(axis=1)
