"""The plain-C oracle (oracle/cese_oracle.c) is pinned bit-for-bit against the UNMODIFIED reference
compiled from /root/reference (oracle/_ref).  CPU only."""
import numpy as np
import pytest

from cases import scalar_ic, smooth_e1d_ic, sod_ic


@pytest.mark.parametrize("alpha", [1, 2])
def test_e1d_sod_march_bit_equal(refbridge, alpha):
    b = refbridge
    n, dt, steps = 2001, 8.0e-4, 300
    r, p = b.RefEuler1D(n, dt), b.PortEuler1D(n, dt)
    sod_ic(r)
    sod_ic(p)
    r.march(alpha, steps)
    p.march(alpha, steps)
    for name in ("so0", "so1", "cfl"):
        assert np.array_equal(getattr(r, name), getattr(p, name)), name
    assert np.abs(r.so0[:, 1]).max() > 0.1  # the tube actually moved


def test_e1d_phases_bit_equal(refbridge):
    b = refbridge
    n, dt = 401, 2.0e-3
    r, p = b.RefEuler1D(n, dt), b.PortEuler1D(n, dt)
    smooth_e1d_ic(r)
    smooth_e1d_ic(p)
    seq = [(b.PH_SO0, 0), (b.PH_CFL, 0), (b.PH_SO1, 0), (b.PH_BC0, 0), (b.PH_BC1, 0),
           (b.PH_SO0, 1), (b.PH_CFL, 1), (b.PH_SO1, 1)]
    for _ in range(5):
        for ph, odd in seq:
            r.phase(ph, odd, 2)
            p.phase(ph, odd, 2)
            for name in ("so0", "so1", "cfl"):
                assert np.array_equal(getattr(r, name), getattr(p, name)), (name, ph, odd)


def test_e1d_even_ncoord_rejected(refbridge):
    with pytest.raises(ValueError, match="ncoord cannot be even"):
        refbridge.RefEuler1D(10, 0.1)
    with pytest.raises(ValueError, match="ncoord cannot be even"):
        refbridge.PortEuler1D(10, 0.1)


@pytest.mark.parametrize("eq", [0, 1])
@pytest.mark.parametrize("alpha", [0, 1, 2])
def test_scalar_march_bit_equal(refbridge, eq, alpha):
    b = refbridge
    n = 777
    dt = 0.5 * 8 * np.pi / n
    r = b.RefScalar1D(eq, 0.0, 8 * np.pi, n, dt)
    p = b.PortScalar1D(eq, 0.0, 8 * np.pi, n, dt)
    assert np.array_equal(r.x, p.x)
    scalar_ic(r.x, r.so0, r.so1, r.cfl)
    scalar_ic(p.x, p.so0, p.so1, p.cfl)
    r.march(alpha, 60)
    p.march(alpha, 60)
    for name in ("so0", "so1", "cfl"):
        assert np.array_equal(getattr(r, name), getattr(p, name)), name


def test_scalar_nonuniform_grid_phases(refbridge):
    b = refbridge
    xloc = np.cumsum(0.1 + 0.05 * np.sin(np.arange(65)))
    for eq in (0, 1):
        r = b.RefScalar1D(eq, dt=0.02, xloc=xloc)
        p = b.PortScalar1D(eq, dt=0.02, xloc=xloc)
        assert np.array_equal(r.x, p.x)
        scalar_ic(r.x, r.so0, r.so1, r.cfl)
        scalar_ic(p.x, p.so0, p.so1, p.cfl)
        seq = [(b.PH_SO0, 0), (b.PH_BC0, 0), (b.PH_CFL, 1), (b.PH_SO1, 0), (b.PH_BC1, 0),
               (b.PH_SO0, 1), (b.PH_CFL, 0), (b.PH_SO1, 1)]
        for _ in range(4):
            for ph, odd in seq:
                r.phase(ph, odd, 2)
                p.phase(ph, odd, 2)
                for name in ("so0", "so1", "cfl"):
                    assert np.array_equal(getattr(r, name), getattr(p, name)), (eq, name, ph, odd)
