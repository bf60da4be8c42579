"""Analytic half of ``solvcon_b200.apps.shocktube.ShockTube`` against the known answers of the reference's
tests/test_onedim_euler.py:74-141 (Sod: gamma 1.4, p1 = rho1 = 1, p5 = 0.1, rho5 = 0.125).  CPU only."""
import numpy as np
import pytest


def _st():
    from solvcon_b200.apps.shocktube import ShockTube
    st = ShockTube()
    st.build_constant(gamma=1.4, pressure1=1.0, density1=1.0, pressure5=0.1, density5=0.125)
    return st


def test_known_answers():
    st = _st()
    np.testing.assert_allclose(st.calc_pressure45(), 3.0313017805064697, rtol=1e-7)
    np.testing.assert_allclose([st.density1, st.density3, st.density4, st.density5],
                               [1.0, 0.42631942817849494, 0.2655737117053072, 0.125], rtol=1e-7)
    assert st.pressure3 == st.pressure4 and st.velocity3 == st.velocity4
    np.testing.assert_allclose([st.pressure1, st.pressure3, st.pressure4, st.pressure5],
                               [1.0, 0.303130178050647, 0.303130178050647, 0.1], rtol=1e-7)
    np.testing.assert_allclose([st.velocity1, st.velocity3, st.velocity4, st.velocity5],
                               [0.0, 0.9274526200489503, 0.9274526200489503, 0.0], rtol=1e-7)
    np.testing.assert_allclose(st.calc_locations(t=0.1),
                               [-0.11832159566199232, -0.007027281256118345, 0.09274526200489504, 0.17521557320301784], rtol=1e-7)
    assert st.speedofsound1 == np.sqrt(1.4 * 1.0 / 1.0) and st.speedofsound5 == np.sqrt(1.4 * 0.1 / 0.125)


def test_field_at_t0_and_later():
    st = _st()
    st.build_field(t=0.0, coord=np.linspace(-1, 1, num=11))
    assert st.svr is None and len(st.coord_field) == 11
    assert st.density_field[:5].tolist() == [1.0] * 5 and st.density_field[5:].tolist() == [0.125] * 6
    assert st.pressure_field[:5].tolist() == [1.0] * 5 and st.pressure_field[5:].tolist() == [0.1] * 6
    assert (st.velocity_field == 0).all()
    x = np.linspace(-0.5, 0.5, 2001)
    st.build_field(t=0.2, coord=x)
    # continuity of the fan with zones 1 and 3, jumps at the contact (density only) and at the shock
    l12, l23, l34, l45 = st.calc_locations(0.2)
    i = np.searchsorted(x, l12)
    assert abs(st.density_field[i] - st.density1) < 2e-3 and abs(st.velocity_field[i]) < 2e-3
    j = np.searchsorted(x, l23) - 1
    assert abs(st.density_field[j] - st.density3) < 2e-3 and abs(st.pressure_field[j] - st.pressure3) < 2e-3
    k = np.searchsorted(x, l34)
    assert st.pressure_field[k - 1] == st.pressure_field[k] and st.density_field[k - 1] != st.density_field[k]
    assert np.all(np.diff(st.pressure_field) <= 1e-15)     # pressure never rises from left to right
    np.testing.assert_allclose(st.entropy_field[x < l34], st.entropy1, rtol=1e-12)   # isentropic up to the contact


def test_numerical_requires_constants():
    from solvcon_b200.apps.shocktube import ShockTube
    with pytest.raises(ValueError, match="constants are not set"):
        ShockTube().build_numerical(xmin=-1, xmax=1, ncoord=21)
