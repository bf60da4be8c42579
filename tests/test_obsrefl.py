"""Oblique-shock-reflection driver (solvcon_b200/apps/obsrefl.py), host side: the mesher and the
analytic relations against goldens produced by the reference's OWN driver source (tests/golden/make_obsrefl_golden.py)
and against the known answers of the reference's tests/apps/obsrefl/test_driver.py.  CPU only."""
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obsrefl.npz")


@pytest.mark.parametrize("cell_type", ["quad", "triangle", "unstructured"])
def test_mesher_matches_reference_driver(cell_type):
    from solvcon_b200.apps.obsrefl import ObliqueShockMesher
    g = np.load(GOLD)
    msher = ObliqueShockMesher(nx=24, ny=8)
    mh = msher.make_mesh(cell_type=cell_type)
    assert np.array_equal(mh.ndcrd.body, g[cell_type + "_ndcrd"])
    assert np.array_equal(mh.clnds.body[:, :5], g[cell_type + "_clnds"])
    for name in msher.BOUNDARY_NAMES:
        assert np.array_equal(mh.bc(name).facn.ndarray[:, 0], g[cell_type + "_" + name]), name
    # test_driver.py:66-170: counter-clockwise cells tiling the 4 x 1 domain, at most one boundary face per cell
    # for the triangular flavours, nx / ny faces per edge
    assert (mh.clvol.body > 0).all() and abs(mh.clvol.body.sum() - 4.0) < 1e-12
    assert [len(mh.bc(n)) for n in msher.BOUNDARY_NAMES] == [8, 24, 24, 8]
    if cell_type != "quad":
        nb = np.bincount(mh.fccls.body[:mh.nface][mh.fccls.body[:mh.nface, 1] < 0, 0], minlength=mh.ncell)
        assert nb.max() == 1


def test_unknown_cell_type_and_missing_constants():
    from solvcon_b200.apps.obsrefl import ObliqueShock, ObliqueShockMesher
    with pytest.raises(ValueError, match="unknown cell_type 'hexagon'"):
        ObliqueShockMesher().make_mesh(cell_type="hexagon")
    with pytest.raises(ValueError, match="constants are not set"):
        ObliqueShock().build_numerical()
    with pytest.raises(ValueError, match="mesh is not built"):
        ObliqueShock().shock_path()


def test_relation_known_answers():
    """tests/apps/obsrefl/test_driver.py:265-316 (Anderson, Modern Compressible Flow, ch. 4)."""
    from solvcon_b200.apps.obsrefl import ObliqueShockRelation
    ob = ObliqueShockRelation(gamma=1.4)
    beta = math.radians(37.8)
    assert abs(ob.calc_density_ratio(3, beta) - 2.4204302545) < 1e-10
    assert abs(ob.calc_pressure_ratio(3, beta) - 3.7777114257) < 1e-10
    assert abs(ob.calc_temperature_ratio(3, beta) - 1.5607602899) < 1e-10
    assert abs(ob.calc_normal_dmach(3) - 0.4751909633) < 1e-10
    assert abs(ob.calc_dmach(3, beta=beta) - 1.9924827009) < 1e-10
    assert abs(ob.calc_dmach(3, theta=math.radians(20)) - 1.9941316656) < 1e-10
    b = ob.calc_shock_angle(4, math.radians(32), delta=1)
    assert abs(math.degrees(b) - 48.2584798722) < 1e-10
    assert abs(math.degrees(ob.calc_flow_angle(4, b)) - 32.0) < 1e-6
    with pytest.raises(ValueError):
        ob.calc_dmach(3)
    with pytest.raises(ValueError):
        ob.calc_dmach(3, beta=0.2, theta=0.1)
    with pytest.raises(ValueError, match="no attached shock"):
        ob.calc_shock_angle(2, math.radians(40))
    ob.gamma = 1.2
    assert abs(ob.calc_density_ratio(3, beta) - 2.7793244902) < 1e-10


def test_constants_match_reference_driver():
    from solvcon_b200.apps.obsrefl import ObliqueShock, ObliqueShockMesher
    g = np.load(GOLD)
    sh = ObliqueShock()
    sh.build_constant()
    assert np.array_equal(np.array(sh.zone_states()), g["zone_states"])
    assert np.array_equal(np.array([sh.theta, sh.shock_angle, sh.shock_angle2, sh.mach2, sh.mach3]), g["angles"])
    sh.mesher = ObliqueShockMesher()
    assert np.array_equal(np.array(sh.shock_path()), g["shock_path"])
    vx, vy = sh.velocity2
    assert vy < 0 and abs(math.atan2(-vy, vx) - math.radians(10.0)) < 1e-14
