"""``spacetime.Grid`` mirror against the known answers of the reference's tests/test_spacetime_grid.py:15-76 (error
texts, the coordinate vector with its mirrored ghosts, counts) and against the oracle's grid fill.  CPU only."""
import numpy as np
import pytest


def test_construction_errors():
    from solvcon_b200.spacetime import Grid
    with pytest.raises(ValueError, match=r"Grid::Grid\(xmin=0, xmax=10, ncelm=0\) invalid argument: ncelm smaller than 1"):
        Grid(0, 10, 0)
    with pytest.raises(ValueError, match=r"Grid::Grid\(xmin=10, xmax=10, ncelm=10\) invalid arguments: xmin >= xmax"):
        Grid(10, 10, 10)
    with pytest.raises(ValueError, match=r"Grid::Grid\(xmin=11, xmax=10, ncelm=10\) invalid arguments: xmin >= xmax"):
        Grid(11, 10, 10)
    Grid(xloc=np.arange(2) * 0.1)
    for s in (0, 1):
        with pytest.raises(ValueError, match=r"Grid::init_from_array\(xloc\) invalid arguments: xloc.size\(\)=%d smaller than 2" % s):
            Grid(xloc=np.arange(s) * 0.1)
    with pytest.raises(ValueError, match=r"Grid::init_from_array\(xloc\) invalid arguments: xloc\[0\]=1 >= xloc\[1\]=0.9"):
        Grid(xloc=np.arange(10, -1, -1) * 0.1)


def test_xcoord_golden_and_counts():
    from solvcon_b200.spacetime import Grid
    g = Grid(xmin=0.0, xmax=10.0, nelm=10)
    nx = (g.ncelm + g.BOUND_COUNT) * 2 + 1
    x = np.arange(0.0, 10.1, 0.5)
    front = x[0] - x[g.BOUND_COUNT:0:-1]
    back = (x[-1] - x[-2:-g.BOUND_COUNT - 2:-1]) + x[-1]
    golden = np.hstack([front, x, back])
    assert nx == len(g.xcoord) == 25
    assert golden.tolist() == g.xcoord.ndarray.tolist()
    assert (g.ncelm, g.nselm, g.xmin, g.xmax) == (10, 11, 0.0, 10.0)
    g.xcoord.ndarray.fill(10)                       # the view is live
    assert [10] * nx == g.xcoord.ndarray.tolist()


@pytest.mark.parametrize("n", [1, 7, 1000, 100_003])
def test_grid_equals_the_oracle(bridge, n):
    from solvcon_b200.spacetime import Grid
    xmax = 8 * np.pi
    chk = (bridge.RefScalar1D if bridge.have_ref() else bridge.PortScalar1D)(0, 0.0, xmax, n, 0.1)
    assert np.array_equal(Grid(0.0, xmax, n).xcoord.ndarray, chk.x)
    xloc = np.cumsum(0.1 + 0.05 * np.sin(np.arange(n + 1)))
    chk2 = bridge.PortScalar1D(0, dt=0.1, xloc=xloc)
    assert np.array_equal(Grid(xloc=xloc).xcoord.ndarray, chk2.x)
