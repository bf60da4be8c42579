"""Gmsh MSH 2.2 reader (solvcon_b200/gmsh.py) against golden tables produced by the UNMODIFIED
reference reader (tests/golden/make_gmsh_golden.py) and against the known answers of the reference's own
tests/test_gmsh.py.  CPU only: the reader and the native StaticMesh builder need no GPU."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
TABLES = ("ndcrd", "fccnd", "fcnml", "fcara", "clcnd", "clvol", "fctpn", "cltpn", "fcnds", "fccls", "clnds", "clfcs",
          "ednds", "bndfcs")


def _read(name):
    with open(os.path.join(GOLD, name + ".msh"), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", ["fan", "tq", "step2d", "mix3d"])
def test_gmsh_tables_match_reference_reader(name):
    from solvcon_b200.gmsh import Gmsh
    m = Gmsh(_read(name)).to_block()
    g = np.load(os.path.join(GOLD, name + ".gmsh.npz"))
    counts = [m.ndim, m.nnode, m.nface, m.ncell, m.nbound, m.ngstnode, m.ngstface, m.ngstcell]
    assert counts == g["counts"].tolist()
    for t in TABLES:
        a = getattr(m, t).ndarray.copy()
        if t == "clnds":
            for r in range(len(a)):
                a[r, 1 + a[r, 0]:] = -1       # the reference leaves the unused slots uninitialised
        assert a.shape == g[t].shape and np.array_equal(a, g[t]), t


def test_gmsh_known_answers_of_the_reference_tests():
    """tests/test_gmsh.py:19-42 of the reference: a fan of three triangles around node 0."""
    from solvcon_b200.gmsh import Gmsh
    blk = Gmsh(_read("fan")).to_block()
    assert (blk.ndim, blk.nnode, blk.ngstnode, blk.ncell) == (2, 4, 3, 3)
    assert blk.cltpn.ndarray[blk.ngstcell:].tolist() == [4, 4, 4]
    assert blk.clnds.ndarray[blk.ngstcell:, :4].tolist() == [[3, 0, 1, 2], [3, 0, 2, 3], [3, 0, 3, 1]]
    np.testing.assert_almost_equal(blk.ndcrd.ndarray[blk.ngstnode:], [[0.25, 0.1], [-1.0, -1.25], [1.5, -1.0], [0.0, 1.75]])
    # only the top-dimension elements are cells (gmsh.cpp:68-80): the three boundary lines are not
    assert blk.nbound == 3 and blk.ngstcell == 3


def test_gmsh_prism_node_order_and_3d_volume():
    from solvcon_b200.gmsh import Gmsh
    blk = Gmsh(_read("mix3d")).to_block()
    vol = blk.clvol.ndarray[blk.ngstcell:]
    assert (vol > 0).all()
    # 3 x 2 x 4 boxes of 0.25 x 0.25 x 0.25 minus the solid block x < x1, z >= z1 (2 x 2 x 1 boxes)
    assert abs(vol.sum() - (24 - 4) * 0.25 ** 3) < 1e-12
    assert sorted(set(blk.cltpn.ndarray[blk.ngstcell:].tolist())) == [5, 6, 7]


def test_gmsh_format_errors():
    from solvcon_b200.gmsh import Gmsh
    good = _read("tq").decode()
    with pytest.raises(ValueError, match="Incorrect msh file format"):
        Gmsh(good.split("$Elements")[0])                      # mandatory section missing
    with pytest.raises(ValueError, match="Incorrect msh file format"):
        Gmsh(good.replace("$Nodes", "$Knoten"))               # unknown keyword ends the parse
    with pytest.raises(ValueError, match="Incorrect msh file format"):
        Gmsh("$Nodes\n0\n$EndNodes\n" + good)                 # sections out of order
    with pytest.raises(ValueError, match="does not support msh file ver 4.1"):
        Gmsh(good.replace("2.2 0 8", "4.1 0 8"))
    assert Gmsh(good.replace("\n", "\r\n")).to_block().ncell == 4   # DOS line ends


def test_gmsh_against_reference_reader_directly(bridge):
    """Where the compiled reference is present: the same comparison on freshly read files."""
    if not bridge.have_ref():
        pytest.skip("oracle/_ref not built")
    from solvcon_b200.gmsh import Gmsh
    for name in ("tq", "step2d"):
        data = _read(name)
        rm = bridge.ref_gmsh_mesh(data)
        m = Gmsh(data).to_block()
        for t in ("fccnd", "fcnml", "clcnd", "clvol", "fccls", "clfcs", "fcnds", "bndfcs"):
            ra, _ = rm.table(t)
            assert np.array_equal(ra, getattr(m, t).ndarray), t
