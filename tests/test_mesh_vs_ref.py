"""The native StaticMesh builder (csrc/scg_hmesh.cpp) against the reference's StaticMesh compiled from
/root/reference (oracle/_ref): every table identical, element for element, and the golden vectors of
the reference's own mesh tests (tests/test_mesh.py:84-182)."""
import numpy as np
import pytest

from cases import md_mesh_case

MESHES = ["quad2d", "tri2d", "mixed2d", "step2d", "hex3d", "tet3d", "prism3d", "mixed3d", "pyramid3d"]


def _build(ndim, ndcrd, cltpn, clnds, bcs=()):
    from solvcon_b200.mesh import make_mesh
    return make_mesh(ndim, ndcrd, cltpn, clnds, bcs, build_edge=True)


@pytest.mark.parametrize("name", MESHES)
def test_tables_identical_to_reference(refbridge, name):
    b = refbridge
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    ref = b.make_ref_mesh(ndim, ndcrd, cltpn, clnds)
    our = _build(ndim, ndcrd, cltpn, clnds)
    for cnt in ("nnode", "nface", "ncell", "nbound", "ngstnode", "ngstface", "ngstcell"):
        assert getattr(our, cnt) == getattr(ref, cnt), cnt
    for tab in b.MESH_TABLES:
        ra, rg = ref.table(tab)
        oa = getattr(our, tab)
        assert oa.nghost == rg, tab
        assert oa.ndarray.shape == ra.shape, (tab, oa.ndarray.shape, ra.shape)
        if tab == "clgrp":
            continue  # the reference leaves clgrp uninitialised
        assert np.array_equal(oa.ndarray, ra), f"{tab} differs"


def test_named_bcs_match_reference(refbridge):
    b = refbridge
    from solvcon_b200 import meshgen
    ndim, ndcrd, cltpn, clnds = md_mesh_case("mixed3d")
    probe = _build(ndim, ndcrd, cltpn, clnds)
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(ndim, probe.fccnd.body, probe.fccls.body, probe.nface, lo, hi)
    bcs = [("inlet", inlet), ("outlet", outlet)]
    ref = b.make_ref_mesh(ndim, ndcrd, cltpn, clnds, bcs)
    our = _build(ndim, ndcrd, cltpn, clnds, bcs)
    assert np.array_equal(our.bndfcs.ndarray, ref.table("bndfcs")[0])
    for tab in ("fccls", "clnds", "clfcs", "fcnds", "ndcrd", "clcnd", "clvol", "fcnml"):
        assert np.array_equal(getattr(our, tab).ndarray, ref.table(tab)[0]), tab
    assert our.nbcs == 3 and [x.name for x in our.bcs] == ["inlet", "outlet", "<NONAME>"]
    assert our.bc("inlet").nbound == len(inlet)
    with pytest.raises(KeyError):
        our.bc("nope")


def test_three_triangle_fan_golden():
    """Golden vectors of the reference's tests/test_mesh.py:84-119."""
    from solvcon_b200.mesh import StaticMesh
    mh = StaticMesh(ndim=2, nnode=4, nface=0, ncell=3)
    mh.ndcrd.ndarray[:, :] = (0, 0), (-1, -1), (1, -1), (0, 1)
    mh.cltpn.ndarray[:] = 4
    mh.clnds.ndarray[:, :4] = (3, 0, 1, 2), (3, 0, 2, 3), (3, 0, 3, 1)
    mh.build_interior()
    mh.build_boundary()
    mh.build_ghost()
    assert (mh.nnode, mh.nface, mh.ncell, mh.nbound) == (4, 6, 3, 3)
    assert (mh.ngstnode, mh.ngstface, mh.ngstcell) == (3, 6, 3)
    assert mh.nedge == 6
    assert mh.bndfcs.ndarray.tolist() == [[1, 0, -1], [3, 0, -1], [5, 0, -1]]
    np.testing.assert_almost_equal(mh.fccnd.body, [[-0.5, -0.5], [0, -1], [0.5, -0.5], [0.5, 0], [0, 0.5], [-0.5, 0]])
    np.testing.assert_almost_equal(mh.fcara.body, [np.sqrt(2), 2, np.sqrt(2), np.sqrt(5), 1, np.sqrt(5)])
    np.testing.assert_almost_equal(mh.clvol.body, [1, 0.5, 0.5])
    np.testing.assert_almost_equal(mh.clcnd.body, [[0, -2.0 / 3], [1.0 / 3, 0], [-1.0 / 3, 0]])


def test_single_tet_golden():
    """Golden vectors of the reference's tests/test_mesh.py:140-182."""
    from solvcon_b200.mesh import StaticMesh
    mh = StaticMesh(ndim=3, nnode=4, nface=0, ncell=1)
    mh.ndcrd.ndarray[:, :] = (0, 0, 0), (0, 1, 0), (-1, 1, 0), (0, 1, 1)
    mh.cltpn.ndarray[:] = 6
    mh.clnds.ndarray[:, :5] = [(4, 0, 1, 2, 3)]
    mh.build_interior()
    mh.build_boundary()
    mh.build_ghost()
    assert (mh.nface, mh.nbound, mh.ngstnode, mh.ngstface, mh.ngstcell) == (4, 4, 4, 12, 4)
    np.testing.assert_almost_equal(mh.clcnd.body, [[-0.25, 0.75, 0.25]])
    np.testing.assert_almost_equal(mh.clvol.body, [1.0 / 6])
    np.testing.assert_almost_equal(mh.fcara.body, [0.5, 0.5, np.sqrt(3) / 2, 0.5])


def test_add_bc_rules():
    """BC-group rules of the reference (tests/test_mesh.py:212-318; StaticMesh_boundary.cpp:14-78)."""
    from solvcon_b200.mesh import make_mesh, StaticMesh
    ndim, ndcrd, cltpn, clnds = md_mesh_case("quad2d")
    m = StaticMesh(ndim, len(ndcrd), 0, len(cltpn))
    m.ndcrd.ndarray[...] = ndcrd
    m.cltpn.ndarray[...] = cltpn
    m.clnds.ndarray[...] = clnds
    m.build_interior()
    bnd = np.nonzero(m.fccls.body[:, 1] < 0)[0]
    inner = np.nonzero(m.fccls.body[:, 1] >= 0)[0]
    m.add_bc("a", bnd[:3])
    with pytest.raises(ValueError, match="duplicated name"):
        m.add_bc("a", bnd[3:5])
    with pytest.raises(ValueError, match="already claimed"):
        m.add_bc("b", bnd[2:4])
    with pytest.raises(ValueError, match="not a boundary face"):
        m.add_bc("c", inner[:1])
    with pytest.raises(ValueError, match="out of range"):
        m.add_bc("d", [m.nface])
    with pytest.raises(ValueError, match="empty name"):
        m.add_bc("", bnd[5:6])
    m.build_boundary()
    with pytest.raises(RuntimeError, match="already built"):
        m.build_boundary()
    with pytest.raises(RuntimeError, match="already built"):
        m.add_bc("e", bnd[6:7])
    assert m.nbcs == 2 and m.bc(1).name == "<NONAME>" and m.bc(0).nbound + m.bc(1).nbound == len(bnd)
