"""``solvcon_b200.gradient_element.GradientElement`` (the Python-visible probe of the drop-in surface)
against the reference's own class through oracle/_ref, bit for bit, on every cell of every test mesh; plus the
structural and bounds checks of the reference's tests/test_multidim.py:284-420.  CPU only."""
import numpy as np
import pytest

from cases import md_mesh_case

MESHES = ["quad2d", "tri2d", "mixed2d", "hex3d", "tet3d", "prism3d", "mixed3d", "pyramid3d", "one_hex", "six_pyramid"]


class _Arr:
    """ndarray + nghost, the two things GradientElement reads from a SimpleArray."""

    def __init__(self, a, nghost):
        self.ndarray, self.nghost = a, nghost


def _setup(bridge, name):
    if not bridge.have_ref():
        pytest.skip("needs oracle/_ref")
    from solvcon_b200.mesh import TableMesh
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    rmesh = bridge.make_ref_mesh(ndim, ndcrd, cltpn, clnds)
    ref = bridge.RefEuler(rmesh, 1.e-3)            # the constructor runs prepare_ce
    mesh = TableMesh(ndim, rmesh.tables())
    cecnd = _Arr(ref.arr("cecnd").copy(), rmesh.ngstcell)
    return ndim, rmesh, ref, mesh, cecnd


@pytest.mark.parametrize("name", MESHES)
def test_matches_reference_bit_for_bit(bridge, name):
    from solvcon_b200.gradient_element import GradientElement
    ndim, rmesh, ref, mesh, cecnd = _setup(bridge, name)
    udf = [0.3, -1.25, 0.7][:ndim]
    cells = range(rmesh.ncell) if rmesh.ncell <= 300 else range(0, rmesh.ncell, 7)
    for icl in cells:
        for tau in (1.0, 0.0, 0.37):
            want = ref.gradient_element(icl, tau, udf)
            ge = GradientElement(mesh=mesh, cecnd=cecnd, icl=icl, tau=tau)
            assert [ge.icl, ge.ndim, ge.clnfc, ge.nfge] == want["meta"].tolist()
            assert ge.nfge_inverse == want["nfge_inverse"]
            for ifl in range(ge.clnfc):
                assert ge.rcl(ifl) == want["rcl"][ifl]
                for d in range(ndim):
                    assert ge.idis(ifl, d) == want["idis"][ifl, d], (icl, tau, ifl, d)
                    assert ge.jdis(ifl, d) == want["jdis"][ifl, d], (icl, tau, ifl, d)
            for ifge in range(ge.nfge):
                assert ge.faces(ifge) == want["faces"][ifge, :ndim].tolist()
                assert np.array_equal(np.array(ge.displacement_matrix(ifge)), want["dmat"][ifge, :ndim, :ndim])
                assert ge.solve_gradient(ifge, udf) == want["grads"][ifge, :ndim].tolist(), (icl, tau, ifge)


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d", "pyramid3d"])
def test_structure_and_bounds(bridge, name):
    """tests/test_multidim.py:291-347 (properties, accessor bounds, FGE table) and :386-419 (linear field)."""
    from solvcon_b200.gradient_element import GradientElement
    ndim, rmesh, ref, mesh, cecnd = _setup(bridge, name)
    ge = GradientElement(mesh=mesh, cecnd=cecnd, icl=0, tau=1.0)
    assert ge.icl == 0 and ge.ndim == ndim and ge.clnfc == mesh.clfcs.ndarray[mesh.clfcs.nghost, 0]
    nfc, nd = ge.clnfc, ge.ndim
    for ifl in (-1, nfc):
        for fn in (lambda: ge.rcl(ifl), lambda: ge.idis(ifl, 0), lambda: ge.jdis(ifl, 0)):
            with pytest.raises(IndexError, match=rf"ifl {ifl} out of range \[0, {nfc}\)"):
                fn()
    for d in (-1, nd):
        with pytest.raises(IndexError, match=rf"d {d} out of range \[0, {nd}\)"):
            ge.idis(0, d)
        with pytest.raises(IndexError):
            ge.jdis(0, d)
    assert ge.nfge > 0 and abs(ge.nfge_inverse - 1.0 / ge.nfge) < 1e-15
    for ifge in range(ge.nfge):
        f = ge.faces(ifge)
        assert len(f) == nd and all(1 <= x <= nfc for x in f)
    for ifge in (-1, ge.nfge):
        for fn in (lambda: ge.faces(ifge), lambda: ge.displacement_matrix(ifge), lambda: ge.solve_gradient(ifge, [0.0] * nd)):
            with pytest.raises(IndexError, match=rf"ifge {ifge} out of range"):
                fn()
    with pytest.raises(ValueError, match=rf"udf size {nd + 1} must equal ndim {nd}"):
        ge.solve_gradient(0, [0.0] * (nd + 1))
    grad = np.array([1.5, -2.7, 0.9][:nd])
    for icl in range(min(rmesh.ncell, 40)):
        g = GradientElement(mesh=mesh, cecnd=cecnd, icl=icl, tau=1.0)
        for ifge in range(g.nfge):
            dst = np.array(g.displacement_matrix(ifge))
            assert abs(np.linalg.det(dst)) > 1e-10
            got = np.array(g.solve_gradient(ifge, (dst @ grad).tolist()))
            assert np.allclose(got, grad, rtol=0, atol=1e-9)
