"""The ``_solvcon`` shim (``dropin/_solvcon``) and the staging of the reference's Python package -- CPU side."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGE = os.path.join(ROOT, "dropin", "_stage")
needs_stage = pytest.mark.skipif(not os.path.isdir(os.path.join(STAGE, "solvcon")),
                                 reason="dropin/_stage not staged (dropin/stage.py needs /root/reference)")


def test_shim_serves_the_hot_path_names():
    import _solvcon
    import solvcon_b200
    from solvcon_b200 import gmsh, gradient_element, mesh, multidim, onedim, spacetime
    assert _solvcon.EulerCore is multidim.EulerCore and _solvcon.StaticMesh is mesh.StaticMesh
    assert _solvcon.StaticMeshBc is mesh.StaticMeshBc and _solvcon.GradientElement is gradient_element.GradientElement
    assert _solvcon.Gmsh is gmsh.Gmsh and _solvcon.onedim.Euler1DCore is onedim.Euler1DCore
    for n in ("Grid", "Celm", "Selm", "Kernel", "Solver", "LinearScalarSolver", "InviscidBurgersSolver"):
        assert getattr(_solvcon.spacetime, n) is getattr(spacetime, n)
    assert _solvcon.SimpleArrayFloat64 is solvcon_b200.SimpleArrayFloat64
    assert _solvcon.HAS_PILOT is False


def test_out_of_scope_names_are_inert():
    import _solvcon
    fft = _solvcon.FourierTransform
    assert "out of scope" in repr(fft)
    with pytest.raises(NotImplementedError, match="outside the CESE hot path"):
        fft()
    with pytest.raises(NotImplementedError):
        _solvcon.ProcessInfo.instance.command_line.populate()
    with pytest.raises(AttributeError):
        _solvcon.no_such_name
    with pytest.raises(ImportError):
        from _solvcon import pilot  # noqa: F401  (solvcon/pilot/_pilot_core.py:15-21 relies on this failing)


def test_point_and_segment_containers():
    import _solvcon
    pad = _solvcon.PointPadFp64(ndim=2)
    pad.append(0.5, 1.5)
    pad.append(2.0, 3.0)
    assert len(pad) == 2 and pad.x_at(1) == 2.0 and pad.y_at(0) == 1.5
    assert pad.pack_array().ndarray.tolist() == [[0.5, 1.5], [2.0, 3.0]]
    with pytest.raises(IndexError):
        pad.x_at(2)
    seg = _solvcon.SegmentPadFp64(ndim=2)
    seg.append(0.0, 1.0, 2.0, 3.0)
    assert len(seg) == 1 and (seg[0].x0, seg[0].y0, seg[0].x1, seg[0].y1) == (0.0, 1.0, 2.0, 3.0)
    assert seg.pack_array().ndarray.tolist() == [[0.0, 1.0, 2.0, 3.0]]
    with pytest.raises(IndexError):
        seg[1]


@needs_stage
def test_reference_package_imports_over_the_shim():
    import solvcon
    assert os.path.realpath(solvcon.__file__).startswith(os.path.realpath(STAGE))
    import solvcon_b200.mesh
    assert solvcon.StaticMesh is solvcon_b200.mesh.StaticMesh
    from solvcon.multidim.euler import EulerField   # noqa: F401
    from solvcon.pilot.apps.obsrefl import ObliqueShock, Reflection, ReflectionSession  # noqa: F401
    from solvcon.onedim import euler1d
    assert euler1d.ShockTube is not None and solvcon.spacetime.SolverProxy is not None


@needs_stage
def test_staged_files_are_verbatim():
    """Every staged file has the SHA-256 recorded at staging time, and -- where /root/reference is present -- that of
    the reference's file: nothing staged was edited."""
    with open(os.path.join(STAGE, "MANIFEST.json")) as f:
        man = json.load(f)
    assert len(man["package"]) > 50 and len(man["tests"]) >= 19

    def sha(p):
        with open(p, "rb") as fh:
            return hashlib.sha256(fh.read()).hexdigest()
    for rel, h in man["package"].items():
        assert sha(os.path.join(STAGE, rel)) == h, rel
        src = os.path.join(man["reference"], rel)
        if os.path.exists(src):
            assert sha(src) == h, rel
    for rel, h in man["tests"].items():
        assert sha(os.path.join(HERE, "_refsuite", rel[len("tests/"):])) == h, rel


def test_mesh_views_survive_a_rebuild():
    """A view of a StaticMesh table taken before build_boundary()/build_ghost() stays valid (it shows the old
    generation), as in the reference where a view keeps the old SimpleArray buffer alive."""
    from solvcon_b200.mesh import StaticMesh
    m = StaticMesh(ndim=2, nnode=4, nface=0, ncell=3)
    m.ndcrd.ndarray[:, :] = (0, 0), (-1, -1), (1, -1), (0, 1)
    m.cltpn.ndarray[:] = StaticMesh.TRIANGLE
    m.clnds.ndarray[:, :4] = (3, 0, 1, 2), (3, 0, 2, 3), (3, 0, 3, 1)
    m.build_interior()
    v = m.ndcrd.ndarray
    c = m.clcnd.ndarray
    keep_v, keep_c = v.copy(), c.copy()
    m.build_boundary()
    m.build_ghost()
    junk = [np.full(1000, 7.0) for _ in range(200)]   # churn the allocator
    assert m.ndcrd.ndarray.shape[0] > v.shape[0] and m.ndcrd.ndarray.ctypes.data != v.ctypes.data
    assert np.array_equal(v, keep_v) and np.array_equal(c, keep_c)
    v[...] = -1.0                                      # writing through the stale view touches only the old buffer
    assert np.array_equal(m.ndcrd.body, keep_v)
    del junk
