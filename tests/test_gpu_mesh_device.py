"""The StaticMesh table builder on the device (csrc/scg_hmesh_dev.cu: faces-from-cells dedup, metrics, boundary list,
ghost mirroring as GPU kernels) against the host builder, which tests/test_mesh_vs_ref.py holds
to the compiled reference table by table: every table, every count and the boundary bookkeeping bit for bit."""
import numpy as np
import pytest

from cases import md_mesh_case

pytestmark = pytest.mark.gpu

TABLES = ("ndcrd", "fccnd", "fcnml", "fcara", "clcnd", "clvol", "fctpn", "cltpn", "clgrp", "fcnds", "fccls", "clnds",
          "clfcs", "bndfcs")
COUNTS = ("nnode", "nface", "ncell", "nbound", "ngstnode", "ngstface", "ngstcell")


def _big(kind):
    from solvcon_b200 import meshgen
    if kind == "mixed3d_big":
        return (3,) + tuple(meshgen.mixed_3d(60, 24, 36, lx=2.5, ly=1.0, lz=1.5, x1=36, z1=18, kinds="hpt", jitter=0.1))
    if kind == "pyramid_big":
        return (3,) + tuple(meshgen.pyramid_3d(24, 20, 16, jitter=0.1))
    return (2,) + tuple(meshgen.mixed_2d(300, 200, jitter=0.3))


def _case(name):
    if name.endswith("_big"):
        return _big(name)
    return md_mesh_case(name)


def _same(a, b, what):
    for c in COUNTS:
        assert getattr(a, c) == getattr(b, c), f"{what}: {c} {getattr(a, c)} != {getattr(b, c)}"
    for n in TABLES:
        x, y = getattr(a, n), getattr(b, n)
        assert x.nghost == y.nghost, f"{what}: {n} ghost rows"
        assert x.ndarray.shape == y.ndarray.shape, f"{what}: {n} shape {x.ndarray.shape} != {y.ndarray.shape}"
        assert np.array_equal(x.ndarray, y.ndarray), f"{what}: table {n} differs"
    assert a.nbcs == b.nbcs
    for i in range(a.nbcs):
        assert a.bc(i).name == b.bc(i).name
        assert np.array_equal(a.bc(i).facn.ndarray, b.bc(i).facn.ndarray)


@pytest.mark.parametrize("name", ["quad2d", "tri2d", "mixed2d", "step2d", "hex3d", "tet3d", "prism3d", "mixed3d", "pyramid3d",
                                  "mixed3d_big", "pyramid_big", "mixed2d_big"])
def test_device_built_tables_equal_host_built(gpu, name):
    from solvcon_b200.mesh import make_mesh
    ndim, ndcrd, cltpn, clnds = _case(name)
    host = make_mesh(ndim, ndcrd, cltpn, clnds, device=False)
    dev = make_mesh(ndim, ndcrd, cltpn, clnds, device=True)
    assert host.on_device == 0 and dev.on_device == 2
    _same(dev, host, name)
    assert dev.on_device == 0, "tables were handed out as host views: the host copies are the mesh now"


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d"])
def test_device_interior_then_named_boundaries(gpu, name):
    """add_bc between build_interior and build_boundary: the interior tables come from the device, the named groups
    are checked against them, and the boundary / ghost stages give the host builder's tables."""
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import StaticMesh
    ndim, ndcrd, cltpn, clnds = _case(name)
    out = []
    for device in (False, True):
        m = StaticMesh(ndim, len(ndcrd), 0, len(cltpn)).set_device(device)
        m.ndcrd.ndarray[...] = ndcrd
        m.cltpn.ndarray[...] = cltpn
        m.clnds.ndarray[...] = clnds
        m.build_interior(True, False)
        assert m.on_device == (1 if device else 0)
        lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
        inlet, outlet, wall = meshgen.boundary_groups(ndim, m.fccnd.body, m.fccls.body, m.nface, lo, hi)
        m.add_bc("inlet", inlet)
        m.add_bc("outlet", outlet)
        with pytest.raises(ValueError):
            m.add_bc("again", inlet[:1])
        m.build_boundary()
        m.build_ghost()
        out.append(m)
    _same(out[1], out[0], name)


@pytest.mark.parametrize("name", ["mixed3d", "step2d"])
@pytest.mark.parametrize("order", ["morton", "rcm"])
def test_march_on_a_device_built_mesh(gpu, name, order):
    """A mesh that never leaves the device (builder -> layout stage -> march) marches like the host-built one."""
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import make_mesh
    from solvcon_b200.multidim import EulerCore
    ndim, ndcrd, cltpn, clnds = _case(name)
    fields = []
    for device in (False, True):
        mesh = make_mesh(ndim, ndcrd, cltpn, clnds, device=device)
        h = (np.prod(ndcrd.max(axis=0) - ndcrd.min(axis=0)) / len(cltpn)) ** (1.0 / ndim)
        svr = EulerCore(mesh, 0.05 * h, order=order)
        assert mesh.on_device == (2 if device else 0)
        vel = [2.0, 0.0, 0.0][:ndim]
        svr.init_solution(1.4, 1.4, vel, 1.0)
        lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
        inlet, outlet, wall = meshgen.boundary_groups(ndim, mesh.fccnd.body, mesh.fccls.body, mesh.nface, lo, hi)
        svr.add_inlet(inlet, [1.4] + vel + [1.0, 1.4])
        svr.add_nonrefl(outlet)
        svr.add_slipwall(wall)
        svr.march(6)
        fields.append({n: getattr(svr, n).ndarray.copy() for n in ("so0n", "so1n", "cflo", "cflc")})
        fields[-1]["perm"] = svr._dmesh.perm
    for n in fields[0]:
        assert np.array_equal(fields[0][n], fields[1][n]), f"{name} {order}: {n} differs between host-built and device-built mesh"


def test_gather_reads_rows_without_downloading_the_table(gpu):
    from solvcon_b200 import meshgen
    from solvcon_b200.mesh import make_mesh
    ndim, ndcrd, cltpn, clnds = _case("mixed3d_big")
    host = make_mesh(ndim, ndcrd, cltpn, clnds, device=False)
    dev = make_mesh(ndim, ndcrd, cltpn, clnds, device=True)
    rng = np.random.default_rng(3)
    for name, n, ng in (("fccnd", host.nface, host.ngstface), ("fccls", host.nface, host.ngstface), ("clcnd", host.ncell, host.ngstcell),
                        ("clvol", host.ncell, host.ngstcell), ("ndcrd", host.nnode, host.ngstnode)):
        idx = rng.integers(-ng, n, size=1000).astype(np.int32)
        want = getattr(host, name).ndarray[idx + ng].reshape(len(idx), -1)
        assert np.array_equal(dev.gather(name, idx), want), name
        assert np.array_equal(host.gather(name, idx), want), name
    with pytest.raises(IndexError):
        dev.gather("fccnd", [host.nface])
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    a = meshgen.boundary_groups_of(dev, lo, hi)
    b = meshgen.boundary_groups(ndim, host.fccnd.body, host.fccls.body, host.nface, lo, hi)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
