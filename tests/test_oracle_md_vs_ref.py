"""The plain-C multidim oracle (oracle/cese_oracle.c) pinned bit-for-bit against the UNMODIFIED
reference EulerCore compiled from /root/reference (oracle/_ref), phase by phase and after a march,
on 2D tri/quad/mixed and 3D hex/tet/prism/mixed meshes with all three boundary kinds.  CPU only."""
import numpy as np
import pytest

from cases import md_bump_ic, md_mesh_case

MESHES = ["quad2d", "tri2d", "mixed2d", "step2d", "hex3d", "tet3d", "prism3d", "mixed3d", "pyramid3d"]


def build_pair(b, name, dt_scale=0.05):
    from solvcon_b200 import meshgen
    ndim, ndcrd, cltpn, clnds = md_mesh_case(name)
    mesh = b.make_ref_mesh(ndim, ndcrd, cltpn, clnds)
    tabs = mesh.tables()
    nface = mesh.nface
    fccnd = mesh.body("fccnd")
    fccls = mesh.body("fccls")
    lo, hi = ndcrd.min(axis=0), ndcrd.max(axis=0)
    inlet, outlet, wall = meshgen.boundary_groups(ndim, fccnd, fccls, nface, lo, hi)
    h = (np.prod(hi - lo) / len(cltpn)) ** (1.0 / ndim)
    dt = dt_scale * h
    r = b.RefEuler(mesh, dt)
    p = b.PortEuler(tabs, ndim, dt)
    for s in (r, p):
        s.init_solution(1.4, 1.4, [2.0 * np.sqrt(1.4 / 1.4), 0.0, 0.0][:ndim], 1.0)
        md_bump_ic(s.body, mesh.body("clcnd"), ndim)
        s.add_bc(b.BC_INLET, inlet, [1.4] + [2.0, 0.0, 0.0][:ndim] + [1.0, 1.4])
        s.add_bc(b.BC_NONREFL, outlet)
        s.add_bc(b.BC_SLIPWALL, wall)
    # the port computes its own CE geometry (prepare_ce), the reference did so in its constructor
    p.phase(b.MD_PREPARE_CE)
    return mesh, r, p, (inlet, outlet, wall)


def assert_same(r, p, tag, names=None):
    from oracle import bridge as b
    for n in names or b.MD_ARRAYS:
        assert np.array_equal(r.arr(n), p.arr(n)), f"{tag}: {n} differs, max abs {np.abs(r.arr(n) - p.arr(n)).max()}"


@pytest.mark.parametrize("name", MESHES)
def test_md_geometry_and_phases_bit_equal(refbridge, name):
    b = refbridge
    mesh, r, p, _ = build_pair(b, name)
    assert mesh.body("clvol").min() > 0
    assert_same(r, p, "prepare_ce", ("cevol", "cecnd", "sfcnd", "sfnml"))
    for s in (r, p):
        s.phase(b.MD_BC_SOLN)
        s.phase(b.MD_BC_DSOLN)
    assert_same(r, p, "prime")
    for it in range(3):
        for ph in (b.MD_UPDATE, b.MD_SOLT, b.MD_SOLN, b.MD_BC_SOLN, b.MD_CFL, b.MD_DSOLN, b.MD_BC_DSOLN):
            r.phase(ph)
            p.phase(ph)
            assert_same(r, p, f"substep {it} phase {ph}")
    assert np.isfinite(r.arr("so0n")).all()
    assert np.abs(r.body("so1n")).max() > 1e-3  # gradients actually developed


@pytest.mark.parametrize("name", ["mixed2d", "mixed3d"])
def test_md_march_bit_equal(refbridge, name):
    b = refbridge
    mesh, r, p, _ = build_pair(b, name)
    for s in (r, p):
        s.phase(b.MD_BC_SOLN)
        s.phase(b.MD_BC_DSOLN)
        s.march(25)
    assert_same(r, p, "march(25)")
    assert np.isfinite(r.arr("so0n")).all()
    assert r.body("cflo").max() < 1.0


def test_md_params_and_normal_matrix(refbridge):
    b = refbridge
    mesh, r, p, (inlet, outlet, wall) = build_pair(b, "mixed3d")
    for s in (r, p):
        s.set_params(2.0, 0.1, 0.7)
        s.phase(b.MD_BC_SOLN)
        s.phase(b.MD_BC_DSOLN)
        s.march(3)
    assert_same(r, p, "march with sigma0=2, taumin=.1, tauscale=.7")
    fcnml = mesh.body("fcnml")
    for ifc in list(wall[:5]) + list(inlet[:2]):
        out = np.zeros((3, 3))
        b.port_lib().orc_normal_matrix(3, b._p(np.ascontiguousarray(fcnml[ifc])), b._p(out))
        assert np.array_equal(out, r.normal_matrix(int(ifc)))
