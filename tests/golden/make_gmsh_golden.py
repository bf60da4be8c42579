"""Generate the Gmsh fixtures and their golden StaticMesh tables (run in the build container, where /root/reference
exists):  python tests/golden/make_gmsh_golden.py

Inputs (*.msh, written HERE by this script from our own generators -- not copies of the reference's data files):
  fan.msh      three triangles around a centre node, with boundary-line elements and physical names
  tq.msh       two quadrilaterals and two triangles in a row
  step2d.msh   ~300-cell mixed tri/quad forward-facing step with boundary lines and corner points
  mix3d.msh    hexahedra + prisms + tetrahedra of the stepped channel (prism node order exercised)
Goldens (*.gmsh.npz): every table of ``Gmsh(data).to_block()`` of the UNMODIFIED reference reader
(cpp/solvcon/inout/gmsh.cpp compiled into oracle/_ref) -- except ``clgrp`` and the unused ``clnds`` slots, which the
reference leaves uninitialised."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import bridge as b          # noqa: E402
from solvcon_b200 import meshgen        # noqa: E402

GMSH_TYPE = {3: 3, 4: 2, 5: 5, 6: 4, 7: 6, 8: 7}      # solvcon cell type -> gmsh element type
PRISM = (0, 2, 1, 3, 5, 4)


def write_msh(path, ndcrd, cltpn, clnds, lines=(), points=(), names=()):
    nd = np.zeros((len(ndcrd), 3))
    nd[:, :ndcrd.shape[1]] = ndcrd
    out = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat"]
    if names:
        out += ["$PhysicalNames", str(len(names))] + [f'{d} {i} "{n}"' for d, i, n in names] + ["$EndPhysicalNames"]
    out += ["$Nodes", str(len(nd))] + [f"{i + 1} {x!r} {y!r} {z!r}" for i, (x, y, z) in enumerate(nd.tolist())] + ["$EndNodes"]
    els = []
    for p in points:
        els.append(f"15 2 1 1 {p + 1}")
    for tag, a, c in lines:
        els.append(f"1 2 {tag} {tag} {a + 1} {c + 1}")
    for t, row in zip(cltpn.tolist(), clnds.tolist()):
        nds = row[1:1 + row[0]]
        if t == 7:
            nds = [nds[PRISM.index(i)] for i in range(6)]
        els.append(f"{GMSH_TYPE[t]} 2 9 1 " + " ".join(str(n + 1) for n in nds))
    out += ["$Elements", str(len(els))] + [f"{i + 1} {e}" for i, e in enumerate(els)] + ["$EndElements", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))


def golden(name):
    data = open(os.path.join(HERE, name + ".msh"), "rb").read()
    rm = b.ref_gmsh_mesh(data)
    arrs = {"counts": np.array([rm.ndim, rm.nnode, rm.nface, rm.ncell, rm.nbound, rm.ngstnode, rm.ngstface, rm.ngstcell])}
    for t in b.MESH_TABLES:
        if t == "clgrp":
            continue
        a, ng = rm.table(t)
        a = a.copy()
        if t == "clnds":
            for r in range(len(a)):
                a[r, 1 + a[r, 0]:] = -1
        arrs[t] = a
    np.savez_compressed(os.path.join(HERE, name + ".gmsh.npz"), **arrs)
    print(name, arrs["counts"])


def main():
    # fan
    nd = np.array([[0.25, 0.1], [-1.0, -1.25], [1.5, -1.0], [0.0, 1.75]])
    cl = np.array([[3, 0, 1, 2] + [-1] * 5, [3, 0, 2, 3] + [-1] * 5, [3, 0, 3, 1] + [-1] * 5], dtype=np.int32)
    write_msh(os.path.join(HERE, "fan.msh"), nd, np.array([4, 4, 4]), cl, lines=[(1, 1, 2), (2, 2, 3), (3, 3, 1)],
              names=[(1, 1, "south"), (1, 2, "east"), (1, 3, "west"), (2, 9, "domain")])
    # tq
    nd = np.array([[0, 0], [1, 0], [2.25, 0], [3, 0], [0, 1], [1, 1.5], [2, 1], [3, 1.25]], dtype=float)
    cl = np.full((4, 9), -1, dtype=np.int32)
    cl[0, :5] = [4, 0, 1, 5, 4]
    cl[1, :5] = [4, 1, 2, 6, 5]
    cl[2, :4] = [3, 2, 3, 6]
    cl[3, :4] = [3, 3, 7, 6]
    write_msh(os.path.join(HERE, "tq.msh"), nd, np.array([3, 3, 4, 4]), cl)
    # step2d
    nd, ct, cn = meshgen.mixed_2d(30, 10, lx=3.0, ly=1.0, tri_every=4, jitter=0.3, step=(0.6, 0.2))
    write_msh(os.path.join(HERE, "step2d.msh"), nd, ct, cn, lines=[(1, 0, 11), (1, 11, 22)], points=[0, 5],
              names=[(1, 1, "inlet"), (2, 9, "fluid")])
    # mix3d
    nd, ct, cn = meshgen.mixed_3d(3, 2, 4, lx=0.75, ly=0.5, lz=1.0, x1=2, z1=3, jitter=0.1)
    write_msh(os.path.join(HERE, "mix3d.msh"), nd, ct, cn, names=[(3, 9, "fluid")])
    for n in ("fan", "tq", "step2d", "mix3d"):
        golden(n)


if __name__ == "__main__":
    main()
