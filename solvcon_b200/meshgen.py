"""Deterministic synthetic meshes (nodes + cells) for the named benchmark configurations.

RNG-free like the reference's own mesher (``solvcon/pilot/apps/obsrefl/_driver.py:149-153``):
every choice is a closed-form hash of the lattice indices.  The output is the input the
reference's ``StaticMesh`` takes (``ndcrd``, ``cltpn``, ``clnds`` -- ``tests/test_multidim.py:152-281``
shows the filling pattern) plus named boundary-face selectors expressed on cell/face geometry.

Cell-type ids and node orders follow ``cpp/solvcon/mesh/StaticMesh.hpp`` (``CellType``):
3 quadrilateral, 4 triangle, 5 hexahedron, 6 tetrahedron, 7 prism, 8 pyramid; ``clnds`` rows are
``[nnode, n1, n2, ...]`` padded with -1 to CLMND+1 = 9 columns.
"""
from __future__ import annotations

import numpy as np

QUAD, TRI, HEX, TET, PRISM, PYRAMID = 3, 4, 5, 6, 7, 8
CLMND = 8


def _jitter2(i, j, amp):
    return amp * np.sin(12.9898 * i + 78.233 * j), amp * np.cos(26.651 * i + 41.347 * j)


def mixed_2d(nx, ny, lx=3.0, ly=1.0, tri_every=4, jitter=0.0, step=None):
    """``nx`` x ``ny`` boxes on ``[0,lx] x [0,ly]``; box (i,j) is a quadrilateral unless
    ``((i*73856093) ^ (j*19349663)) & (tri_every-1) == 0`` -- then two triangles along its
    ll-ur diagonal.  ``tri_every=0`` gives quads only, ``1`` triangles only.
    ``step=(x0, y1)`` removes the forward-facing step ``x >= x0, y < y1`` (box-aligned).
    Returns ``(ndcrd[nnode,2], cltpn[ncell], clnds[ncell,9])``."""
    ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    x = ii * (lx / nx)
    y = jj * (ly / ny)
    if jitter:
        jx, jy = _jitter2(ii, jj, jitter)
        inner = (ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)
        if step is not None:
            # keep the step's walls straight: no jitter on lattice lines touching the cut-out
            sx, sy = int(round(step[0] / lx * nx)), int(round(step[1] / ly * ny))
            inner &= ~((ii >= sx) & (jj <= sy))
        x = x + np.where(inner, jx, 0.0) * (lx / nx)
        y = y + np.where(inner, jy, 0.0) * (ly / ny)
    nid = (ii * (ny + 1) + jj).astype(np.int64)
    bi, bj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    bi, bj = bi.ravel(), bj.ravel()
    if step is not None:
        sx, sy = int(round(step[0] / lx * nx)), int(round(step[1] / ly * ny))
        keep = ~((bi >= sx) & (bj < sy))
        bi, bj = bi[keep], bj[keep]
    ll = nid[bi, bj]
    lr = nid[bi + 1, bj]
    ur = nid[bi + 1, bj + 1]
    ul = nid[bi, bj + 1]
    if tri_every == 0:
        split = np.zeros(len(bi), dtype=bool)
    elif tri_every == 1:
        split = np.ones(len(bi), dtype=bool)
    else:
        split = (((bi * 73856093) ^ (bj * 19349663)) & (tri_every - 1)) == 0
    ncl = np.where(split, 2, 1)
    first = np.concatenate(([0], np.cumsum(ncl)[:-1]))
    ncell = int(ncl.sum())
    cltpn = np.empty(ncell, dtype=np.int32)
    clnds = np.full((ncell, CLMND + 1), -1, dtype=np.int32)
    q = first[~split]
    cltpn[q] = QUAD
    clnds[q, 0] = 4
    clnds[q, 1:5] = np.stack([ll, lr, ur, ul], axis=1)[~split]
    t0 = first[split]
    cltpn[t0] = TRI
    cltpn[t0 + 1] = TRI
    clnds[t0, 0] = 3
    clnds[t0 + 1, 0] = 3
    clnds[t0, 1:4] = np.stack([ll, lr, ur], axis=1)[split]
    clnds[t0 + 1, 1:4] = np.stack([ll, ur, ul], axis=1)[split]
    ndcrd = np.stack([x.ravel(), y.ravel()], axis=1)
    if step is not None:
        ndcrd, clnds = _compact_nodes(ndcrd, clnds)
    return np.ascontiguousarray(ndcrd), cltpn, clnds


def _compact_nodes(ndcrd, clnds):
    used = np.zeros(len(ndcrd), dtype=bool)
    valid = clnds[:, 1:] >= 0
    used[clnds[:, 1:][valid]] = True
    remap = np.cumsum(used) - 1
    out = clnds.copy()
    out[:, 1:][valid] = remap[clnds[:, 1:][valid]]
    return ndcrd[used], out.astype(np.int32)


# Kuhn (Freudenthal) split of the unit box into 6 tetrahedra sharing the v000-v111 diagonal; the
# z-faces are cut along the ll-ur diagonal, matching the prisms' triangles.  Corner c = 4*dx+2*dy+dz.
_KUHN = []
for perm in ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)):
    c, path = 0, [0]
    for axis in perm:
        c += (4, 2, 1)[axis]
        path.append(c)
    _KUHN.append(path)


def _mixed_3d_native(dims, xyz, return_box_j, return_box):
    """The connectivity half of ``mixed_3d`` in native code on all host threads (csrc/scg_meshgen.cpp); the numpy code
    below stays the definition (tests/test_meshgen_native.py holds the two equal, array for array)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    nlat, ncell = C.c_int64(), C.c_int64()
    _lib.check(lib.scg_meshgen_mixed3d_count(*dims, C.byref(nlat), C.byref(ncell)))
    xyz = np.ascontiguousarray(xyz)
    assert len(xyz) == nlat.value
    cltpn = np.empty(ncell.value, dtype=np.int32)
    clnds = np.empty((ncell.value, CLMND + 1), dtype=np.int32)
    used = np.empty(nlat.value, dtype=np.uint8)
    want_box = return_box or return_box_j
    box = np.empty((4, ncell.value), dtype=np.int64) if want_box else None
    nn = C.c_int64()
    _lib.check(lib.scg_meshgen_mixed3d_cells(*dims, _lib.ptr(xyz), _lib.ptr(cltpn), _lib.ptr(clnds), used.ctypes.data,
                                             box.ctypes.data if want_box else None, C.byref(nn)))
    out = (np.ascontiguousarray(xyz[used.view(bool)]), cltpn, clnds)
    if return_box_j:
        out += (box[1].copy(),)
    if return_box:
        out += ((box[0], box[1], box[2], box[3]),)
    return out


def mixed_3d(nx, ny, nz, lx=None, ly=None, lz=None, x1=None, z1=None, kinds="hpt", jitter=0.0, j_range=None,
             return_box_j=False, box_range=None, return_box=False, native=True):
    """Stepped channel of ``nx`` x ``ny`` x ``nz`` boxes: block A hexahedra
    (``i < x1, k < z1``), block B z-axis prisms (``i >= x1, k < z1``; each box split on its vertical
    ll-ur diagonal plane), block C tetrahedra (``i >= x1, k >= z1``; Kuhn split, 6 per box); the
    region ``i < x1, k >= z1`` is solid.  All interfaces conform without pyramids (hex|prism share
    quads at ``i = x1``, prism|tet share triangles at ``k = z1``).  ``kinds`` selects the blocks:
    ``"h"`` = hex box only (``x1 = nx, z1 = nz``), ``"hp"``, ``"hpt"``; ``"t"`` / ``"p"`` = one kind
    everywhere.

    Sub-lattices (domain decomposition: no rank ever generates the global mesh): ``box_range=((i0, i1), (j0, j1),
    (k0, k1))`` generates only the boxes of that index range of the SAME global lattice -- coordinates, jitter, cell
    kinds and the relative order of nodes and cells are those of the full mesh (both are lexicographic in the global
    lattice indices).  ``j_range=(j0, j1)`` is the y-only form.  ``return_box_j`` also returns the global box index j
    of every cell; ``return_box`` returns ``(bi, bj, bk, sub)`` per cell (global box indices and the cell's position
    inside its box), from which ``cell_key`` makes an id that ascends with the global cell order.
    Returns ``(ndcrd[nnode,3], cltpn, clnds)``."""
    lx = float(nx) if lx is None else lx
    ly = float(ny) if ly is None else ly
    lz = float(nz) if lz is None else lz
    if kinds == "h":
        x1, z1 = nx, nz
    elif kinds in ("t", "p"):
        x1, z1 = 0, (0 if kinds == "t" else nz)
    elif kinds == "hp":
        x1 = nx // 2 if x1 is None else x1
        z1 = nz
    else:
        x1 = nx // 2 if x1 is None else x1
        z1 = nz // 2 if z1 is None else z1
    if box_range is None:
        box_range = ((0, nx), (0, ny) if j_range is None else j_range, (0, nz))
    (i0, i1), (j0, j1), (k0, k1) = [(max(0, int(a)), min(n, int(b))) for (a, b), n in zip(box_range, (nx, ny, nz))]
    ii, jj, kk = np.meshgrid(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), np.arange(k0, k1 + 1), indexing="ij")
    x = ii * (lx / nx)
    y = jj * (ly / ny)
    z = kk * (lz / nz)
    if jitter:
        inner = (ii > 0) & (ii < nx) & (jj > 0) & (jj < ny) & (kk > 0) & (kk < nz)
        inner &= ~((ii <= x1) & (kk >= z1))  # keep the solid block's walls planar
        x = x + np.where(inner, jitter * np.sin(12.9898 * ii + 78.233 * jj + 37.719 * kk), 0.0) * (lx / nx)
        y = y + np.where(inner, jitter * np.cos(26.651 * ii + 41.347 * jj + 11.135 * kk), 0.0) * (ly / ny)
        z = z + np.where(inner, jitter * np.sin(63.7264 * ii + 10.873 * jj + 53.127 * kk), 0.0) * (lz / nz)
    nxl, nyl, nzl = i1 - i0, j1 - j0, k1 - k0
    if native and (nxl + 1) * (nyl + 1) * (nzl + 1) < 2 ** 31:
        return _mixed_3d_native((nx, ny, nz, x1, z1, i0, i1, j0, j1, k0, k1), np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1),
                                return_box_j, return_box)
    # local lattice id, same relative order as the global one
    nid = (((ii - i0) * (nyl + 1) + (jj - j0)) * (nzl + 1) + (kk - k0)).astype(np.int64)
    li, lj, lk = [a.ravel() for a in np.meshgrid(np.arange(nxl), np.arange(nyl), np.arange(nzl), indexing="ij")]
    bi, bj, bk = li + i0, lj + j0, lk + k0
    solid = (bi < x1) & (bk >= z1)
    li, lj, lk, bi, bj, bk = (a[~solid] for a in (li, lj, lk, bi, bj, bk))
    # box corners c = 4*dx + 2*dy + dz
    cor = np.stack([nid[li + dx, lj + dy, lk + dz] for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)], axis=1)
    is_hex = (bi < x1) & (bk < z1)
    is_pri = (bi >= x1) & (bk < z1)
    is_tet = (bi >= x1) & (bk >= z1)
    ncl = np.where(is_hex, 1, np.where(is_pri, 2, 6))
    first = np.concatenate(([0], np.cumsum(ncl)[:-1])) if len(ncl) else np.zeros(0, dtype=np.int64)
    ncell = int(ncl.sum())
    cltpn = np.empty(ncell, dtype=np.int32)
    clnds = np.full((ncell, CLMND + 1), -1, dtype=np.int32)
    # hexahedron: bottom 000,100,110,010 then top (StaticMesh_interior.cpp add_hexahedron)
    h = first[is_hex]
    cltpn[h] = HEX
    clnds[h, 0] = 8
    clnds[h, 1:9] = cor[is_hex][:, [0, 4, 6, 2, 1, 5, 7, 3]]
    # prisms: triangles (000,100,110) and (000,110,010) extruded in z
    p = first[is_pri]
    cp = cor[is_pri]
    for s, tri in enumerate(((0, 4, 6), (0, 6, 2))):
        cltpn[p + s] = PRISM
        clnds[p + s, 0] = 6
        clnds[p + s, 1:4] = cp[:, list(tri)]
        clnds[p + s, 4:7] = cp[:, [c + 1 for c in tri]]
    # tetrahedra
    t = first[is_tet]
    ct = cor[is_tet]
    for s, path in enumerate(_KUHN):
        cltpn[t + s] = TET
        clnds[t + s, 0] = 4
        clnds[t + s, 1:5] = ct[:, path]
    ndcrd = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    # positive orientation for the tetrahedra (swap two nodes where the triple product is negative)
    tets = np.nonzero(cltpn == TET)[0]
    if len(tets):
        n = clnds[tets, 1:5]
        a, b, c, d = (ndcrd[n[:, q]] for q in range(4))
        vol = np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a)
        neg = vol < 0
        n[neg, 2], n[neg, 3] = n[neg, 3].copy(), n[neg, 2].copy()
        clnds[tets, 1:5] = n
    ndcrd, clnds = _compact_nodes(ndcrd, clnds)
    out = (np.ascontiguousarray(ndcrd), cltpn, clnds)
    if return_box_j:
        out += (np.repeat(bj, ncl),)
    if return_box:
        sub = np.arange(ncell, dtype=np.int64) - np.repeat(first, ncl)
        out += ((np.repeat(bi, ncl), np.repeat(bj, ncl), np.repeat(bk, ncl), sub),)
    return out


def cell_key(box, ny, nz):
    """An int64 id per cell that ascends with the GLOBAL cell order of ``mixed_3d`` (lexicographic in the global box
    indices, then the position inside the box): equal on every rank that generates the cell."""
    bi, bj, bk, sub = box
    return ((bi.astype(np.int64) * ny + bj) * nz + bk) * 6 + sub


def boundary_groups(ndim, fccnd, fccls, nface, lo, hi, tol=1e-9):
    """Split the boundary faces (``fccls[:,1] < 0``) of a box-shaped domain into inlet (x = lo[0]),
    outlet (x = hi[0]) and wall (everything else, incl. the step / solid block) face lists.
    ``fccnd``/``fccls`` are BODY rows of the StaticMesh tables."""
    bnd = np.nonzero(fccls[:nface, 1] < 0)[0].astype(np.int32)
    xc = fccnd[bnd, 0]
    inlet = bnd[np.abs(xc - lo[0]) < tol]
    outlet = bnd[np.abs(xc - hi[0]) < tol]
    wall = bnd[(np.abs(xc - lo[0]) >= tol) & (np.abs(xc - hi[0]) >= tol)]
    return inlet, outlet, wall


def boundary_groups_of(mesh, lo, hi, tol=1e-9):
    """``boundary_groups`` of a built StaticMesh, reading only the boundary faces' rows (``StaticMesh.gather``): the same
    three face lists, in ascending face id."""
    if not hasattr(mesh, "gather") or getattr(mesh, "on_device", 0) == 0:
        return boundary_groups(mesh.ndim, mesh.fccnd.body, mesh.fccls.body, mesh.nface, lo, hi, tol)
    bnd = np.sort(mesh.bndfcs.ndarray[:, 0]).astype(np.int32)
    xc = mesh.gather("fccnd", bnd)[:, 0]
    inlet = bnd[np.abs(xc - lo[0]) < tol]
    outlet = bnd[np.abs(xc - hi[0]) < tol]
    wall = bnd[(np.abs(xc - lo[0]) >= tol) & (np.abs(xc - hi[0]) >= tol)]
    return inlet, outlet, wall


def pyramid_3d(nx, ny, nz, lx=None, ly=None, lz=None, jitter=0.0):
    """``nx`` x ``ny`` x ``nz`` boxes, each cut into six pyramids whose apex is the box centre and whose bases are
    the six box faces (the only conforming all-pyramid lattice; quadrilateral faces are shared between boxes,
    triangular ones inside a box).  Node order of a pyramid: base 0-3 counter-clockwise seen from the apex, apex 4
    (``StaticMesh_interior.cpp:261-305``).  With ``jitter`` the lattice nodes are displaced like in ``mixed_3d`` and the
    centres are the mean of the eight displaced corners, so the quadrilateral faces are not planar.
    Returns ``(ndcrd[nnode,3], cltpn, clnds)``."""
    lx = float(nx) if lx is None else lx
    ly = float(ny) if ly is None else ly
    lz = float(nz) if lz is None else lz
    ii, jj, kk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    x, y, z = ii * (lx / nx), jj * (ly / ny), kk * (lz / nz)
    if jitter:
        inner = (ii > 0) & (ii < nx) & (jj > 0) & (jj < ny) & (kk > 0) & (kk < nz)
        x = x + np.where(inner, jitter * np.sin(12.9898 * ii + 78.233 * jj + 37.719 * kk), 0.0) * (lx / nx)
        y = y + np.where(inner, jitter * np.cos(26.651 * ii + 41.347 * jj + 17.113 * kk), 0.0) * (ly / ny)
        z = z + np.where(inner, jitter * np.sin(9.173 * ii + 53.519 * jj + 61.271 * kk), 0.0) * (lz / nz)
    nid = (ii * (ny + 1) + jj) * (nz + 1) + kk
    corners = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    bi, bj, bk = (a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"))
    c = {(a, b, d): nid[bi + a, bj + b, bk + d] for a in (0, 1) for b in (0, 1) for d in (0, 1)}
    nbox = len(bi)
    centre = len(corners) + np.arange(nbox)
    cxyz = sum(corners[c[k]] for k in c) / 8.0
    ndcrd = np.concatenate([corners, cxyz], axis=0)
    # box faces as node loops whose right-hand normal points OUT of the box
    out = [((0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)), ((0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)),
           ((0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)), ((0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)),
           ((0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)), ((1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1))]
    clnds = np.full((6 * nbox, CLMND + 1), -1, dtype=np.int32)
    for f, (a, b, cc, d) in enumerate(out):
        rows = slice(f, None, 6)
        clnds[rows, 0] = 5
        # the pyramid's own base face is (1, 4, 3, 2) in 1-based local numbering and must point out of the box
        clnds[rows, 1], clnds[rows, 2], clnds[rows, 3], clnds[rows, 4] = c[a], c[d], c[cc], c[b]
        clnds[rows, 5] = centre
    cltpn = np.full(6 * nbox, PYRAMID, dtype=np.int32)
    return ndcrd, cltpn, clnds
