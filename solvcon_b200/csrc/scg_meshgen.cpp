// scg_meshgen.cpp -- the connectivity of the synthetic stepped channel (solvcon_b200/meshgen.py: mixed_3d), natively and
// on all host threads.  The Python generator stays the definition (tests/test_meshgen_native.py holds the two equal,
// array for array); at 5e7 cells its fancy-indexing takes 20 s, this takes about one.  Node COORDINATES are still
// evaluated by numpy (vectorised sin / cos of the lattice indices): only integers are produced here.
//
// Lattice: boxes (bi, bj, bk) of the global lattice nx x ny x nz, range [i0,i1) x [j0,j1) x [k0,k1); box kinds by index:
// hexahedron (bi < x1, bk < z1), two z-prisms (bi >= x1, bk < z1), six Kuhn tetrahedra (bi >= x1, bk >= z1), solid
// (bi < x1, bk >= z1: no cells).  Cells in lexicographic box order, then position in the box; local lattice node id
// ((ii - i0) * (nyl + 1) + (jj - j0)) * (nzl + 1) + (kk - k0); unused lattice nodes are dropped and the rest renumbered
// in order (`used` says which).  Cell types and node orders: cpp/solvcon/mesh/StaticMesh.hpp (CellType),
// StaticMesh_interior.cpp:61-330.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/scg.h"

namespace scg
{
int fail(char const * fmt, ...);
}

namespace
{

constexpr int HEX = 5, TET = 6, PRISM = 7, CLMND = 8;

void par_for(int64_t n, std::function<void(int64_t, int64_t)> const & fn)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 32) { nt = 32; }
    if (nt < 2 || n < 8)
    {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    int64_t const chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t)
    {
        int64_t const b = t * chunk, e = std::min<int64_t>(n, b + chunk);
        if (b >= e) { break; }
        th.emplace_back([=, &fn]() { fn(b, e); });
    }
    for (auto & t : th) { t.join(); }
}

struct Lattice
{
    int64_t nx, ny, nz, x1, z1, i0, i1, j0, j1, k0, k1;
    int64_t nxl() const { return i1 - i0; }
    int64_t nyl() const { return j1 - j0; }
    int64_t nzl() const { return k1 - k0; }
    int cells_of(int64_t bi, int64_t bk) const
    {
        if (bi < x1) { return bk < z1 ? 1 : 0; }
        return bk < z1 ? 2 : 6;
    }
    int64_t nid(int64_t li, int64_t lj, int64_t lk) const { return (li * (nyl() + 1) + lj) * (nzl() + 1) + lk; }
};

// Kuhn (Freudenthal) split: the six monotone corner paths 000 -> 111, corner c = 4 dx + 2 dy + dz, in the order of
// itertools-style permutations of the axes (meshgen.py: _KUHN)
void kuhn_paths(int (&path)[6][4])
{
    int const perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    int const bit[3] = {4, 2, 1};
    for (int p = 0; p < 6; ++p)
    {
        int c = 0;
        path[p][0] = 0;
        for (int a = 0; a < 3; ++a)
        {
            c += bit[perms[p][a]];
            path[p][a + 1] = c;
        }
    }
}

} // namespace

extern "C" {

int scg_meshgen_mixed3d_count(int64_t nx, int64_t ny, int64_t nz, int64_t x1, int64_t z1, int64_t i0, int64_t i1, int64_t j0,
                              int64_t j1, int64_t k0, int64_t k1, int64_t * nnode_lattice, int64_t * ncell)
{
    if (!(nnode_lattice && ncell)) { return scg::fail("scg_meshgen_mixed3d_count: null argument"); }
    if (!(0 <= i0 && i0 <= i1 && i1 <= nx && 0 <= j0 && j0 <= j1 && j1 <= ny && 0 <= k0 && k0 <= k1 && k1 <= nz))
    {
        return scg::fail("scg_meshgen_mixed3d_count: box range outside the lattice");
    }
    Lattice const L{nx, ny, nz, x1, z1, i0, i1, j0, j1, k0, k1};
    int64_t n = 0;
    for (int64_t li = 0; li < L.nxl(); ++li)
    {
        int64_t row = 0;
        for (int64_t lk = 0; lk < L.nzl(); ++lk) { row += L.cells_of(li + i0, lk + k0); }
        n += row * L.nyl();
    }
    *ncell = n;
    *nnode_lattice = (L.nxl() + 1) * (L.nyl() + 1) * (L.nzl() + 1);
    return 0;
}

// xyz: coordinates of the LOCAL lattice nodes [nnode_lattice][3] (tetrahedra are given positive orientation with them).
// Outputs: cltpn[ncell], clnds[ncell][9] (node ids already renumbered over the used nodes), used[nnode_lattice] (0/1),
// box[4][ncell] = (bi, bj, bk, position in the box) or null.
int scg_meshgen_mixed3d_cells(int64_t nx, int64_t ny, int64_t nz, int64_t x1, int64_t z1, int64_t i0, int64_t i1, int64_t j0,
                              int64_t j1, int64_t k0, int64_t k1, const double * xyz, int32_t * cltpn, int32_t * clnds,
                              uint8_t * used, int64_t * box, int64_t * nnode_used)
{
    if (!(xyz && cltpn && clnds && used && nnode_used)) { return scg::fail("scg_meshgen_mixed3d_cells: null argument"); }
    Lattice const L{nx, ny, nz, x1, z1, i0, i1, j0, j1, k0, k1};
    int64_t const nxl = L.nxl(), nyl = L.nyl(), nzl = L.nzl();
    int64_t const nlat = (nxl + 1) * (nyl + 1) * (nzl + 1);
    if (nlat >= (int64_t(1) << 31)) { return scg::fail("scg_meshgen_mixed3d_cells: more than 2^31 lattice nodes"); }
    // cell offset of box (li, lj, lk): slab_off[li] + lj * row[li] + rowpre[kind(li)][lk]
    std::vector<int64_t> slab_off(static_cast<size_t>(nxl) + 1, 0), row(static_cast<size_t>(nxl), 0);
    std::vector<int64_t> rowpre[2];
    for (int kind = 0; kind < 2; ++kind)
    {
        rowpre[kind].assign(static_cast<size_t>(nzl) + 1, 0);
        int64_t const bi = kind ? x1 : x1 - 1; // any index of that kind
        for (int64_t lk = 0; lk < nzl; ++lk) { rowpre[kind][lk + 1] = rowpre[kind][lk] + L.cells_of(bi, lk + k0); }
    }
    for (int64_t li = 0; li < nxl; ++li)
    {
        int const kind = (li + i0 >= x1) ? 1 : 0;
        row[li] = rowpre[kind][nzl];
        slab_off[li + 1] = slab_off[li] + row[li] * nyl;
    }
    int64_t const ncell = slab_off[nxl];
    int path[6][4];
    kuhn_paths(path);
    std::memset(used, 0, static_cast<size_t>(nlat));
    int64_t * const bbi = box, * const bbj = box ? box + ncell : nullptr, * const bbk = box ? box + 2 * ncell : nullptr,
            * const bsub = box ? box + 3 * ncell : nullptr;
    par_for(nxl, [&](int64_t lb, int64_t le)
            {
                for (int64_t li = lb; li < le; ++li)
                {
                    int64_t const bi = li + i0;
                    int const kind = (bi >= x1) ? 1 : 0;
                    for (int64_t lj = 0; lj < nyl; ++lj)
                    {
                        for (int64_t lk = 0; lk < nzl; ++lk)
                        {
                            int64_t const bk = lk + k0;
                            int const nc = L.cells_of(bi, bk);
                            if (0 == nc) { continue; }
                            int64_t const c0 = slab_off[li] + lj * row[li] + rowpre[kind][lk];
                            int32_t cor[8]; // c = 4 dx + 2 dy + dz
                            for (int dx = 0; dx < 2; ++dx)
                            {
                                for (int dy = 0; dy < 2; ++dy)
                                {
                                    for (int dz = 0; dz < 2; ++dz)
                                    {
                                        int64_t const id = L.nid(li + dx, lj + dy, lk + dz);
                                        cor[4 * dx + 2 * dy + dz] = static_cast<int32_t>(id);
                                        used[id] = 1; // benign race: every writer stores 1
                                    }
                                }
                            }
                            for (int s = 0; s < nc; ++s)
                            {
                                int64_t const c = c0 + s;
                                int32_t * const r = clnds + c * (CLMND + 1);
                                for (int q = 0; q <= CLMND; ++q) { r[q] = -1; }
                                if (1 == nc)
                                {
                                    // hexahedron: bottom 000,100,110,010 then top (StaticMesh_interior.cpp add_hexahedron)
                                    static int const h[8] = {0, 4, 6, 2, 1, 5, 7, 3};
                                    cltpn[c] = HEX;
                                    r[0] = 8;
                                    for (int q = 0; q < 8; ++q) { r[1 + q] = cor[h[q]]; }
                                }
                                else if (2 == nc)
                                {
                                    // prisms: triangles (000,100,110) and (000,110,010) extruded in z
                                    static int const tri[2][3] = {{0, 4, 6}, {0, 6, 2}};
                                    cltpn[c] = PRISM;
                                    r[0] = 6;
                                    for (int q = 0; q < 3; ++q)
                                    {
                                        r[1 + q] = cor[tri[s][q]];
                                        r[4 + q] = cor[tri[s][q] + 1];
                                    }
                                }
                                else
                                {
                                    cltpn[c] = TET;
                                    r[0] = 4;
                                    for (int q = 0; q < 4; ++q) { r[1 + q] = cor[path[s][q]]; }
                                    // positive orientation: swap two nodes where the triple product is negative
                                    double const * const a = xyz + int64_t(r[1]) * 3, * const b = xyz + int64_t(r[2]) * 3,
                                                 * const cc = xyz + int64_t(r[3]) * 3, * const d = xyz + int64_t(r[4]) * 3;
                                    double const u0 = b[0] - a[0], u1 = b[1] - a[1], u2 = b[2] - a[2];
                                    double const v0 = cc[0] - a[0], v1 = cc[1] - a[1], v2 = cc[2] - a[2];
                                    double const w0 = d[0] - a[0], w1 = d[1] - a[1], w2 = d[2] - a[2];
                                    double const x0 = u1 * v2 - u2 * v1, xx1 = u2 * v0 - u0 * v2, x2 = u0 * v1 - u1 * v0;
                                    double const vol = x0 * w0 + xx1 * w1 + x2 * w2;
                                    if (vol < 0) { std::swap(r[3], r[4]); } // the last two nodes, like meshgen.py
                                }
                                if (box)
                                {
                                    bbi[c] = bi;
                                    bbj[c] = lj + j0;
                                    bbk[c] = bk;
                                    bsub[c] = s;
                                }
                            }
                        }
                    }
                }
            });
    // renumber over the used lattice nodes, in order
    std::vector<int32_t> remap(static_cast<size_t>(nlat));
    int32_t nn = 0;
    for (int64_t q = 0; q < nlat; ++q)
    {
        remap[q] = nn;
        nn += used[q];
    }
    *nnode_used = nn;
    par_for(ncell, [&](int64_t b, int64_t e)
            {
                for (int64_t c = b; c < e; ++c)
                {
                    int32_t * const r = clnds + c * (CLMND + 1);
                    for (int q = 1; q <= r[0]; ++q) { r[q] = remap[r[q]]; }
                }
            });
    return 0;
}

} // extern "C"
