// scg_hmesh.cpp -- host-side unstructured mesh: the native twin of solvcon::StaticMesh
// (cpp/solvcon/mesh/StaticMesh.hpp:270-455, StaticMesh_interior.cpp, StaticMesh_boundary.cpp).
//
// The march only consumes the finished tables; this builder exists so the product is a drop-in for the
// reference's `StaticMesh(ndim, nnode, nface, ncell)` -> fill ndcrd/cltpn/clnds -> build_interior ->
// add_bc* -> build_boundary -> build_ghost sequence, and so that 10^7..10^8-cell synthetic meshes can be
// tabulated in seconds.  Table contents (numbering, node order inside faces, owner cell, normals,
// centroids, ghost mirroring) are IDENTICAL to the reference's: every floating-point expression keeps
// the reference's operation order and the file is built with -ffp-contract=off.
//
// Differences in *how* (not in what):
//   * duplicate faces are found by a counting sort on the smallest node id followed by an exact match
//     of the sorted node tuple inside each (tiny) bucket, instead of a node->faces hash table; roots are
//     kept in order of first appearance, so the face numbering is the same;
//   * build_boundary looks faces up through an index table instead of std::find (O(n) not O(n^2));
//   * face/cell metric loops that have no cross-iteration dependence run on all host threads.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/scg.h"

#include "scg_hmesh_dev.cuh"

namespace scg
{
int fail(char const * fmt, ...);
}

namespace
{

constexpr int FCMND = 4, CLMND = 8, CLMFC = 6, FCREL = 4, BFREL = 3;

void parallel_for(int64_t n, std::function<void(int64_t, int64_t)> const & fn)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 32) { nt = 32; }
    if (nt < 2 || n < 200000)
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

// Table storage: allocated uninitialised and filled / copied on all host threads (a std::vector would value-initialise
// on ONE thread: at 5e7 cells that is 14 GB of serial first-touch page faults).
template <typename T>
struct PBuf
{
    T * p = nullptr;
    size_t n = 0;
    PBuf() = default;
    PBuf(PBuf const &) = delete;
    PBuf & operator=(PBuf const &) = delete;
    PBuf(PBuf && o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    PBuf & operator=(PBuf && o) noexcept
    {
        if (this != &o)
        {
            std::free(p);
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~PBuf() { std::free(p); }
    void alloc(size_t count)
    {
        std::free(p);
        p = count ? static_cast<T *>(std::malloc(count * sizeof(T))) : nullptr;
        n = count;
    }
    void assign(size_t count, T init)
    {
        alloc(count);
        fill(0, count, init);
    }
    void fill(size_t b, size_t e, T init)
    {
        T * const q = p;
        parallel_for(static_cast<int64_t>(e - b), [=](int64_t lo, int64_t hi) { std::fill(q + b + lo, q + b + hi, init); });
    }
    T * data() { return p; }
    T const * data() const { return p; }
    T * begin() { return p; }
    T * end() { return p + n; }
    size_t size() const { return n; }
    bool empty() const { return 0 == n; }
    T & operator[](size_t k) { return p[k]; }
    T const & operator[](size_t k) const { return p[k]; }
};

template <typename T>
struct Tab
{
    PBuf<T> v;
    int64_t rows = 0, cols = 1, ngh = 0;
    bool one_d = false;
    // A caller may hold zero-copy views of `v` (scg_hmesh_table).  In the reference an ndarray view keeps the old
    // SimpleArray buffer alive across build_* calls (python/common.hpp:209-215); here a buffer that has been handed
    // out is RETIRED instead of freed when the table is re-made, so stale views stay readable (they show the old
    // generation, like the reference's) and writes through them cannot corrupt the heap.
    bool exported = false;
    // the table lives on the device (built by scg_hmesh_dev.cu) and has not been brought to the host yet: rows / cols /
    // ngh are set, `v` is empty until somebody asks for the table
    bool stale = false;
    std::vector<PBuf<T>> retired;
    void on_device(int64_t r, int64_t ng)
    {
        retire();
        rows = r;
        ngh = ng;
        stale = true;
    }
    void retire()
    {
        if (exported && !v.empty()) { retired.emplace_back(std::move(v)); }
        v = PBuf<T>();
        exported = false;
    }
    void make(int64_t r, int64_t c, T init, bool is1d = false)
    {
        retire();
        rows = r;
        cols = c;
        ngh = 0;
        one_d = is1d;
        v.assign(static_cast<size_t>(r * c), init);
    }
    // prepend `ng` ghost rows filled with `init` (StaticMesh_boundary.cpp:219-234)
    void grow_front(int64_t ng, T init)
    {
        PBuf<T> nv;
        nv.alloc(static_cast<size_t>((rows + ng) * cols));
        nv.fill(0, static_cast<size_t>(ng * cols), init);
        {
            T const * const src = v.data();
            T * const dst = nv.data() + ng * cols;
            parallel_for(static_cast<int64_t>(v.size()), [=](int64_t lo, int64_t hi) { std::copy(src + lo, src + hi, dst + lo); });
        }
        retire();
        v = std::move(nv);
        rows += ng;
        ngh = ng;
    }
    void drop_retired()
    {
        retired.clear();
        retired.shrink_to_fit();
    }
    T * body() { return v.data() + ngh * cols; }
    T const * body() const { return v.data() + ngh * cols; }
    T & operator()(int64_t i) { return body()[i * cols]; }
    T & operator()(int64_t i, int64_t j) { return body()[i * cols + j]; }
    T const & operator()(int64_t i) const { return body()[i * cols]; }
    T const & operator()(int64_t i, int64_t j) const { return body()[i * cols + j]; }
    int64_t nbody() const { return rows - ngh; }
};

struct FaceDef
{
    int nface;
    int fnnd[CLMFC];
    int nd[CLMFC][FCMND]; // 1-based positions in clnds
};

// StaticMesh_interior.cpp:61-330 (add_quadrilateral .. add_pyramid)
FaceDef const FACEDEF[9] = {
    {0, {0}, {{0}}},
    /* 1 point (StaticMesh_interior.cpp:27-39) */ {1, {1}, {{1}}},
    /* 2 line  (StaticMesh_interior.cpp:41-60) */ {2, {1, 1}, {{1}, {2}}},
    /* 3 quadrilateral */ {4, {2, 2, 2, 2}, {{1, 2}, {2, 3}, {3, 4}, {4, 1}}},
    /* 4 triangle      */ {3, {2, 2, 2}, {{1, 2}, {2, 3}, {3, 1}}},
    /* 5 hexahedron    */ {6, {4, 4, 4, 4, 4, 4}, {{1, 4, 3, 2}, {2, 3, 7, 6}, {5, 6, 7, 8}, {1, 5, 8, 4}, {1, 2, 6, 5}, {3, 4, 8, 7}}},
    /* 6 tetrahedron   */ {4, {3, 3, 3, 3}, {{1, 3, 2}, {1, 2, 4}, {1, 4, 3}, {2, 3, 4}}},
    /* 7 prism         */ {5, {3, 3, 4, 4, 4}, {{1, 2, 3}, {4, 6, 5}, {1, 4, 5, 2}, {1, 3, 6, 4}, {2, 5, 6, 3}}},
    /* 8 pyramid       */ {5, {3, 3, 3, 3, 4}, {{1, 5, 4}, {2, 5, 1}, {3, 5, 2}, {4, 5, 3}, {1, 4, 3, 2}}},
};

inline int face_type_of(int ndim, int nnd)
{
    if (1 == nnd) { return 1; /*POINT: the faces of a line cell*/ }
    return 2 == ndim ? 2 /*LINE*/ : (3 == nnd ? 4 /*TRIANGLE*/ : 3 /*QUADRILATERAL*/);
}
inline int cell_nface(int t) { return (t >= 0 && t <= 8) ? FACEDEF[t].nface : 0; }

// SCG_HMESH_TIMING=1: wall time of the builder's stages on stderr
struct StageTimer
{
    bool on;
    std::chrono::steady_clock::time_point t;
    StageTimer() : on(std::getenv("SCG_HMESH_TIMING") && std::getenv("SCG_HMESH_TIMING")[0] == '1'), t(std::chrono::steady_clock::now()) {}
    void lap(char const * what)
    {
        if (!on) { return; }
        auto const n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[hmesh] %-28s %8.3f s\n", what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};

struct Bc
{
    std::string name;
    Tab<int32_t> facn; // [n][3]
};

} // namespace

struct scg_hmesh
{
    int ndim = 0;
    int64_t nnode = 0, nface = 0, ncell = 0, nbound = 0, ngstnode = 0, ngstface = 0, ngstcell = 0;
    Tab<double> ndcrd, fccnd, fcnml, fcara, clcnd, clvol;
    Tab<int32_t> fctpn, cltpn, clgrp, fcnds, fccls, clnds, clfcs, ednds, bndfcs;
    std::vector<std::unique_ptr<Bc>> bcs;
    bool use_incenter = false;
    // device builder (scg_hmesh_dev.cu): -1 = automatic (a CUDA device is present and the mesh is large), 0 = host, 1 = device
    int dev_mode = -1;
    std::unique_ptr<scg::hdev::Tables> dev;
    bool device_wanted() const;
    int fetch(char const * name); // bring one table to the host if it only lives on the device
    int fetch_all();

    void build_faces_from_cells();
    void calc_metric();
    void build_edge();
    void face_centroid_normal(int64_t ifc);
    void cell_centroid(int64_t icl);
    void fill_ghost();
};

// --------------------------------------------------------------------------------------------------
// faces from cells, StaticMesh_interior.cpp:351-563
// --------------------------------------------------------------------------------------------------
void scg_hmesh::build_faces_from_cells()
{
    StageTimer tm;
    PBuf<int64_t> first;
    first.alloc(static_cast<size_t>(ncell) + 1);
    first[0] = 0;
    for (int64_t icl = 0; icl < ncell; ++icl) { first[icl + 1] = first[icl] + cell_nface(cltpn(icl)); }
    int64_t const mface = first[ncell];
    PBuf<int32_t> tnds, ttpn, key;
    tnds.alloc(static_cast<size_t>(mface) * (FCMND + 1));
    ttpn.alloc(static_cast<size_t>(mface));
    key.alloc(static_cast<size_t>(mface) * 4);
    clfcs.make(ncell, CLMFC + 1, -1);
    // populate every candidate face in cell order, with its sorted node tuple
    parallel_for(ncell, [&](int64_t b, int64_t e)
                 {
                     for (int64_t icl = b; icl < e; ++icl)
                     {
                         FaceDef const & fd = FACEDEF[(cltpn(icl) >= 0 && cltpn(icl) <= 8) ? cltpn(icl) : 0];
                         if (0 == fd.nface) { continue; }
                         clfcs(icl, 0) = fd.nface;
                         for (int f = 0; f < fd.nface; ++f)
                         {
                             int64_t const ifc = first[icl] + f;
                             clfcs(icl, f + 1) = static_cast<int32_t>(ifc);
                             ttpn[ifc] = face_type_of(ndim, fd.fnnd[f]);
                             tnds[ifc * 5] = fd.fnnd[f];
                             int32_t k[4] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX};
                             for (int q = 0; q < FCMND; ++q)
                             {
                                 int32_t const nd = (q < fd.fnnd[f]) ? clnds(icl, fd.nd[f][q]) : -1;
                                 tnds[ifc * 5 + 1 + q] = nd;
                                 if (q < fd.fnnd[f]) { k[q] = nd; }
                             }
                             std::sort(k, k + 4);
                             std::memcpy(&key[ifc * 4], k, sizeof(k));
                         }
                     }
                 });
    tm.lap("faces: populate + keys");
    // counting sort of the candidates by smallest node id (stable: original order inside a bucket)
    PBuf<int64_t> bstart;
    bstart.assign(static_cast<size_t>(nnode) + 2, 0);
    for (int64_t f = 0; f < mface; ++f) { bstart[key[f * 4] + 2] += 1; }
    for (int64_t n = 0; n < nnode; ++n) { bstart[n + 2] += bstart[n + 1]; }
    PBuf<int32_t> sorted, root, map2;
    sorted.alloc(static_cast<size_t>(mface));
    for (int64_t f = 0; f < mface; ++f) { sorted[bstart[key[f * 4] + 1]++] = static_cast<int32_t>(f); }
    tm.lap("faces: counting sort");
    // bstart[n] is now the start of bucket n, bstart[n+1] its end
    root.alloc(static_cast<size_t>(mface));
    parallel_for(mface, [&](int64_t b, int64_t e)
                 {
                     for (int64_t f = b; f < e; ++f) { root[f] = static_cast<int32_t>(f); }
                 });
    parallel_for(nnode, [&](int64_t nb, int64_t ne)
                 {
                     for (int64_t n = nb; n < ne; ++n)
                     {
                         int64_t const b = bstart[n], e = bstart[n + 1];
                         for (int64_t p = b; p < e; ++p)
                         {
                             int32_t const a = sorted[p];
                             if (root[a] != a) { continue; }
                             for (int64_t q = p + 1; q < e; ++q)
                             {
                                 int32_t const c = sorted[q];
                                 if (root[c] == c && ttpn[c] == ttpn[a] && 0 == std::memcmp(&key[int64_t(a) * 4], &key[int64_t(c) * 4], 16))
                                 {
                                     root[c] = a;
                                 }
                             }
                         }
                     }
                 });
    tm.lap("faces: match duplicates");
    // compact numbering of the roots in order of first appearance: a chunked prefix sum over the root flags, then the
    // duplicates take their root's number (a root always precedes its duplicates)
    map2.alloc(static_cast<size_t>(mface));
    {
        int64_t const nchunk = 256;
        int64_t const clen = (mface + nchunk - 1) / nchunk;
        std::vector<int64_t> cnt(static_cast<size_t>(nchunk) + 1, 0);
        parallel_for(nchunk, [&](int64_t cb, int64_t ce)
                     {
                         for (int64_t c = cb; c < ce; ++c)
                         {
                             int64_t n = 0;
                             for (int64_t f = c * clen; f < std::min(mface, (c + 1) * clen); ++f) { n += (root[f] == f) ? 1 : 0; }
                             cnt[c + 1] = n;
                         }
                     });
        for (int64_t c = 0; c < nchunk; ++c) { cnt[c + 1] += cnt[c]; }
        nface = cnt[nchunk];
        parallel_for(nchunk, [&](int64_t cb, int64_t ce)
                     {
                         for (int64_t c = cb; c < ce; ++c)
                         {
                             int64_t n = cnt[c];
                             for (int64_t f = c * clen; f < std::min(mface, (c + 1) * clen); ++f)
                             {
                                 if (root[f] == f) { map2[f] = static_cast<int32_t>(n++); }
                             }
                         }
                     });
        parallel_for(mface, [&](int64_t b, int64_t e)
                     {
                         for (int64_t f = b; f < e; ++f)
                         {
                             if (root[f] != f) { map2[f] = map2[root[f]]; }
                         }
                     });
    }
    fctpn.make(nface, 1, 0, true);
    fcnds.make(nface, FCMND + 1, -1);
    fccls.make(nface, FCREL, -1);
    parallel_for(mface, [&](int64_t b, int64_t e)
                 {
                     for (int64_t f = b; f < e; ++f)
                     {
                         if (root[f] != f) { continue; }
                         int64_t const j = map2[f];
                         fctpn(j) = ttpn[f];
                         for (int q = 0; q <= FCMND; ++q) { fcnds(j, q) = tnds[f * 5 + q]; }
                     }
                 });
    // clfcs takes the compact numbers; fccls[:,0] is the first cell that sees the face (the root candidate's: cells
    // are visited in ascending order and a root precedes its duplicates), fccls[:,1] the second one
    parallel_for(ncell, [&](int64_t b, int64_t e)
                 {
                     for (int64_t icl = b; icl < e; ++icl)
                     {
                         for (int ifl = 1; ifl <= clfcs(icl, 0); ++ifl)
                         {
                             int64_t const cand = clfcs(icl, ifl);
                             int32_t const jfc = map2[cand];
                             clfcs(icl, ifl) = jfc;
                             if (root[cand] == cand) { fccls(jfc, 0) = static_cast<int32_t>(icl); }
                             else
                             {
                                 // the lowest second cell wins, as in the sequential visit (more than two cells on a face
                                 // only happens in an invalid mesh)
                                 int32_t * const slot = &fccls(jfc, 1);
                                 int32_t cur = __atomic_load_n(slot, __ATOMIC_RELAXED);
                                 while ((-1 == cur || static_cast<int32_t>(icl) < cur)
                                        && !__atomic_compare_exchange_n(slot, &cur, static_cast<int32_t>(icl), true, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                                 {
                                 }
                             }
                         }
                     }
                 });
    fccnd.make(nface, ndim, 0.0);
    fcnml.make(nface, ndim, 0.0);
    fcara.make(nface, 1, 0.0, true);
    tm.lap("faces: number + tables");
}

// face centroid, unit normal and area of one face (body or ghost index):
// StaticMesh_interior.cpp:623-766, StaticMesh_boundary.cpp:411-519
void scg_hmesh::face_centroid_normal(int64_t ifc)
{
    if (2 == ndim)
    {
        int64_t const ind1 = fcnds(ifc, 1), ind2 = fcnds(ifc, 2);
        fccnd(ifc, 0) = ndcrd(ind1, 0);
        fccnd(ifc, 1) = ndcrd(ind1, 1);
        fccnd(ifc, 0) += ndcrd(ind2, 0);
        fccnd(ifc, 1) += ndcrd(ind2, 1);
        fccnd(ifc, 0) /= 2;
        fccnd(ifc, 1) /= 2;
        fcnml(ifc, 0) = ndcrd(ind2, 1) - ndcrd(ind1, 1);
        fcnml(ifc, 1) = ndcrd(ind1, 0) - ndcrd(ind2, 0);
        fcara(ifc) = std::sqrt(fcnml(ifc, 0) * fcnml(ifc, 0) + fcnml(ifc, 1) * fcnml(ifc, 1));
        fcnml(ifc, 0) /= fcara(ifc);
        fcnml(ifc, 1) /= fcara(ifc);
        return;
    }
    double cfd[FCMND + 2][3];
    cfd[0][0] = cfd[0][1] = cfd[0][2] = 0.0;
    int const nnd = fcnds(ifc, 0);
    for (int inf = 1; inf <= nnd; ++inf)
    {
        int64_t const ind = fcnds(ifc, inf);
        for (int d = 0; d < 3; ++d)
        {
            cfd[inf][d] = ndcrd(ind, d);
            cfd[0][d] += ndcrd(ind, d);
        }
    }
    for (int d = 0; d < 3; ++d)
    {
        cfd[nnd + 1][d] = cfd[1][d];
        cfd[0][d] /= nnd;
    }
    double voc = 0.0;
    double fc[3] = {0.0, 0.0, 0.0};
    for (int inf = 1; inf <= nnd; ++inf)
    {
        double crd[3];
        for (int d = 0; d < 3; ++d) { crd[d] = (cfd[0][d] + cfd[inf][d] + cfd[inf + 1][d]) / 3; }
        double const du0 = cfd[inf][0] - cfd[0][0];
        double const du1 = cfd[inf][1] - cfd[0][1];
        double const du2 = cfd[inf][2] - cfd[0][2];
        double const dv0 = cfd[inf + 1][0] - cfd[0][0];
        double const dv1 = cfd[inf + 1][1] - cfd[0][1];
        double const dv2 = cfd[inf + 1][2] - cfd[0][2];
        double const dw0 = du1 * dv2 - du2 * dv1;
        double const dw1 = du2 * dv0 - du0 * dv2;
        double const dw2 = du0 * dv1 - du1 * dv0;
        double const vob = std::sqrt(dw0 * dw0 + dw1 * dw1 + dw2 * dw2);
        fc[0] += crd[0] * vob;
        fc[1] += crd[1] * vob;
        fc[2] += crd[2] * vob;
        voc += vob;
    }
    for (int d = 0; d < 3; ++d)
    {
        fc[d] /= voc;
        fccnd(ifc, d) = fc[d];
    }
    double rv[FCMND][3];
    for (int inf = 0; inf < nnd; ++inf)
    {
        int64_t const ind = fcnds(ifc, inf + 1);
        for (int d = 0; d < 3; ++d) { rv[inf][d] = ndcrd(ind, d) - fc[d]; }
    }
    double nm[3];
    nm[0] = rv[nnd - 1][1] * rv[0][2] - rv[nnd - 1][2] * rv[0][1];
    nm[1] = rv[nnd - 1][2] * rv[0][0] - rv[nnd - 1][0] * rv[0][2];
    nm[2] = rv[nnd - 1][0] * rv[0][1] - rv[nnd - 1][1] * rv[0][0];
    for (int q = 1; q < nnd; ++q)
    {
        nm[0] += rv[q - 1][1] * rv[q][2] - rv[q - 1][2] * rv[q][1];
        nm[1] += rv[q - 1][2] * rv[q][0] - rv[q - 1][0] * rv[q][2];
        nm[2] += rv[q - 1][0] * rv[q][1] - rv[q - 1][1] * rv[q][0];
    }
    double ara = std::sqrt(nm[0] * nm[0] + nm[1] * nm[1] + nm[2] * nm[2]);
    for (int d = 0; d < 3; ++d) { fcnml(ifc, d) = nm[d] / ara; }
    fcara(ifc) = ara / 2.0;
}

// volume-weighted cell centroid (body or ghost index): StaticMesh_interior.cpp:800-835, 889-927,
// StaticMesh_boundary.cpp:522-600
void scg_hmesh::cell_centroid(int64_t icl)
{
    double crd[3] = {0.0, 0.0, 0.0};
    int const nnd = clnds(icl, 0);
    for (int inc = 1; inc <= nnd; ++inc)
    {
        int64_t const ind = clnds(icl, inc);
        for (int d = 0; d < ndim; ++d) { crd[d] += ndcrd(ind, d); }
    }
    for (int d = 0; d < ndim; ++d) { crd[d] /= nnd; }
    double voc = 0.0;
    double cc[3] = {0.0, 0.0, 0.0};
    int const nfc = clfcs(icl, 0);
    double const part = ndim + 1;
    for (int ifl = 1; ifl <= nfc; ++ifl)
    {
        int64_t const ifc = clfcs(icl, ifl);
        double du[3];
        double dot = 0.0;
        for (int d = 0; d < ndim; ++d) { du[d] = crd[d] - fccnd(ifc, d); }
        if (2 == ndim) { dot = du[0] * fcnml(ifc, 0) + du[1] * fcnml(ifc, 1); }
        else { dot = du[0] * fcnml(ifc, 0) + du[1] * fcnml(ifc, 1) + du[2] * fcnml(ifc, 2); }
        // body cells: fabs(dot) * fcara; 3D ghost cells: fabs(dot * fcara) -- the same value in IEEE
        // arithmetic because fcara >= 0 (|a*b| == |a|*|b| exactly, rounding is sign-symmetric)
        double const vob = std::fabs(dot) * fcara(ifc);
        voc += vob;
        for (int d = 0; d < ndim; ++d)
        {
            double const dv = fccnd(ifc, d) + du[d] / part;
            cc[d] += dv * vob;
        }
    }
    for (int d = 0; d < ndim; ++d) { clcnd(icl, d) = cc[d] / voc; }
}

void scg_hmesh::calc_metric()
{
    StageTimer tm;
    // one-dimensional meshes have no face / cell metric (the reference's loops are `if (2 == ndim) .. else if
    // (3 == ndim)`, StaticMesh_interior.cpp:626-927); only the volume loop below runs, on zeros
    if (ndim >= 2)
    {
        parallel_for(nface, [&](int64_t b, int64_t e)
                     {
                         for (int64_t f = b; f < e; ++f) { face_centroid_normal(f); }
                     });
        // incenter variants (StaticMesh_interior.cpp:771-797, 841-886) are unreachable: m_use_incenter is
        // never set (StaticMesh.hpp:409)
        parallel_for(ncell, [&](int64_t b, int64_t e)
                     {
                         for (int64_t c = b; c < e; ++c) { cell_centroid(c); }
                     });
    }
    tm.lap("metric: face + cell centroids");
    // Cell volumes and face orientation (StaticMesh_interior.cpp:930-976).  The reference walks the cells in order and
    // lets every cell re-orient a face it sees; per face that is: the owner (first cell) visits, then the neighbour.
    // Re-orienting negates the normal, which negates the face's volume term EXACTLY (every product and the ordered sum
    // change sign, rounding is symmetric), and the term enters the volume by its magnitude -- so the volumes do not
    // depend on the visiting order and the final orientation of a face depends on its own two visits only.
    parallel_for(ncell, [&](int64_t b, int64_t e)
                 {
                     for (int64_t icl = b; icl < e; ++icl)
                     {
                         double acc = 0.0;
                         int const nfc = clfcs(icl, 0);
                         for (int it = 1; it <= nfc; ++it)
                         {
                             int64_t const ifc = clfcs(icl, it);
                             double vol = 0.0;
                             for (int d = 0; d < ndim; ++d) { vol += (fccnd(ifc, d) - clcnd(icl, d)) * fcnml(ifc, d); }
                             vol *= fcara(ifc);
                             if (vol < 0.0) { vol = -vol; }
                             acc += vol;
                         }
                         clvol(icl) = acc / ndim;
                     }
                 });
    parallel_for(nface, [&](int64_t b, int64_t e)
                 {
                     for (int64_t ifc = b; ifc < e; ++ifc)
                     {
                         bool flipped = false;
                         for (int side = 0; side < 2; ++side)
                         {
                             int64_t const icl = fccls(ifc, side);
                             if (icl < 0) { continue; }
                             double vol = 0.0;
                             for (int d = 0; d < ndim; ++d) { vol += (fccnd(ifc, d) - clcnd(icl, d)) * fcnml(ifc, d); }
                             vol *= fcara(ifc);
                             if (flipped) { vol = -vol; }          // what the visit computes with the current normal
                             bool const mine = (0 == side);
                             bool const flip = (vol < 0.0) ? mine : !mine;
                             if (flip) { flipped = !flipped; }
                         }
                         if (flipped)
                         {
                             int const nnd = fcnds(ifc, 0);
                             int32_t tmp[FCMND];
                             for (int jt = 0; jt < nnd; ++jt) { tmp[jt] = fcnds(ifc, nnd - jt); }
                             for (int jt = 0; jt < nnd; ++jt) { fcnds(ifc, jt + 1) = tmp[jt]; }
                             for (int d = 0; d < ndim; ++d) { fcnml(ifc, d) = -fcnml(ifc, d); }
                         }
                     }
                 });
    tm.lap("metric: volumes + orientation");
}

// StaticMesh_interior.cpp:566-609: unique edges in order of first appearance over the faces
void scg_hmesh::build_edge()
{
    std::unordered_set<uint64_t> known;
    std::vector<int32_t> edges;
    for (int64_t ifc = 0; ifc < nface; ++ifc)
    {
        int const nnd = fcnds(ifc, 0);
        for (int inf = 1; inf <= nnd; ++inf)
        {
            int const nxt = (nnd == inf) ? 1 : inf + 1;
            int32_t nd0 = fcnds(ifc, inf), nd1 = fcnds(ifc, nxt);
            if (nd0 > nd1) { std::swap(nd0, nd1); }
            uint64_t const k = (static_cast<uint64_t>(static_cast<uint32_t>(nd0)) << 32) | static_cast<uint32_t>(nd1);
            if (known.insert(k).second)
            {
                edges.push_back(nd0);
                edges.push_back(nd1);
            }
        }
    }
    ednds.make(static_cast<int64_t>(edges.size() / 2), 2, 0);
    std::copy(edges.begin(), edges.end(), ednds.v.begin());
}

// StaticMesh_boundary.cpp:269-627
void scg_hmesh::fill_ghost()
{
    int32_t const none = static_cast<int32_t>(nnode);
    std::vector<int32_t> gstndmap(static_cast<size_t>(nnode), none);
    int32_t ignd = -1, igfc = -1;
    for (int32_t igcl = -1; igcl >= -static_cast<int32_t>(ngstcell); --igcl)
    {
        int32_t const ibfc = bndfcs(-igcl - 1, 0);
        int32_t const icl = fccls(ibfc, 0);
        cltpn(igcl) = cltpn(icl);
        clgrp(igcl) = clgrp(icl);
        for (int inl = 0; inl <= CLMND; ++inl) { clnds(igcl, inl) = clnds(icl, inl); }
        for (int inl = 1; inl <= clnds(icl, 0); ++inl)
        {
            int32_t const ind = clnds(icl, inl);
            bool found = false;
            for (int inf = 1; inf <= fcnds(ibfc, 0); ++inf)
            {
                if (ind == fcnds(ibfc, inf))
                {
                    found = true;
                    break;
                }
            }
            if (!found)
            {
                gstndmap[ind] = ignd;
                clnds(igcl, inl) = ignd;
                double dist = 0.0;
                for (int d = 0; d < ndim; ++d) { dist += (fccnd(ibfc, d) - ndcrd(ind, d)) * fcnml(ibfc, d); }
                for (int d = 0; d < ndim; ++d) { ndcrd(ignd, d) = ndcrd(ind, d) + 2 * dist * fcnml(ibfc, d); }
                ignd -= 1;
            }
        }
        fccls(ibfc, 1) = igcl;
        for (int ifl = 0; ifl <= CLMFC; ++ifl) { clfcs(igcl, ifl) = clfcs(icl, ifl); }
        for (int ifl = 1; ifl <= clfcs(icl, 0); ++ifl)
        {
            int32_t const ifc = clfcs(icl, ifl);
            if (ifc == ibfc) { continue; }
            fctpn(igfc) = fctpn(ifc);
            fccls(igfc, 0) = igcl;
            clfcs(igcl, ifl) = igfc;
            for (int inf = 0; inf <= FCMND; ++inf) { fcnds(igfc, inf) = fcnds(ifc, inf); }
            for (int inf = 1; inf <= fcnds(igfc, 0); ++inf)
            {
                int32_t const ind = fcnds(igfc, inf);
                if (gstndmap[ind] != none) { fcnds(igfc, inf) = gstndmap[ind]; }
            }
            igfc -= 1;
        }
        for (int inl = 1; inl <= clnds(icl, 0); ++inl) { gstndmap[clnds(icl, inl)] = none; }
    }
    for (int64_t ifc = -1; ifc >= -ngstface; --ifc) { face_centroid_normal(ifc); }
    for (int64_t icl = -1; icl >= -ngstcell; --icl) { cell_centroid(icl); }
    for (int64_t icl = -1; icl >= -ngstcell; --icl)
    {
        clvol(icl) = 0.0;
        for (int it = 1; it <= clfcs(icl, 0); ++it)
        {
            int64_t const ifc = clfcs(icl, it);
            double vol = 0.0;
            for (int d = 0; d < ndim; ++d) { vol += (fccnd(ifc, d) - clcnd(icl, d)) * fcnml(ifc, d); }
            vol *= fcara(ifc);
            if (vol < 0.0)
            {
                if (fccls(ifc, 0) == icl)
                {
                    for (int d = 0; d < ndim; ++d) { fcnml(ifc, d) = -fcnml(ifc, d); }
                }
                vol = -vol;
            }
            clvol(icl) += vol;
        }
        clvol(icl) /= ndim;
    }
}

bool scg_hmesh::device_wanted() const
{
    if (ndim < 2) { return false; }
    int mode = dev_mode;
    if (mode < 0)
    {
        char const * const env = std::getenv("SCG_HMESH_DEVICE");
        if (env && ('0' == env[0] || '1' == env[0])) { mode = env[0] - '0'; }
    }
    if (0 == mode) { return false; }
    int ndev = 0;
    if (cudaSuccess != cudaGetDeviceCount(&ndev) || ndev < 1)
    {
        cudaGetLastError();
        return false;
    }
    return 1 == mode || ncell >= 200000;
}

namespace
{
template <typename T>
int fetch_tab(Tab<T> & t, scg::DevBuf<T> const & d)
{
    if (!t.stale) { return 0; }
    size_t const n = static_cast<size_t>(t.rows * t.cols);
    t.v.alloc(n);
    if (n > 0)
    {
        if (static_cast<int64_t>(n) != d.n) { return scg::fail("StaticMesh: device table size mismatch (%lld != %lld)", static_cast<long long>(n), static_cast<long long>(d.n)); }
        cudaError_t const e = cudaMemcpy(t.v.data(), d.p, n * sizeof(T), cudaMemcpyDeviceToHost);
        if (cudaSuccess != e) { return scg::fail("StaticMesh: download of a device-built table failed: %s", cudaGetErrorString(e)); }
    }
    t.stale = false;
    return 0;
}
} // namespace

int scg_hmesh::fetch(char const * name)
{
    if (!dev) { return 0; }
    std::string const n(name);
#define HM_FETCH(N)                        \
    if (n == #N) { return fetch_tab(N, dev->N); }
    HM_FETCH(ndcrd) HM_FETCH(fccnd) HM_FETCH(fcnml) HM_FETCH(fcara) HM_FETCH(clcnd) HM_FETCH(clvol)
    HM_FETCH(fctpn) HM_FETCH(cltpn) HM_FETCH(clgrp) HM_FETCH(fcnds) HM_FETCH(fccls) HM_FETCH(clnds) HM_FETCH(clfcs)
#undef HM_FETCH
    return 0;
}

int scg_hmesh::fetch_all()
{
    for (char const * n : {"ndcrd", "fccnd", "fcnml", "fcara", "clcnd", "clvol", "fctpn", "cltpn", "clgrp", "fcnds", "fccls", "clnds", "clfcs"})
    {
        if (int rc = fetch(n)) { return rc; }
    }
    return 0;
}

scg::hdev::Tables * scg::hdev::tables_of(scg_hmesh * m) { return m ? m->dev.get() : nullptr; }

using scg::fail;

#define HM_REQUIRE(cond, ...)          \
    do                                 \
    {                                  \
        if (!(cond)) { return fail(__VA_ARGS__); } \
    } while (0)

extern "C" {

int scg_hmesh_create(scg_hmesh ** out, int ndim, int64_t nnode, int64_t nface, int64_t ncell)
{
    HM_REQUIRE(out != nullptr, "scg_hmesh_create: null output");
    HM_REQUIRE(ndim >= 1 && ndim <= 3, "StaticMesh: ndim must be 1, 2 or 3");
    HM_REQUIRE(nnode >= 0 && nface >= 0 && ncell >= 0, "StaticMesh: negative size");
    auto * m = new scg_hmesh;
    m->ndim = ndim;
    m->nnode = nnode;
    m->nface = nface;
    m->ncell = ncell;
    // StaticMesh.hpp:300-320
    m->ndcrd.make(nnode, ndim, 0.0);
    m->fccnd.make(nface, ndim, 0.0);
    m->fcnml.make(nface, ndim, 0.0);
    m->fcara.make(nface, 1, 0.0, true);
    m->clcnd.make(ncell, ndim, 0.0);
    m->clvol.make(ncell, 1, 0.0, true);
    m->fctpn.make(nface, 1, 0, true);
    m->cltpn.make(ncell, 1, 0, true);
    m->clgrp.make(ncell, 1, 0, true);
    m->fcnds.make(nface, FCMND + 1, 0);
    m->fccls.make(nface, FCREL, 0);
    m->clnds.make(ncell, CLMND + 1, 0);
    m->clfcs.make(ncell, CLMFC + 1, 0);
    m->ednds.make(0, 2, 0);
    m->bndfcs.make(0, BFREL, 0);
    *out = m;
    return 0;
}

int scg_hmesh_destroy(scg_hmesh * m)
{
    delete m;
    return 0;
}

int scg_hmesh_count(scg_hmesh * m, int which, int64_t * out)
{
    HM_REQUIRE(m && out, "scg_hmesh_count: null argument");
    switch (which)
    {
    case 0: *out = m->ndim; break;
    case 1: *out = m->nnode; break;
    case 2: *out = m->nface; break;
    case 3: *out = m->ncell; break;
    case 4: *out = m->nbound; break;
    case 5: *out = m->ngstnode; break;
    case 6: *out = m->ngstface; break;
    case 7: *out = m->ngstcell; break;
    case 8: *out = m->ednds.rows; break;
    case 9: *out = static_cast<int64_t>(m->bcs.size()); break;
    default: return fail("scg_hmesh_count: unknown counter %d", which);
    }
    return 0;
}

int scg_hmesh_table(scg_hmesh * m, const char * name, void ** ptr, int64_t * rows, int64_t * cols, int64_t * nghost, int * is_real,
                    int * one_d)
{
    HM_REQUIRE(m && name && ptr && rows && cols && nghost && is_real && one_d, "scg_hmesh_table: null argument");
    if (int rc = m->fetch(name)) { return rc; }
    std::string const n(name);
#define HM_TAB(N, REAL)                                \
    if (n == #N)                                       \
    {                                                  \
        *ptr = m->N.v.data();                          \
        m->N.exported = true;                          \
        *rows = m->N.rows;                             \
        *cols = m->N.cols;                             \
        *nghost = m->N.ngh;                            \
        *is_real = REAL;                               \
        *one_d = m->N.one_d ? 1 : 0;                   \
        return 0;                                      \
    }
    HM_TAB(ndcrd, 1) HM_TAB(fccnd, 1) HM_TAB(fcnml, 1) HM_TAB(fcara, 1) HM_TAB(clcnd, 1) HM_TAB(clvol, 1)
    HM_TAB(fctpn, 0) HM_TAB(cltpn, 0) HM_TAB(clgrp, 0) HM_TAB(fcnds, 0) HM_TAB(fccls, 0) HM_TAB(clnds, 0)
    HM_TAB(clfcs, 0) HM_TAB(ednds, 0) HM_TAB(bndfcs, 0)
#undef HM_TAB
    return fail("StaticMesh: unknown table %s", name);
}

int scg_hmesh_set_device(scg_hmesh * m, int mode)
{
    HM_REQUIRE(m, "scg_hmesh_set_device: null mesh");
    HM_REQUIRE(mode >= -1 && mode <= 1, "scg_hmesh_set_device: mode must be -1 (automatic), 0 (host) or 1 (device)");
    m->dev_mode = mode;
    return 0;
}

int scg_hmesh_on_device(scg_hmesh * m, int * out)
{
    HM_REQUIRE(m && out, "scg_hmesh_on_device: null argument");
    // a table that was handed out as a (writable) view may have been changed on the host: from then on the host copies
    // are the mesh and the device copies are no longer offered
    bool const handed_out = m->ndcrd.exported || m->cltpn.exported || m->clgrp.exported || m->clnds.exported || m->fctpn.exported
                            || m->fcnds.exported || m->fccls.exported || m->clfcs.exported || m->fccnd.exported || m->fcnml.exported
                            || m->fcara.exported || m->clcnd.exported || m->clvol.exported;
    *out = (m->dev && m->dev->interior && !(m->dev->ghost && handed_out)) ? (m->dev->ghost ? 2 : 1) : 0;
    return 0;
}

// rows idx[0..n) (entity indices: >= 0 body, < 0 ghost) of one table -> out[n][cols], without bringing a device-built
// table to the host
int scg_hmesh_gather(scg_hmesh * m, const char * name, const int32_t * idx, int64_t n, void * out)
{
    HM_REQUIRE(m && name && out && (idx || 0 == n), "scg_hmesh_gather: null argument");
    std::string const nm(name);
#define HM_GATHER(N, T)                                                                                          \
    if (nm == #N)                                                                                                \
    {                                                                                                            \
        for (int64_t q = 0; q < n; ++q)                                                                          \
        {                                                                                                        \
            HM_REQUIRE(idx[q] >= -m->N.ngh && idx[q] < m->N.nbody(), "StaticMesh: index %d of table %s is out of range", idx[q], name); \
        }                                                                                                        \
        if (m->N.stale && m->dev) { return scg::hdev::gather_rows(m->dev->N.p, m->N.ngh, static_cast<int>(m->N.cols), idx, n, static_cast<T *>(out)); } \
        T * const o = static_cast<T *>(out);                                                                     \
        for (int64_t q = 0; q < n; ++q)                                                                          \
        {                                                                                                        \
            for (int64_t c = 0; c < m->N.cols; ++c) { o[q * m->N.cols + c] = m->N(idx[q], c); }                   \
        }                                                                                                        \
        return 0;                                                                                                \
    }
    HM_GATHER(ndcrd, double) HM_GATHER(fccnd, double) HM_GATHER(fcnml, double) HM_GATHER(fcara, double) HM_GATHER(clcnd, double)
    HM_GATHER(clvol, double) HM_GATHER(fctpn, int32_t) HM_GATHER(cltpn, int32_t) HM_GATHER(clgrp, int32_t) HM_GATHER(fcnds, int32_t)
    HM_GATHER(fccls, int32_t) HM_GATHER(clnds, int32_t) HM_GATHER(clfcs, int32_t)
#undef HM_GATHER
    return fail("StaticMesh: unknown table %s", name);
}

int scg_hmesh_build_interior(scg_hmesh * m, int do_metric, int do_edge)
{
    HM_REQUIRE(m, "scg_hmesh_build_interior: null mesh");
    if (m->device_wanted())
    {
        if (int rc = m->fetch_all()) { return rc; } // a rebuild starts from the host copies of nodes and cells
        m->dev.reset(new scg::hdev::Tables);
        StageTimer tm;
        int const rc = scg::hdev::build_interior(*m->dev, m->ndim, m->nnode, m->ncell, m->ndcrd.body(), m->cltpn.body(), m->clgrp.body(),
                                                 m->clnds.body(), 0 != do_metric);
        if (0 == rc)
        {
            tm.lap("device: interior tables");
            m->nface = m->dev->nface;
            m->fctpn.on_device(m->nface, 0);
            m->fcnds.on_device(m->nface, 0);
            m->fccls.on_device(m->nface, 0);
            m->clfcs.on_device(m->ncell, 0);
            m->fccnd.on_device(m->nface, 0);
            m->fcnml.on_device(m->nface, 0);
            m->fcara.on_device(m->nface, 0);
            m->clcnd.on_device(m->ncell, 0);
            m->clvol.on_device(m->ncell, 0);
            if (do_edge)
            {
                if (int rc2 = m->fetch("fcnds")) { return rc2; }
                m->build_edge();
            }
            return 0;
        }
        // An out-of-range node: the host loop below gives the reference's message.  Anything else (no memory left on the
        // device, a mesh too large for the int32 candidate numbering): the host builder makes the same tables.
        m->dev.reset();
        cudaGetLastError();
        if (std::getenv("SCG_HMESH_TIMING"))
        {
            std::fprintf(stderr, "[hmesh] device builder gave up (%s); building on the host\n", scg_last_error());
        }
    }
    m->dev.reset();
    for (int64_t icl = 0; icl < m->ncell; ++icl)
    {
        int const t = m->cltpn(icl);
        int const nnd = (1 == t) ? 1 : (2 == t) ? 2 : (3 == t) ? 4 : (4 == t) ? 3 : (5 == t) ? 8 : (6 == t) ? 4 : (7 == t) ? 6 : (8 == t) ? 5 : 0;
        for (int q = 1; q <= nnd; ++q)
        {
            int32_t const ind = m->clnds(icl, q);
            HM_REQUIRE(ind >= 0 && ind < m->nnode, "StaticMesh::build_interior: cell %lld node %d out of range [0, %lld)",
                       static_cast<long long>(icl), ind, static_cast<long long>(m->nnode));
        }
    }
    m->build_faces_from_cells();
    if (do_metric) { m->calc_metric(); }
    if (do_edge) { m->build_edge(); }
    return 0;
}

int scg_hmesh_build_edge(scg_hmesh * m)
{
    HM_REQUIRE(m, "scg_hmesh_build_edge: null mesh");
    if (int rc = m->fetch("fcnds")) { return rc; }
    m->build_edge();
    return 0;
}

// StaticMesh::add_bc, StaticMesh_boundary.cpp:14-78 (same checks, same messages)
int scg_hmesh_add_bc(scg_hmesh * m, const char * name, const int32_t * faces, int64_t n)
{
    HM_REQUIRE(m && name, "scg_hmesh_add_bc: null argument");
    HM_REQUIRE(0 == m->nbound, "StaticMesh::add_bc: boundary is already built");
    HM_REQUIRE(name[0] != '\0', "StaticMesh::add_bc: empty name");
    if (int rc = m->fetch("fccls")) { return rc; }
    for (auto const & b : m->bcs)
    {
        HM_REQUIRE(b->name != name, "StaticMesh::add_bc: duplicated name \"%s\"", name);
    }
    std::vector<char> claimed(static_cast<size_t>(m->nface), 0);
    for (auto const & b : m->bcs)
    {
        for (int64_t q = 0; q < b->facn.rows; ++q) { claimed[b->facn(q, 0)] = 1; }
    }
    for (int64_t q = 0; q < n; ++q)
    {
        int32_t const ifc = faces[q];
        HM_REQUIRE(ifc >= 0 && ifc < m->nface, "StaticMesh::add_bc: face %d out of range [0, %lld)", ifc,
                   static_cast<long long>(m->nface));
        HM_REQUIRE(m->fccls(ifc, 1) < 0, "StaticMesh::add_bc: face %d is not a boundary face", ifc);
        HM_REQUIRE(!claimed[ifc], "StaticMesh::add_bc: face %d is already claimed", ifc);
        claimed[ifc] = 1;
    }
    auto bc = std::make_unique<Bc>();
    bc->name = name;
    bc->facn.make(n, BFREL, -1);
    for (int64_t q = 0; q < n; ++q) { bc->facn(q, 0) = faces[q]; }
    m->bcs.push_back(std::move(bc));
    return 0;
}

int scg_hmesh_bc(scg_hmesh * m, int64_t ibc, const char ** name, int32_t ** facn, int64_t * nbound)
{
    HM_REQUIRE(m && name && facn && nbound, "scg_hmesh_bc: null argument");
    HM_REQUIRE(ibc >= 0 && ibc < static_cast<int64_t>(m->bcs.size()), "StaticMesh::bc: index %lld out of range",
               static_cast<long long>(ibc));
    Bc & b = *m->bcs[static_cast<size_t>(ibc)];
    *name = b.name.c_str();
    *facn = b.facn.v.data();
    *nbound = b.facn.rows;
    return 0;
}

// StaticMeshBc::set_name (StaticMesh.hpp:243; no validation, like the reference)
int scg_hmesh_bc_rename(scg_hmesh * m, int64_t ibc, const char * name)
{
    HM_REQUIRE(m && name, "scg_hmesh_bc_rename: null argument");
    HM_REQUIRE(ibc >= 0 && ibc < static_cast<int64_t>(m->bcs.size()), "StaticMesh::bc: index %lld out of range",
               static_cast<long long>(ibc));
    m->bcs[static_cast<size_t>(ibc)]->name = name;
    return 0;
}

// Free the table generations that were handed out as views and have since been re-made.  Only for callers that
// know no view of an earlier generation is alive (the bulk path of make_mesh()).
int scg_hmesh_drop_retired(scg_hmesh * m)
{
    HM_REQUIRE(m, "scg_hmesh_drop_retired: null mesh");
    m->ndcrd.drop_retired(); m->fccnd.drop_retired(); m->fcnml.drop_retired(); m->fcara.drop_retired();
    m->clcnd.drop_retired(); m->clvol.drop_retired(); m->fctpn.drop_retired(); m->cltpn.drop_retired();
    m->clgrp.drop_retired(); m->fcnds.drop_retired(); m->fccls.drop_retired(); m->clnds.drop_retired();
    m->clfcs.drop_retired(); m->ednds.drop_retired(); m->bndfcs.drop_retired();
    return 0;
}

// StaticMesh::build_boundary, StaticMesh_boundary.cpp:103-182
int scg_hmesh_build_boundary(scg_hmesh * m)
{
    HM_REQUIRE(m, "scg_hmesh_build_boundary: null mesh");
    HM_REQUIRE(0 == m->nbound, "StaticMesh::build_boundary: boundary is already built");
    // all boundary faces (no second cell) in ascending id: from the device tables when the interior was built there
    std::vector<int32_t> allfacn;
    if (m->dev && m->dev->interior && !m->fccls.exported)
    {
        if (int rc = scg::hdev::boundary_faces(*m->dev, allfacn)) { return rc; }
    }
    else
    {
        if (int rc = m->fetch("fccls")) { return rc; }
        for (int64_t f = 0; f < m->nface; ++f)
        {
            if (m->fccls(f, 1) < 0) { allfacn.push_back(static_cast<int32_t>(f)); }
        }
    }
    int64_t const nb = static_cast<int64_t>(allfacn.size());
    m->nbound = nb;
    m->bndfcs.make(nb, BFREL, -1);
    auto where = [&](int32_t f) -> int64_t
    {
        auto const it = std::lower_bound(allfacn.begin(), allfacn.end(), f);
        return (it != allfacn.end() && *it == f) ? (it - allfacn.begin()) : -1;
    };
    std::vector<char> specified(static_cast<size_t>(nb), 0);
    int64_t ibfc = 0, nleft = nb;
    for (size_t ibnd = 0; ibnd < m->bcs.size(); ++ibnd)
    {
        auto & facn = m->bcs[ibnd]->facn;
        for (int64_t q = 0; q < facn.rows; ++q)
        {
            m->bndfcs(ibfc, 0) = facn(q, 0);
            m->bndfcs(ibfc, 1) = static_cast<int32_t>(ibnd);
            facn(q, 1) = static_cast<int32_t>(ibfc);
            int64_t const pos = where(facn(q, 0));
            if (pos >= 0)
            {
                specified[pos] = 1;
                --nleft;
            }
            ++ibfc;
        }
    }
    if (nleft != 0)
    {
        auto bc = std::make_unique<Bc>();
        bc->name = "<NONAME>";
        bc->facn.make(nleft, BFREL, -1);
        int64_t q = 0;
        int32_t const ibnd = static_cast<int32_t>(m->bcs.size());
        for (int64_t s = 0; s < nb; ++s)
        {
            if (!specified[s])
            {
                m->bndfcs(ibfc, 0) = allfacn[s];
                m->bndfcs(ibfc, 1) = ibnd;
                bc->facn(q, 0) = allfacn[s];
                bc->facn(q, 1) = static_cast<int32_t>(ibfc);
                ++ibfc;
                ++q;
            }
        }
        m->bcs.push_back(std::move(bc));
    }
    return 0;
}

// StaticMesh::build_ghost, StaticMesh_boundary.cpp:185-267
int scg_hmesh_build_ghost(scg_hmesh * m)
{
    HM_REQUIRE(m, "scg_hmesh_build_ghost: null mesh");
    HM_REQUIRE(0 == m->ngstcell, "StaticMesh::build_ghost: ghost entities are already built");
    // a view of a device-built table that was handed out may have been written through: then the host copies rule
    bool const handed_out = m->fctpn.exported || m->fcnds.exported || m->fccls.exported || m->clfcs.exported || m->fccnd.exported
                            || m->fcnml.exported || m->fcara.exported || m->clcnd.exported || m->clvol.exported;
    if (m->dev && m->dev->interior && !m->dev->ghost && !handed_out)
    {
        // the device copies of the interior tables are current: ghost entities are built there too
        StageTimer tmd;
        std::vector<int32_t> bnd(static_cast<size_t>(m->nbound));
        for (int64_t ib = 0; ib < m->nbound; ++ib) { bnd[static_cast<size_t>(ib)] = m->bndfcs(ib, 0); }
        if (int rc = scg::hdev::build_ghost(*m->dev, bnd.data(), m->nbound)) { return rc; }
        tmd.lap("device: ghost entities");
        scg::hdev::Tables const & t = *m->dev;
        m->ngstnode = t.ngstnode;
        m->ngstface = t.ngstface;
        m->ngstcell = t.ngstcell;
        m->ndcrd.on_device(t.nnode + t.ngstnode, t.ngstnode);
        m->fccnd.on_device(t.nface + t.ngstface, t.ngstface);
        m->fcnml.on_device(t.nface + t.ngstface, t.ngstface);
        m->fcara.on_device(t.nface + t.ngstface, t.ngstface);
        m->clcnd.on_device(t.ncell + t.ngstcell, t.ngstcell);
        m->clvol.on_device(t.ncell + t.ngstcell, t.ngstcell);
        m->fctpn.on_device(t.nface + t.ngstface, t.ngstface);
        m->cltpn.on_device(t.ncell + t.ngstcell, t.ngstcell);
        m->clgrp.on_device(t.ncell + t.ngstcell, t.ngstcell);
        m->fcnds.on_device(t.nface + t.ngstface, t.ngstface);
        m->fccls.on_device(t.nface + t.ngstface, t.ngstface);
        m->clnds.on_device(t.ncell + t.ngstcell, t.ngstcell);
        m->clfcs.on_device(t.ncell + t.ngstcell, t.ngstcell);
        return 0;
    }
    // host path: every table on the host; device copies (if any) are out of date from here on
    if (int rc = m->fetch_all()) { return rc; }
    m->dev.reset();
    int64_t gn = 0, gf = 0;
    for (int64_t ib = 0; ib < m->nbound; ++ib)
    {
        int32_t const ifc = m->bndfcs(ib, 0);
        int32_t const icl = m->fccls(ifc, 0);
        gf += cell_nface(m->cltpn(icl)) - 1;
        gn += m->clnds(icl, 0) - m->fcnds(ifc, 0);
    }
    m->ngstnode = gn;
    m->ngstface = gf;
    m->ngstcell = m->nbound;
    auto grow_d = [](Tab<double> & t, int64_t ng, double init) { t.grow_front(ng, init); };
    auto grow_i = [](Tab<int32_t> & t, int64_t ng, int32_t init) { t.grow_front(ng, init); };
    StageTimer tm;
    grow_d(m->ndcrd, gn, 0.0);
    grow_d(m->fccnd, gf, 0.0);
    grow_d(m->fcnml, gf, 0.0);
    grow_d(m->fcara, gf, 0.0);
    grow_d(m->clcnd, m->ngstcell, 0.0);
    grow_d(m->clvol, m->ngstcell, 0.0);
    grow_i(m->fctpn, gf, 0);
    grow_i(m->cltpn, m->ngstcell, 0);
    grow_i(m->clgrp, m->ngstcell, -1);
    grow_i(m->fcnds, gf, -1);
    grow_i(m->fccls, gf, -1);
    grow_i(m->clnds, m->ngstcell, -1);
    grow_i(m->clfcs, m->ngstcell, -1);
    tm.lap("ghost: grow tables");
    m->fill_ghost();
    tm.lap("ghost: fill");
    return 0;
}

} // extern "C"
