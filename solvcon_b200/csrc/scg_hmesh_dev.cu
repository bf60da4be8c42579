// scg_hmesh_dev.cu -- StaticMesh tables built by GPU kernels (see scg_hmesh_dev.cuh).
//
// What is computed is exactly what scg_hmesh.cpp computes on the host, which is exactly what the reference computes:
// the same face numbering (order of first appearance over the cells), node order inside faces, owner / neighbour
// cells, and every floating-point expression of the metric code in the reference's operation order (the strict build
// compiles with -fmad=false; fp64 division and square root are IEEE on the device).  Only the *how* differs:
//   * candidate faces are bucketed by their smallest node id with a stable radix sort (cub) and matched inside the
//     bucket by one thread per candidate (first matching candidate = the root, like the sequential visit);
//   * roots are numbered by a prefix sum over the candidates in cell order;
//   * ghost entities are numbered by prefix sums over the boundary list instead of running counters.
#include "scg_hmesh_dev.cuh"

#include <cub/cub.cuh>

#include <climits>

namespace scg
{
namespace hdev
{

namespace
{

constexpr int FCMND = 4, CLMND = 8, CLMFC = 6, FCREL = 4;

struct FaceDefD
{
    int nface;
    int fnnd[CLMFC];
    int nd[CLMFC][FCMND]; // 1-based positions in clnds
};

// StaticMesh_interior.cpp:61-330 (add_quadrilateral .. add_pyramid); same table as scg_hmesh.cpp
__constant__ FaceDefD c_facedef[9] = {
    {0, {0}, {{0}}},
    {1, {1}, {{1}}},
    {2, {1, 1}, {{1}, {2}}},
    {4, {2, 2, 2, 2}, {{1, 2}, {2, 3}, {3, 4}, {4, 1}}},
    {3, {2, 2, 2}, {{1, 2}, {2, 3}, {3, 1}}},
    {6, {4, 4, 4, 4, 4, 4}, {{1, 4, 3, 2}, {2, 3, 7, 6}, {5, 6, 7, 8}, {1, 5, 8, 4}, {1, 2, 6, 5}, {3, 4, 8, 7}}},
    {4, {3, 3, 3, 3}, {{1, 3, 2}, {1, 2, 4}, {1, 4, 3}, {2, 3, 4}}},
    {5, {3, 3, 4, 4, 4}, {{1, 2, 3}, {4, 6, 5}, {1, 4, 5, 2}, {1, 3, 6, 4}, {2, 5, 6, 3}}},
    {5, {3, 3, 3, 3, 4}, {{1, 5, 4}, {2, 5, 1}, {3, 5, 2}, {4, 5, 3}, {1, 4, 3, 2}}},
};

__host__ __device__ inline int cell_nnode(int t)
{
    return (1 == t) ? 1 : (2 == t) ? 2 : (3 == t) ? 4 : (4 == t) ? 3 : (5 == t) ? 8 : (6 == t) ? 4 : (7 == t) ? 6 : (8 == t) ? 5 : 0;
}
__device__ inline int face_type_of(int ndim, int nnd)
{
    if (1 == nnd) { return 1; }
    return 2 == ndim ? 2 : (3 == nnd ? 4 : 3);
}

inline unsigned grid_of(int64_t n, int tb = 256) { return static_cast<unsigned>(n > 0 ? (n + tb - 1) / tb : 1); }

// A table in the stored layout: element (idx, j) with idx in [-ngh, nbody)
template <typename T>
struct View
{
    T * p;
    int64_t ngh;
    int cols;
    __device__ __forceinline__ T & operator()(int64_t i, int j = 0) const { return p[(ngh + i) * cols + j]; }
};
template <typename T>
View<T> view(DevBuf<T> & b, int64_t ngh, int cols) { return View<T>{b.p, ngh, cols}; }

// ------------------------------------------------------------------------------------------------
// faces from cells
// ------------------------------------------------------------------------------------------------
__global__ void k_cell_nface(int64_t ncell, int64_t nnode, int32_t const * cltpn, int32_t const * clnds, int32_t * nfc, int * err)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= ncell) { return; }
    int const t = cltpn[i];
    nfc[i] = (t >= 0 && t <= 8) ? c_facedef[t].nface : 0;
    int const nnd = cell_nnode(t);
    for (int q = 1; q <= nnd; ++q)
    {
        int32_t const ind = clnds[i * (CLMND + 1) + q];
        if (ind < 0 || ind >= nnode) { atomicMin(err, static_cast<int>(i)); }
    }
}

__global__ void k_populate(int ndim, int64_t ncell, int32_t const * cltpn, int32_t const * clnds, int32_t const * first, int32_t * clfcs,
                           int32_t * tnds, int32_t * ttpn, int32_t * key4, uint32_t * key0, int32_t * cand)
{
    int64_t const icl = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (icl >= ncell) { return; }
    int const t = cltpn[icl];
    FaceDefD const & fd = c_facedef[(t >= 0 && t <= 8) ? t : 0];
    for (int q = 0; q <= CLMFC; ++q) { clfcs[icl * (CLMFC + 1) + q] = -1; }
    if (0 == fd.nface) { return; }
    clfcs[icl * (CLMFC + 1)] = fd.nface;
    for (int f = 0; f < fd.nface; ++f)
    {
        int64_t const ifc = int64_t(first[icl]) + f;
        clfcs[icl * (CLMFC + 1) + f + 1] = static_cast<int32_t>(ifc);
        ttpn[ifc] = face_type_of(ndim, fd.fnnd[f]);
        tnds[ifc * 5] = fd.fnnd[f];
        int32_t k[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
        for (int q = 0; q < FCMND; ++q)
        {
            int32_t const nd = (q < fd.fnnd[f]) ? clnds[icl * (CLMND + 1) + fd.nd[f][q]] : -1;
            tnds[ifc * 5 + 1 + q] = nd;
            if (q < fd.fnnd[f]) { k[q] = nd; }
        }
        // sorting network for 4 keys
#define SCG_CSWAP(a, b)          \
    if (k[a] > k[b])             \
    {                            \
        int32_t const x_ = k[a]; \
        k[a] = k[b];             \
        k[b] = x_;               \
    }
        SCG_CSWAP(0, 1) SCG_CSWAP(2, 3) SCG_CSWAP(0, 2) SCG_CSWAP(1, 3) SCG_CSWAP(1, 2)
#undef SCG_CSWAP
        for (int q = 0; q < 4; ++q) { key4[ifc * 4 + q] = k[q]; }
        key0[ifc] = static_cast<uint32_t>(k[0]);
        cand[ifc] = static_cast<int32_t>(ifc);
    }
}

// start of the bucket (run of equal smallest node ids) every sorted position lies in: heads, then a running maximum
__global__ void k_heads(int64_t n, uint32_t const * skey, int32_t * head)
{
    int64_t const q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (q >= n) { return; }
    head[q] = (0 == q || skey[q] != skey[q - 1]) ? static_cast<int32_t>(q) : 0;
}

// root of a candidate = the first candidate of its bucket with the same type and the same sorted node tuple
__global__ void k_match(int64_t n, int32_t const * sorted, int32_t const * start, int32_t const * ttpn, int32_t const * key4, int32_t * root)
{
    int64_t const q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (q >= n) { return; }
    int32_t const c = sorted[q];
    int4 const kc = *reinterpret_cast<int4 const *>(key4 + int64_t(c) * 4);
    int const tc = ttpn[c];
    int32_t r = c;
    for (int64_t p = start[q]; p < q; ++p)
    {
        int32_t const a = sorted[p];
        int4 const ka = *reinterpret_cast<int4 const *>(key4 + int64_t(a) * 4);
        if (ttpn[a] == tc && ka.x == kc.x && ka.y == kc.y && ka.z == kc.z && ka.w == kc.w)
        {
            r = a;
            break;
        }
    }
    root[c] = r;
}

__global__ void k_rootflag(int64_t n, int32_t const * root, int32_t * flag)
{
    int64_t const f = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (f < n) { flag[f] = (root[f] == f) ? 1 : 0; }
}

// faces take the data of their root candidate; non-root candidates take their root's number
__global__ void k_faces(int64_t n, int32_t const * root, int32_t * map2, int32_t const * ttpn, int32_t const * tnds, int32_t * fctpn,
                        int32_t * fcnds, int32_t * fccls)
{
    int64_t const f = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (f >= n || root[f] != f) { return; }
    int64_t const j = map2[f];
    fctpn[j] = ttpn[f];
    for (int q = 0; q <= FCMND; ++q) { fcnds[j * (FCMND + 1) + q] = tnds[f * 5 + q]; }
    fccls[j * FCREL + 0] = -1;
    fccls[j * FCREL + 1] = INT_MAX; // lowest second cell wins (atomicMin below); INT_MAX -> -1 afterwards
    fccls[j * FCREL + 2] = -1;
    fccls[j * FCREL + 3] = -1;
}

__global__ void k_clfcs(int64_t ncell, int32_t const * root, int32_t const * map2, int32_t * clfcs, int32_t * fccls)
{
    int64_t const icl = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (icl >= ncell) { return; }
    int const nfc = clfcs[icl * (CLMFC + 1)];
    for (int ifl = 1; ifl <= nfc; ++ifl)
    {
        int32_t const cand = clfcs[icl * (CLMFC + 1) + ifl];
        int32_t const r = root[cand];
        int32_t const jfc = map2[r];
        clfcs[icl * (CLMFC + 1) + ifl] = jfc;
        if (r == cand) { fccls[int64_t(jfc) * FCREL] = static_cast<int32_t>(icl); }
        else { atomicMin(&fccls[int64_t(jfc) * FCREL + 1], static_cast<int32_t>(icl)); }
    }
}

__global__ void k_fccls_fix(int64_t nface, int32_t * fccls)
{
    int64_t const f = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (f < nface && INT_MAX == fccls[f * FCREL + 1]) { fccls[f * FCREL + 1] = -1; }
}

// ------------------------------------------------------------------------------------------------
// metrics: scg_hmesh::face_centroid_normal / cell_centroid / calc_metric, operation by operation
// ------------------------------------------------------------------------------------------------
struct Geo
{
    int ndim;
    View<double> ndcrd, fccnd, fcnml, fcara, clcnd, clvol;
    View<int32_t> fcnds, fccls, clnds, clfcs;
};

__device__ void face_centroid_normal(Geo const & G, int64_t ifc)
{
    if (2 == G.ndim)
    {
        int64_t const ind1 = G.fcnds(ifc, 1), ind2 = G.fcnds(ifc, 2);
        double c0 = G.ndcrd(ind1, 0);
        double c1 = G.ndcrd(ind1, 1);
        c0 += G.ndcrd(ind2, 0);
        c1 += G.ndcrd(ind2, 1);
        c0 /= 2;
        c1 /= 2;
        G.fccnd(ifc, 0) = c0;
        G.fccnd(ifc, 1) = c1;
        double n0 = G.ndcrd(ind2, 1) - G.ndcrd(ind1, 1);
        double n1 = G.ndcrd(ind1, 0) - G.ndcrd(ind2, 0);
        double const ara = sqrt(n0 * n0 + n1 * n1);
        G.fcara(ifc) = ara;
        G.fcnml(ifc, 0) = n0 / ara;
        G.fcnml(ifc, 1) = n1 / ara;
        return;
    }
    double cfd[FCMND + 2][3];
    cfd[0][0] = cfd[0][1] = cfd[0][2] = 0.0;
    int const nnd = G.fcnds(ifc, 0);
    for (int inf = 1; inf <= nnd; ++inf)
    {
        int64_t const ind = G.fcnds(ifc, inf);
        for (int d = 0; d < 3; ++d)
        {
            cfd[inf][d] = G.ndcrd(ind, d);
            cfd[0][d] += G.ndcrd(ind, d);
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
        double const vob = sqrt(dw0 * dw0 + dw1 * dw1 + dw2 * dw2);
        fc[0] += crd[0] * vob;
        fc[1] += crd[1] * vob;
        fc[2] += crd[2] * vob;
        voc += vob;
    }
    for (int d = 0; d < 3; ++d)
    {
        fc[d] /= voc;
        G.fccnd(ifc, d) = fc[d];
    }
    double rv[FCMND][3];
    for (int inf = 0; inf < nnd; ++inf)
    {
        int64_t const ind = G.fcnds(ifc, inf + 1);
        for (int d = 0; d < 3; ++d) { rv[inf][d] = G.ndcrd(ind, d) - fc[d]; }
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
    double const ara = sqrt(nm[0] * nm[0] + nm[1] * nm[1] + nm[2] * nm[2]);
    for (int d = 0; d < 3; ++d) { G.fcnml(ifc, d) = nm[d] / ara; }
    G.fcara(ifc) = ara / 2.0;
}

__device__ void cell_centroid(Geo const & G, int64_t icl)
{
    int const ndim = G.ndim;
    double crd[3] = {0.0, 0.0, 0.0};
    int const nnd = G.clnds(icl, 0);
    for (int inc = 1; inc <= nnd; ++inc)
    {
        int64_t const ind = G.clnds(icl, inc);
        for (int d = 0; d < ndim; ++d) { crd[d] += G.ndcrd(ind, d); }
    }
    for (int d = 0; d < ndim; ++d) { crd[d] /= nnd; }
    double voc = 0.0;
    double cc[3] = {0.0, 0.0, 0.0};
    int const nfc = G.clfcs(icl, 0);
    double const part = ndim + 1;
    for (int ifl = 1; ifl <= nfc; ++ifl)
    {
        int64_t const ifc = G.clfcs(icl, ifl);
        double du[3];
        double dot = 0.0;
        for (int d = 0; d < ndim; ++d) { du[d] = crd[d] - G.fccnd(ifc, d); }
        if (2 == ndim) { dot = du[0] * G.fcnml(ifc, 0) + du[1] * G.fcnml(ifc, 1); }
        else { dot = du[0] * G.fcnml(ifc, 0) + du[1] * G.fcnml(ifc, 1) + du[2] * G.fcnml(ifc, 2); }
        double const vob = fabs(dot) * G.fcara(ifc);
        voc += vob;
        for (int d = 0; d < ndim; ++d)
        {
            double const dv = G.fccnd(ifc, d) + du[d] / part;
            cc[d] += dv * vob;
        }
    }
    for (int d = 0; d < ndim; ++d) { G.clcnd(icl, d) = cc[d] / voc; }
}

// idx0 + q for q in [0, n): body rows with idx0 = 0, ghost rows with idx0 = -n
__global__ void k_face_metric(Geo G, int64_t idx0, int64_t n)
{
    int64_t const q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (q < n) { face_centroid_normal(G, idx0 + q); }
}
__global__ void k_cell_centroid(Geo G, int64_t idx0, int64_t n)
{
    int64_t const q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (q < n) { cell_centroid(G, idx0 + q); }
}

// body cells: StaticMesh_interior.cpp:930-976 (volume by magnitudes; see scg_hmesh::calc_metric for why the order of
// the visits does not matter)
__global__ void k_cell_volume(Geo G, int64_t ncell)
{
    int64_t const icl = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (icl >= ncell) { return; }
    int const ndim = G.ndim;
    double acc = 0.0;
    int const nfc = G.clfcs(icl, 0);
    for (int it = 1; it <= nfc; ++it)
    {
        int64_t const ifc = G.clfcs(icl, it);
        double vol = 0.0;
        for (int d = 0; d < ndim; ++d) { vol += (G.fccnd(ifc, d) - G.clcnd(icl, d)) * G.fcnml(ifc, d); }
        vol *= G.fcara(ifc);
        if (vol < 0.0) { vol = -vol; }
        acc += vol;
    }
    G.clvol(icl) = acc / ndim;
}

__global__ void k_face_orient(Geo G, int64_t nface)
{
    int64_t const ifc = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (ifc >= nface) { return; }
    int const ndim = G.ndim;
    bool flipped = false;
    for (int side = 0; side < 2; ++side)
    {
        int64_t const icl = G.fccls(ifc, side);
        if (icl < 0) { continue; }
        double vol = 0.0;
        for (int d = 0; d < ndim; ++d) { vol += (G.fccnd(ifc, d) - G.clcnd(icl, d)) * G.fcnml(ifc, d); }
        vol *= G.fcara(ifc);
        if (flipped) { vol = -vol; }
        bool const mine = (0 == side);
        bool const flip = (vol < 0.0) ? mine : !mine;
        if (flip) { flipped = !flipped; }
    }
    if (flipped)
    {
        int const nnd = G.fcnds(ifc, 0);
        int32_t tmp[FCMND];
        for (int jt = 0; jt < nnd; ++jt) { tmp[jt] = G.fcnds(ifc, nnd - jt); }
        for (int jt = 0; jt < nnd; ++jt) { G.fcnds(ifc, jt + 1) = tmp[jt]; }
        for (int d = 0; d < ndim; ++d) { G.fcnml(ifc, d) = -G.fcnml(ifc, d); }
    }
}

// ------------------------------------------------------------------------------------------------
// boundary list and ghost entities
// ------------------------------------------------------------------------------------------------
__global__ void k_bflag(int64_t nface, int32_t const * fccls_body, int32_t * flag)
{
    int64_t const f = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (f < nface) { flag[f] = fccls_body[f * FCREL + 1] < 0 ? 1 : 0; }
}
__global__ void k_compact(int64_t n, int32_t const * flag, int32_t const * off, int32_t * out)
{
    int64_t const f = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (f < n && flag[f]) { out[off[f]] = static_cast<int32_t>(f); }
}

// ghost faces / nodes each boundary face brings: StaticMesh_boundary.cpp:185-217
__global__ void k_ghost_count(int64_t nb, int32_t const * bnd, int32_t const * fccls, int32_t const * cltpn, int32_t const * clnds,
                              int32_t const * fcnds, int32_t * gf, int32_t * gn)
{
    int64_t const k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= nb) { return; }
    int64_t const ifc = bnd[k];
    int64_t const icl = fccls[ifc * FCREL];
    int const t = cltpn[icl];
    gf[k] = ((t >= 0 && t <= 8) ? c_facedef[t].nface : 0) - 1;
    gn[k] = clnds[icl * (CLMND + 1)] - fcnds[ifc * (FCMND + 1)];
}

template <typename T>
__global__ void k_fill(int64_t n, T * p, T v)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) { p[i] = v; }
}

struct GhostTabs
{
    int ndim;
    View<double> ndcrd, fccnd, fcnml;
    View<int32_t> fctpn, cltpn, clgrp, fcnds, fccls, clnds, clfcs;
};

// scg_hmesh::fill_ghost, connectivity part (StaticMesh_boundary.cpp:269-410); ghost cell -(k+1), its faces from
// -(1 + offf[k]) downwards and its nodes from -(1 + offn[k]) downwards, in the order of the sequential code
__global__ void k_ghost_fill(GhostTabs G, int64_t nb, int32_t const * bnd, int32_t const * offf, int32_t const * offn)
{
    int64_t const k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= nb) { return; }
    int32_t const igcl = static_cast<int32_t>(-k - 1);
    int32_t const ibfc = bnd[k];
    int32_t const icl = G.fccls(ibfc, 0);
    int32_t ignd = -1 - offn[k], igfc = -1 - offf[k];
    G.cltpn(igcl) = G.cltpn(icl);
    G.clgrp(igcl) = G.clgrp(icl);
    for (int inl = 0; inl <= CLMND; ++inl) { G.clnds(igcl, inl) = G.clnds(icl, inl); }
    int32_t from[CLMND], to[CLMND]; // gstndmap of this cell
    int nmap = 0;
    int const ncn = G.clnds(icl, 0);
    int const nfn = G.fcnds(ibfc, 0);
    for (int inl = 1; inl <= ncn; ++inl)
    {
        int32_t const ind = G.clnds(icl, inl);
        bool found = false;
        for (int inf = 1; inf <= nfn; ++inf) { found = found || (ind == G.fcnds(ibfc, inf)); }
        if (!found)
        {
            from[nmap] = ind;
            to[nmap] = ignd;
            ++nmap;
            G.clnds(igcl, inl) = ignd;
            double dist = 0.0;
            for (int d = 0; d < G.ndim; ++d) { dist += (G.fccnd(ibfc, d) - G.ndcrd(ind, d)) * G.fcnml(ibfc, d); }
            for (int d = 0; d < G.ndim; ++d) { G.ndcrd(ignd, d) = G.ndcrd(ind, d) + 2 * dist * G.fcnml(ibfc, d); }
            ignd -= 1;
        }
    }
    G.fccls(ibfc, 1) = igcl;
    for (int ifl = 0; ifl <= CLMFC; ++ifl) { G.clfcs(igcl, ifl) = G.clfcs(icl, ifl); }
    int const nfc = G.clfcs(icl, 0);
    for (int ifl = 1; ifl <= nfc; ++ifl)
    {
        int32_t const ifc = G.clfcs(icl, ifl);
        if (ifc == ibfc) { continue; }
        G.fctpn(igfc) = G.fctpn(ifc);
        G.fccls(igfc, 0) = igcl;
        G.clfcs(igcl, ifl) = igfc;
        for (int inf = 0; inf <= FCMND; ++inf) { G.fcnds(igfc, inf) = G.fcnds(ifc, inf); }
        int const nn = G.fcnds(igfc, 0);
        for (int inf = 1; inf <= nn; ++inf)
        {
            int32_t const ind = G.fcnds(igfc, inf);
            for (int q = 0; q < nmap; ++q)
            {
                if (from[q] == ind) { G.fcnds(igfc, inf) = to[q]; }
            }
        }
        igfc -= 1;
    }
}

// ghost cells: volume, and the orientation of the ghost faces they own (StaticMesh_boundary.cpp:602-627)
__global__ void k_ghost_volume(Geo G, int64_t nb)
{
    int64_t const k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= nb) { return; }
    int64_t const icl = -k - 1;
    int const ndim = G.ndim;
    double acc = 0.0;
    int const nfc = G.clfcs(icl, 0);
    for (int it = 1; it <= nfc; ++it)
    {
        int64_t const ifc = G.clfcs(icl, it);
        double vol = 0.0;
        for (int d = 0; d < ndim; ++d) { vol += (G.fccnd(ifc, d) - G.clcnd(icl, d)) * G.fcnml(ifc, d); }
        vol *= G.fcara(ifc);
        if (vol < 0.0)
        {
            if (G.fccls(ifc, 0) == icl)
            {
                for (int d = 0; d < ndim; ++d) { G.fcnml(ifc, d) = -G.fcnml(ifc, d); }
            }
            vol = -vol;
        }
        acc += vol;
    }
    G.clvol(icl) = acc / ndim;
}

template <typename T>
int exclusive_sum(DevBuf<T> & in, DevBuf<T> & out, int64_t n, DevBuf<unsigned char> & tmp)
{
    size_t bytes = 0;
    SCG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in.p, out.p, static_cast<int>(n), stream()));
    if (static_cast<int64_t>(bytes) > tmp.n) { SCG_CUDA(tmp.alloc(static_cast<int64_t>(bytes))); }
    SCG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in.p, out.p, static_cast<int>(n), stream()));
    return 0;
}

Geo geo_of(Tables & t)
{
    Geo G;
    G.ndim = t.ndim;
    G.ndcrd = view(t.ndcrd, t.ngstnode, t.ndim);
    G.fccnd = view(t.fccnd, t.ngstface, t.ndim);
    G.fcnml = view(t.fcnml, t.ngstface, t.ndim);
    G.fcara = view(t.fcara, t.ngstface, 1);
    G.clcnd = view(t.clcnd, t.ngstcell, t.ndim);
    G.clvol = view(t.clvol, t.ngstcell, 1);
    G.fcnds = view(t.fcnds, t.ngstface, FCMND + 1);
    G.fccls = view(t.fccls, t.ngstface, FCREL);
    G.clnds = view(t.clnds, t.ngstcell, CLMND + 1);
    G.clfcs = view(t.clfcs, t.ngstcell, CLMFC + 1);
    return G;
}

// prepend `ng` rows filled with `init` to a stored table of `rows` rows
template <typename T>
int grow_front(DevBuf<T> & b, int64_t rows, int cols, int64_t ng, T init)
{
    DevBuf<T> nb;
    SCG_CUDA(nb.alloc((rows + ng) * cols));
    if (ng > 0) { k_fill<T><<<grid_of(ng * cols), 256, 0, stream()>>>(ng * cols, nb.p, init); }
    if (rows > 0) { SCG_CUDA(cudaMemcpyAsync(nb.p + ng * cols, b.p, sizeof(T) * rows * cols, cudaMemcpyDeviceToDevice, stream())); }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    b.release();
    b.p = nb.p;
    b.n = nb.n;
    nb.p = nullptr;
    nb.n = 0;
    return 0;
}

template <typename T>
__global__ void k_gather(int64_t n, int cols, int64_t ngh, T const * tab, int32_t const * idx, T * out)
{
    int64_t const q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (q >= n) { return; }
    int64_t const r = ngh + idx[q];
    for (int c = 0; c < cols; ++c) { out[q * cols + c] = tab[r * cols + c]; }
}

template <typename T>
int gather_impl(T const * tab, int64_t ngh, int cols, int32_t const * idx, int64_t n, T * out)
{
    if (n <= 0) { return 0; }
    cudaStream_t const st = stream();
    DevBuf<int32_t> d_idx;
    DevBuf<T> d_out;
    SCG_CUDA(d_idx.alloc(n));
    SCG_CUDA(d_out.alloc(n * cols));
    SCG_CUDA(cudaMemcpyAsync(d_idx.p, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    k_gather<T><<<grid_of(n), 256, 0, st>>>(n, cols, ngh, tab, d_idx.p, d_out.p);
    SCG_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(T) * n * cols, cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    SCG_CUDA(cudaGetLastError());
    return 0;
}

} // namespace

int gather_rows(double const * tab, int64_t ngh, int cols, int32_t const * idx, int64_t n, double * out)
{
    return gather_impl<double>(tab, ngh, cols, idx, n, out);
}
int gather_rows(int32_t const * tab, int64_t ngh, int cols, int32_t const * idx, int64_t n, int32_t * out)
{
    return gather_impl<int32_t>(tab, ngh, cols, idx, n, out);
}

int build_interior(Tables & t, int ndim, int64_t nnode, int64_t ncell, double const * ndcrd, int32_t const * cltpn, int32_t const * clgrp,
                   int32_t const * clnds, bool do_metric)
{
    if (int rc = ensure_init()) { return rc; }
    // candidate faces are numbered with int32 (at most 6 per cell)
    SCG_REQUIRE(ncell < (int64_t(1) << 31) / 6 - 1, "StaticMesh (device builder): %lld cells exceed the int32 numbering of the candidate faces",
                static_cast<long long>(ncell));
    cudaStream_t const st = stream();
    // `t` is a fresh object (scg_hmesh.cpp makes a new one for every build)
    t.ndim = ndim;
    t.nnode = nnode;
    t.ncell = ncell;
    SCG_CUDA(t.ndcrd.alloc(nnode * ndim));
    SCG_CUDA(t.cltpn.alloc(ncell));
    SCG_CUDA(t.clgrp.alloc(ncell));
    SCG_CUDA(t.clnds.alloc(ncell * (CLMND + 1)));
    SCG_CUDA(t.clfcs.alloc(ncell * (CLMFC + 1)));
    SCG_CUDA(cudaMemcpyAsync(t.ndcrd.p, ndcrd, sizeof(double) * nnode * ndim, cudaMemcpyHostToDevice, st));
    SCG_CUDA(cudaMemcpyAsync(t.cltpn.p, cltpn, sizeof(int32_t) * ncell, cudaMemcpyHostToDevice, st));
    SCG_CUDA(cudaMemcpyAsync(t.clgrp.p, clgrp, sizeof(int32_t) * ncell, cudaMemcpyHostToDevice, st));
    SCG_CUDA(cudaMemcpyAsync(t.clnds.p, clnds, sizeof(int32_t) * ncell * (CLMND + 1), cudaMemcpyHostToDevice, st));

    DevBuf<unsigned char> tmp;
    DevBuf<int> err;
    SCG_CUDA(err.alloc(1));
    int h_err = INT_MAX;
    SCG_CUDA(cudaMemcpyAsync(err.p, &h_err, sizeof(int), cudaMemcpyHostToDevice, st));
    // ---- candidate faces
    DevBuf<int32_t> nfc, first;
    SCG_CUDA(nfc.alloc(ncell + 1));
    SCG_CUDA(first.alloc(ncell + 1));
    k_cell_nface<<<grid_of(ncell), 256, 0, st>>>(ncell, nnode, t.cltpn.p, t.clnds.p, nfc.p, err.p);
    if (int rc = exclusive_sum(nfc, first, ncell + 1, tmp)) { return rc; }
    int32_t mface32 = 0;
    SCG_CUDA(cudaMemcpyAsync(&mface32, first.p + ncell, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    if (h_err != INT_MAX)
    {
        return fail("StaticMesh::build_interior: a node of cell %d is out of range [0, %lld)", h_err, static_cast<long long>(nnode));
    }
    int64_t const mface = mface32;
    DevBuf<int32_t> tnds, ttpn, key4, cand, sorted, root, map2, flag;
    DevBuf<uint32_t> key0, skey;
    SCG_CUDA(tnds.alloc(mface * 5));
    SCG_CUDA(ttpn.alloc(mface));
    SCG_CUDA(key4.alloc(mface * 4));
    SCG_CUDA(key0.alloc(mface));
    SCG_CUDA(skey.alloc(mface));
    SCG_CUDA(cand.alloc(mface));
    SCG_CUDA(sorted.alloc(mface));
    k_populate<<<grid_of(ncell), 256, 0, st>>>(ndim, ncell, t.cltpn.p, t.clnds.p, first.p, t.clfcs.p, tnds.p, ttpn.p, key4.p, key0.p, cand.p);
    if (mface > 0)
    {
        int bits = 1;
        while ((int64_t(1) << bits) < nnode + 1 && bits < 32) { ++bits; }
        size_t bytes = 0;
        SCG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key0.p, skey.p, cand.p, sorted.p, static_cast<int>(mface), 0, bits, st));
        if (static_cast<int64_t>(bytes) > tmp.n) { SCG_CUDA(tmp.alloc(static_cast<int64_t>(bytes))); }
        SCG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key0.p, skey.p, cand.p, sorted.p, static_cast<int>(mface), 0, bits, st));
    }
    // ---- duplicates: bucket starts by a running maximum over the run heads, then one thread per candidate
    DevBuf<int32_t> head, start;
    SCG_CUDA(head.alloc(mface));
    SCG_CUDA(start.alloc(mface));
    SCG_CUDA(root.alloc(mface));
    if (mface > 0)
    {
        k_heads<<<grid_of(mface), 256, 0, st>>>(mface, skey.p, head.p);
        size_t bytes = 0;
        SCG_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, head.p, start.p, cub::Max(), static_cast<int>(mface), st));
        if (static_cast<int64_t>(bytes) > tmp.n) { SCG_CUDA(tmp.alloc(static_cast<int64_t>(bytes))); }
        SCG_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, bytes, head.p, start.p, cub::Max(), static_cast<int>(mface), st));
        k_match<<<grid_of(mface), 256, 0, st>>>(mface, sorted.p, start.p, ttpn.p, key4.p, root.p);
    }
    head.release();
    start.release();
    skey.release();
    key0.release();
    // ---- compact numbering of the roots in order of first appearance
    SCG_CUDA(flag.alloc(mface + 1));
    SCG_CUDA(map2.alloc(mface + 1));
    k_rootflag<<<grid_of(mface), 256, 0, st>>>(mface, root.p, flag.p);
    if (int rc = exclusive_sum(flag, map2, mface + 1, tmp)) { return rc; }
    int32_t nface32 = 0;
    SCG_CUDA(cudaMemcpyAsync(&nface32, map2.p + mface, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    int64_t const nface = nface32;
    t.nface = nface;
    SCG_CUDA(t.fctpn.alloc(nface));
    SCG_CUDA(t.fcnds.alloc(nface * (FCMND + 1)));
    SCG_CUDA(t.fccls.alloc(nface * FCREL));
    k_faces<<<grid_of(mface), 256, 0, st>>>(mface, root.p, map2.p, ttpn.p, tnds.p, t.fctpn.p, t.fcnds.p, t.fccls.p);
    k_clfcs<<<grid_of(ncell), 256, 0, st>>>(ncell, root.p, map2.p, t.clfcs.p, t.fccls.p);
    k_fccls_fix<<<grid_of(nface), 256, 0, st>>>(nface, t.fccls.p);
    SCG_CUDA(t.fccnd.alloc_zero(nface * ndim));
    SCG_CUDA(t.fcnml.alloc_zero(nface * ndim));
    SCG_CUDA(t.fcara.alloc_zero(nface));
    SCG_CUDA(t.clcnd.alloc_zero(ncell * ndim));
    SCG_CUDA(t.clvol.alloc_zero(ncell));
    if (do_metric)
    {
        Geo const G = geo_of(t);
        if (ndim >= 2)
        {
            k_face_metric<<<grid_of(nface, 128), 128, 0, st>>>(G, 0, nface);
            k_cell_centroid<<<grid_of(ncell, 128), 128, 0, st>>>(G, 0, ncell);
        }
        k_cell_volume<<<grid_of(ncell, 128), 128, 0, st>>>(G, ncell);
        k_face_orient<<<grid_of(nface, 128), 128, 0, st>>>(G, nface);
    }
    SCG_CUDA(cudaStreamSynchronize(st));
    SCG_CUDA(cudaGetLastError());
    t.interior = true;
    return 0;
}

int boundary_faces(Tables & t, std::vector<int32_t> & faces)
{
    cudaStream_t const st = stream();
    int64_t const nface = t.nface;
    DevBuf<int32_t> flag, off, out;
    DevBuf<unsigned char> tmp;
    SCG_CUDA(flag.alloc(nface + 1));
    SCG_CUDA(off.alloc(nface + 1));
    k_bflag<<<grid_of(nface), 256, 0, st>>>(nface, t.fccls.p + t.ngstface * FCREL, flag.p);
    if (int rc = exclusive_sum(flag, off, nface + 1, tmp)) { return rc; }
    int32_t nb = 0;
    SCG_CUDA(cudaMemcpyAsync(&nb, off.p + nface, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    faces.resize(static_cast<size_t>(nb));
    if (nb > 0)
    {
        SCG_CUDA(out.alloc(nb));
        k_compact<<<grid_of(nface), 256, 0, st>>>(nface, flag.p, off.p, out.p);
        SCG_CUDA(cudaMemcpyAsync(faces.data(), out.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        SCG_CUDA(cudaStreamSynchronize(st));
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int build_ghost(Tables & t, int32_t const * bnd, int64_t nbound)
{
    cudaStream_t const st = stream();
    t.nbound = nbound;
    DevBuf<int32_t> d_bnd, gf, gn, offf, offn;
    DevBuf<unsigned char> tmp;
    SCG_CUDA(d_bnd.alloc(nbound + 1));
    SCG_CUDA(gf.alloc(nbound + 1));
    SCG_CUDA(gn.alloc(nbound + 1));
    SCG_CUDA(offf.alloc(nbound + 1));
    SCG_CUDA(offn.alloc(nbound + 1));
    if (nbound > 0) { SCG_CUDA(cudaMemcpyAsync(d_bnd.p, bnd, sizeof(int32_t) * nbound, cudaMemcpyHostToDevice, st)); }
    k_fill<int32_t><<<1, 32, 0, st>>>(1, gf.p + nbound, 0);
    k_fill<int32_t><<<1, 32, 0, st>>>(1, gn.p + nbound, 0);
    k_ghost_count<<<grid_of(nbound), 256, 0, st>>>(nbound, d_bnd.p, t.fccls.p, t.cltpn.p, t.clnds.p, t.fcnds.p, gf.p, gn.p);
    if (int rc = exclusive_sum(gf, offf, nbound + 1, tmp)) { return rc; }
    if (int rc = exclusive_sum(gn, offn, nbound + 1, tmp)) { return rc; }
    int32_t tf = 0, tn = 0;
    SCG_CUDA(cudaMemcpyAsync(&tf, offf.p + nbound, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaMemcpyAsync(&tn, offn.p + nbound, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    int64_t const gfc = tf, gnd = tn, gcl = nbound;
    int const nd = t.ndim;
    if (grow_front(t.ndcrd, t.nnode, nd, gnd, 0.0) || grow_front(t.fccnd, t.nface, nd, gfc, 0.0) || grow_front(t.fcnml, t.nface, nd, gfc, 0.0)
        || grow_front(t.fcara, t.nface, 1, gfc, 0.0) || grow_front(t.clcnd, t.ncell, nd, gcl, 0.0) || grow_front(t.clvol, t.ncell, 1, gcl, 0.0)
        || grow_front(t.fctpn, t.nface, 1, gfc, 0) || grow_front(t.cltpn, t.ncell, 1, gcl, 0) || grow_front(t.clgrp, t.ncell, 1, gcl, -1)
        || grow_front(t.fcnds, t.nface, FCMND + 1, gfc, -1) || grow_front(t.fccls, t.nface, FCREL, gfc, -1)
        || grow_front(t.clnds, t.ncell, CLMND + 1, gcl, -1) || grow_front(t.clfcs, t.ncell, CLMFC + 1, gcl, -1))
    {
        return -1;
    }
    t.ngstnode = gnd;
    t.ngstface = gfc;
    t.ngstcell = gcl;
    GhostTabs H;
    H.ndim = nd;
    H.ndcrd = view(t.ndcrd, gnd, nd);
    H.fccnd = view(t.fccnd, gfc, nd);
    H.fcnml = view(t.fcnml, gfc, nd);
    H.fctpn = view(t.fctpn, gfc, 1);
    H.cltpn = view(t.cltpn, gcl, 1);
    H.clgrp = view(t.clgrp, gcl, 1);
    H.fcnds = view(t.fcnds, gfc, FCMND + 1);
    H.fccls = view(t.fccls, gfc, FCREL);
    H.clnds = view(t.clnds, gcl, CLMND + 1);
    H.clfcs = view(t.clfcs, gcl, CLMFC + 1);
    k_ghost_fill<<<grid_of(nbound, 128), 128, 0, st>>>(H, nbound, d_bnd.p, offf.p, offn.p);
    Geo const G = geo_of(t);
    if (nd >= 2)
    {
        k_face_metric<<<grid_of(gfc, 128), 128, 0, st>>>(G, -gfc, gfc);
        k_cell_centroid<<<grid_of(gcl, 128), 128, 0, st>>>(G, -gcl, gcl);
    }
    k_ghost_volume<<<grid_of(gcl, 128), 128, 0, st>>>(G, gcl);
    SCG_CUDA(cudaStreamSynchronize(st));
    SCG_CUDA(cudaGetLastError());
    t.ghost = true;
    return 0;
}

} // namespace hdev
} // namespace scg
