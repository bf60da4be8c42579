// scg_md.cuh -- shared types of the multidim device path (scg_mesh.cu, scg_md.cu).
//
// Device data model:
//   * cells are stored in DEVICE ORDER: grouped by cell type (so every warp is type-uniform and the
//     per-type loops unroll), and inside a group sorted by the chosen key (reference order, Morton
//     key of the centroid, or reverse Cuthill-McKee rank) so neighbour gathers hit L1/L2;
//   * ghost cell g (reference index -(g+1)) is device index ncell + g;
//   * field arrays are SoA with one component per plane of stride `stride` (= padded ncell+ngstcell);
//   * per-cell geometry and connectivity are stored "group compact": a group of n cells with nfc
//     faces / nsf sub-faces holds nbr[nfc][n], cevol[nfc+1][n], cecnd[nfc*ndim][n],
//     sfcnd[nsf*ndim][n], sfnml[nsf*ndim][n] -- nothing is padded to CLMFC*FCMND like the host layout
//     (euler_ce.cpp:47-59), so a warp reads 32 consecutive doubles per component.
#pragma once

#include "scg_common.cuh"
#include "scg_par.hpp"

#include <vector>

namespace scg
{

// cell type ids, cpp/solvcon/mesh/StaticMesh.hpp (CellType)
enum : int
{
    CT_QUAD = 3,
    CT_TRI = 4,
    CT_HEX = 5,
    CT_TET = 6,
    CT_PRISM = 7,
    CT_PYRAMID = 8,
};

constexpr int CLMFC = 6;
constexpr int FCMND = 4;

__host__ __device__ constexpr int ct_ndim(int t) { return (t == CT_QUAD || t == CT_TRI) ? 2 : 3; }
__host__ __device__ constexpr int ct_nfc(int t)
{
    return t == CT_QUAD ? 4 : t == CT_TRI ? 3 : t == CT_HEX ? 6 : t == CT_TET ? 4 : 5;
}
// nodes (= sub-faces) of local face f (0-based), StaticMesh_interior.cpp add_* functions
__host__ __device__ constexpr int ct_fnnd(int t, int f)
{
    return (t == CT_QUAD || t == CT_TRI) ? 2
         : t == CT_HEX                  ? 4
         : t == CT_TET                  ? 3
         : t == CT_PRISM                ? (f < 2 ? 3 : 4)
                                        : (f < 4 ? 3 : 4);
}
__host__ __device__ constexpr int ct_sfoff(int t, int f) // first compact sub-face slot of local face f
{
    int o = 0;
    for (int q = 0; q < f; ++q) { o += ct_fnnd(t, q); }
    return o;
}
__host__ __device__ constexpr int ct_nsf(int t) { return ct_sfoff(t, ct_nfc(t)); }
// number of fundamental gradient elements, GradientElement.cpp:16-79
__host__ __device__ constexpr int ct_nfge(int t)
{
    return t == CT_QUAD ? 4 : t == CT_TRI ? 3 : t == CT_HEX ? 8 : t == CT_TET ? 4 : 6;
}
// algorithmic bytes per cell-substep, SURVEY.md 8(d)
__host__ __device__ constexpr int ct_balg(int t)
{
    int const nd = ct_ndim(t), neq = nd + 2, nfc = ct_nfc(t), nsf = ct_nsf(t);
    return 8 * neq * (1 + nd) + 8 + 8 * (nfc + 1) + 8 * nd * (nfc + 1) + 16 * nd * nsf + 4 * (nfc + 2) // reads
         + 8 * neq + 8 * neq * nd + 16;                                                              // writes
}

struct MeshGroup
{
    int type = 0;
    int32_t beg = 0; // first device cell
    int32_t n = 0;   // cells in the group
    int32_t ld = 0;  // leading dimension of the group's arrays: n padded to a multiple of 32 (every row starts 256-byte aligned)
    int32_t nact = 0; // owned cells (stored first); the rest are halo copies owned by another partition
    int32_t nif = 0;  // owned interface cells (stored first among the owned ones)
    int64_t off_nbr = 0;  // into scg_mesh::d_nbr   (ints)
    int64_t off_geo = 0;  // into scg_md::geo        (doubles): cevol, cecnd, sfcnd, sfnml in this order
    __host__ __device__ int64_t geo_cevol() const { return off_geo; }
    __host__ __device__ int64_t geo_cecnd() const { return off_geo + int64_t(ct_nfc(type) + 1) * ld; }
    __host__ __device__ int64_t geo_sfcnd() const { return geo_cecnd() + int64_t(ct_nfc(type)) * ct_ndim(type) * ld; }
    __host__ __device__ int64_t geo_sfnml() const { return geo_sfcnd() + int64_t(ct_nsf(type)) * ct_ndim(type) * ld; }
    __host__ __device__ int64_t geo_size() const { return geo_sfnml() + int64_t(ct_nsf(type)) * ct_ndim(type) * ld - off_geo; }
};

constexpr int MAX_GROUPS = 8;

struct GroupTable // passed by value to the layout-conversion kernels
{
    int ngroup;
    MeshGroup g[MAX_GROUPS];
};

} // namespace scg

struct scg_mesh;
// fccls[f][0..1] and fcnml[f][:] of the listed body faces, from the device copies (scg_mesh.cu)
int mesh_fetch_faces(scg_mesh * m, int32_t const * faces, int64_t n, int32_t * fccls2, double * fcnml);

struct scg_mesh
{
    int ndim = 0;
    int order = 0;
    int64_t nnode = 0, nface = 0, ncell = 0, ngstnode = 0, ngstface = 0, ngstcell = 0;
    int64_t T = 0;      // ncell + ngstcell
    int64_t stride = 0; // T padded to a multiple of 32
    std::vector<int32_t> perm;  // device slot -> reference cell
    std::vector<int32_t> iperm; // reference cell -> device slot
    std::vector<scg::MeshGroup> groups;
    int64_t geo_total = 0;
    int64_t nbr_total = 0;
    // device
    scg::DevBuf<int32_t> d_perm;   // [ncell]
    scg::DevBuf<int32_t> d_rowmap; // [T] stored host row (ghost rows first) -> device index
    scg::DevBuf<int32_t> d_nbr;    // group compact neighbour device indices
    // raw reference tables kept for prepare_ce (full stored arrays, ghost rows first)
    scg::DevBuf<int32_t> d_clfcs, d_fccls, d_fcnds;
    scg::DevBuf<double> d_ndcrd, d_fccnd, d_clcnd;
    scg::DevBuf<double> d_fcnml; // [nface][ndim] body faces (boundary-condition setup, normal matrices)
    int refcount = 0; // scg_md objects using this mesh
};
