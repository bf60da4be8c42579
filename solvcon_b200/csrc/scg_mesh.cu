// scg_mesh.cu -- device mesh for the multidim march: the mesh reorderer (type grouping + Morton /
// reverse Cuthill-McKee cell renumbering) and the group-compact neighbour table.
//
// Input is the reference's StaticMesh tables (cpp/solvcon/mesh/StaticMesh.hpp:425-447), ghost rows
// first exactly as SimpleArray stores them (StaticMesh_boundary.cpp:219-234).  The march kernels
// never touch clfcs/fccls/fcnds: the neighbour of cell icl across local face f,
// fcrcl = fccls[ifc,0] + fccls[ifc,1] - icl (StaticMesh.hpp:355-358), is resolved here once into
// nbr[f][cell] in device numbering.
#include "scg_hmesh_dev.cuh"
#include "scg_md.cuh"
#include "scg_par.hpp"

#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>

using namespace scg;

namespace
{

inline uint64_t spread3(uint64_t v) // 21 bits -> every third bit
{
    v &= 0x1fffff;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

inline uint64_t spread2(uint64_t v) // 31 bits -> every second bit
{
    v &= 0x7fffffff;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// Morton key of every body cell centroid over the bounding box of the centroids
void morton_keys(int ndim, int64_t ncell, double const * clcnd_body, std::vector<uint64_t> & key)
{
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < ncell; ++i)
    {
        for (int d = 0; d < ndim; ++d)
        {
            double const v = clcnd_body[i * ndim + d];
            lo[d] = v < lo[d] ? v : lo[d];
            hi[d] = v > hi[d] ? v : hi[d];
        }
    }
    double const bits = (3 == ndim) ? 2097151.0 : 2147483647.0;
    double sc[3];
    for (int d = 0; d < ndim; ++d) { sc[d] = (hi[d] > lo[d]) ? bits / (hi[d] - lo[d]) : 0.0; }
    key.resize(static_cast<size_t>(ncell));
    host_parallel_for(ncell, [&](int64_t b, int64_t e)
                      {
                          for (int64_t i = b; i < e; ++i)
                          {
                              uint64_t q[3] = {0, 0, 0};
                              for (int d = 0; d < ndim; ++d) { q[d] = static_cast<uint64_t>((clcnd_body[i * ndim + d] - lo[d]) * sc[d]); }
                              key[i] = (3 == ndim) ? (spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2))
                                                   : (spread2(q[0]) | (spread2(q[1]) << 1));
                          }
                      });
}

// reverse Cuthill-McKee rank of every body cell over the cell-face-cell graph
void rcm_rank(int64_t ncell, int32_t const * clfcs_body, int32_t const * fccls_body, std::vector<uint64_t> & key)
{
    auto neighbour = [&](int64_t icl, int f) -> int32_t
    {
        int32_t const ifc = clfcs_body[icl * 7 + 1 + f];
        return fccls_body[ifc * 4 + 0] + fccls_body[ifc * 4 + 1] - static_cast<int32_t>(icl);
    };
    std::vector<int32_t> degree(static_cast<size_t>(ncell));
    host_parallel_for(ncell, [&](int64_t b, int64_t e)
                      {
                          for (int64_t i = b; i < e; ++i)
                          {
                              int deg = 0;
                              for (int f = 0; f < clfcs_body[i * 7]; ++f) { deg += neighbour(i, f) >= 0 ? 1 : 0; }
                              degree[i] = deg;
                          }
                      });
    std::vector<int32_t> order;
    order.reserve(static_cast<size_t>(ncell));
    std::vector<char> seen(static_cast<size_t>(ncell), 0);
    // seeds in ascending (degree, id) so every component starts from a low-degree cell: a counting sort (degree <= 6)
    std::vector<int32_t> seeds(static_cast<size_t>(ncell));
    {
        int64_t cnt[CLMFC + 2] = {0};
        for (int64_t i = 0; i < ncell; ++i) { cnt[degree[i] + 1] += 1; }
        for (int d = 0; d <= CLMFC; ++d) { cnt[d + 1] += cnt[d]; }
        for (int64_t i = 0; i < ncell; ++i) { seeds[cnt[degree[i]]++] = static_cast<int32_t>(i); }
    }
    for (int32_t const seed : seeds)
    {
        if (seen[seed]) { continue; }
        size_t head = order.size();
        order.push_back(seed);
        seen[seed] = 1;
        while (head < order.size())
        {
            int32_t const c = order[head++];
            int32_t nb[CLMFC];
            int nnb = 0;
            for (int f = 0; f < clfcs_body[int64_t(c) * 7]; ++f)
            {
                int32_t const j = neighbour(c, f);
                if (j >= 0 && !seen[j])
                {
                    seen[j] = 1;
                    nb[nnb++] = j;
                }
            }
            std::sort(nb, nb + nnb, [&](int32_t a, int32_t b) { return degree[a] != degree[b] ? degree[a] < degree[b] : a < b; });
            for (int q = 0; q < nnb; ++q) { order.push_back(nb[q]); }
        }
    }
    key.resize(static_cast<size_t>(ncell));
    for (int64_t r = 0; r < ncell; ++r) { key[order[r]] = static_cast<uint64_t>(ncell - 1 - r); } // reversed
}


// ------------------------------------------------------------------------------------------------
// The same layout stage on the device (default; SCG_MESH_HOST_LAYOUT=1 selects the host code above, which is kept as
// the checker: tests/test_gpu_md.py compares the two permutations).  The raw tables are uploaded first -- the march
// keeps them for prepare_ce anyway -- and ordering keys, the sort into (type, class) groups, perm / iperm, the
// neighbour table and the row map are kernels; only perm / iperm (host twins for the API) and 27 counters come back.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t d_spread3(uint64_t v)
{
    v &= 0x1fffff;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__device__ __forceinline__ uint64_t d_spread2(uint64_t v)
{
    v &= 0x7fffffff;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// order-preserving map double -> int64 (for atomicMin / atomicMax on coordinates)
__device__ __forceinline__ long long d_ordered(double v)
{
    long long const b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffll);
}
inline double h_unordered(long long k)
{
    long long const b = k >= 0 ? k : (k ^ 0x7fffffffffffffffll);
    double v;
    std::memcpy(&v, &b, sizeof(v));
    return v;
}

// error flags: [0] first cell with an unsupported type / face count, [1] first cell with a face index out of range,
// [2] first cell with a neighbour out of range (all "first" = smallest cell id, -1 = none)
__global__ void mesh_validate_kernel(int ndim, int64_t ncell, int32_t const * cltpn_b, int32_t const * clfcs_b, int * err)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= ncell) { return; }
    int const t = cltpn_b[i];
    bool const ok = (2 == ndim) ? (t == CT_QUAD || t == CT_TRI) : (t >= CT_HEX && t <= CT_PYRAMID);
    if (!ok || clfcs_b[i * 7] != ct_nfc(t)) { atomicMin(&err[0], static_cast<int>(i)); }
}

__global__ void mesh_bbox_kernel(int ndim, int64_t ncell, double const * clcnd_b, long long * lohi)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    long long lo[3], hi[3];
    for (int d = 0; d < 3; ++d)
    {
        lo[d] = LLONG_MAX;
        hi[d] = LLONG_MIN;
    }
    if (i < ncell)
    {
        for (int d = 0; d < ndim; ++d) { lo[d] = hi[d] = d_ordered(clcnd_b[i * ndim + d]); }
    }
    for (int d = 0; d < ndim; ++d)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            long long const a = __shfl_xor_sync(0xffffffffu, lo[d], o);
            long long const b = __shfl_xor_sync(0xffffffffu, hi[d], o);
            lo[d] = a < lo[d] ? a : lo[d];
            hi[d] = b > hi[d] ? b : hi[d];
        }
        if (0 == (threadIdx.x & 31))
        {
            atomicMin(&lohi[d], lo[d]);
            atomicMax(&lohi[3 + d], hi[d]);
        }
    }
}

struct MortonBox
{
    double lo[3], sc[3];
};

__global__ void mesh_morton_kernel(int ndim, int64_t ncell, double const * clcnd_b, MortonBox B, uint64_t * key, int32_t * id)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= ncell) { return; }
    uint64_t q[3] = {0, 0, 0};
    for (int d = 0; d < ndim; ++d) { q[d] = static_cast<uint64_t>((clcnd_b[i * ndim + d] - B.lo[d]) * B.sc[d]); }
    key[i] = (3 == ndim) ? (d_spread3(q[0]) | (d_spread3(q[1]) << 1) | (d_spread3(q[2]) << 2)) : (d_spread2(q[0]) | (d_spread2(q[1]) << 1));
    id[i] = static_cast<int32_t>(i);
}

__device__ __forceinline__ int32_t d_neighbour(int32_t const * clfcs_b, int32_t const * fccls_b, int64_t icl, int f)
{
    int32_t const ifc = clfcs_b[icl * 7 + 1 + f];
    return fccls_b[int64_t(ifc) * 4] + fccls_b[int64_t(ifc) * 4 + 1] - static_cast<int32_t>(icl);
}

// ---- reverse Cuthill-McKee, level by level.  Same order as rcm_rank: components are started from the unseen cell
// of smallest (degree, id); the cells of a level are visited in order and each appends its not-yet-seen neighbours in
// ascending (degree, id).  A cell reachable from several cells of a level belongs to the first of them: claim[] holds
// the smallest position among the level's cells that touch it.
__global__ void rcm_degree_kernel(int64_t ncell, int32_t const * clfcs_b, int32_t const * fccls_b, int8_t * degree, int32_t * pos, int32_t * claim)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= ncell) { return; }
    int deg = 0;
    int const nfc = clfcs_b[i * 7];
    for (int f = 0; f < nfc; ++f) { deg += d_neighbour(clfcs_b, fccls_b, i, f) >= 0 ? 1 : 0; }
    degree[i] = static_cast<int8_t>(deg);
    pos[i] = -1;
    claim[i] = INT_MAX;
}

// smallest (degree, id) among the unseen cells -> *best (64-bit atomicMin); ULLONG_MAX when every cell has been seen
__global__ void rcm_seed_kernel(int64_t ncell, int8_t const * degree, int32_t const * pos, unsigned long long * best)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    unsigned long long k = ULLONG_MAX;
    if (i < ncell && pos[i] < 0) { k = (static_cast<unsigned long long>(degree[i]) << 32) | static_cast<unsigned long long>(i); }
    for (int o = 16; o > 0; o >>= 1)
    {
        unsigned long long const a = __shfl_xor_sync(0xffffffffu, k, o);
        k = a < k ? a : k;
    }
    if (0 == (threadIdx.x & 31) && k != ULLONG_MAX) { atomicMin(best, k); }
}

__global__ void rcm_start_kernel(unsigned long long const * best, int32_t at, int32_t * order, int32_t * pos)
{
    int32_t const c = static_cast<int32_t>(*best & 0xffffffffull);
    order[at] = c;
    pos[c] = at;
}

__global__ void rcm_claim_kernel(int32_t lb, int32_t le, int32_t const * order, int32_t const * clfcs_b, int32_t const * fccls_b,
                                 int32_t const * pos, int32_t * claim)
{
    int32_t const q = lb + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= le) { return; }
    int64_t const c = order[q];
    int const nfc = clfcs_b[c * 7];
    for (int f = 0; f < nfc; ++f)
    {
        int32_t const j = d_neighbour(clfcs_b, fccls_b, c, f);
        if (j >= 0 && pos[j] < 0) { atomicMin(&claim[j], q); }
    }
}

// children of the level's cell at position q: its neighbours claimed by q, without repeats, in ascending (degree, id)
__device__ __forceinline__ int rcm_children(int32_t q, int64_t c, int32_t const * clfcs_b, int32_t const * fccls_b, int32_t const * pos,
                                            int32_t const * claim, int8_t const * degree, int32_t (&nb)[CLMFC])
{
    int nnb = 0;
    int const nfc = clfcs_b[c * 7];
    for (int f = 0; f < nfc; ++f)
    {
        int32_t const j = d_neighbour(clfcs_b, fccls_b, c, f);
        if (j < 0 || pos[j] >= 0 || claim[j] != q) { continue; }
        bool dup = false;
        for (int r = 0; r < nnb; ++r) { dup = dup || nb[r] == j; }
        if (!dup) { nb[nnb++] = j; }
    }
    for (int a = 1; a < nnb; ++a) // insertion sort, at most 6 entries
    {
        int32_t const v = nb[a];
        int const dv = degree[v];
        int b = a - 1;
        while (b >= 0 && (degree[nb[b]] > dv || (degree[nb[b]] == dv && nb[b] > v)))
        {
            nb[b + 1] = nb[b];
            --b;
        }
        nb[b + 1] = v;
    }
    return nnb;
}

__global__ void rcm_count_kernel(int32_t lb, int32_t le, int32_t const * order, int32_t const * clfcs_b, int32_t const * fccls_b,
                                 int32_t const * pos, int32_t const * claim, int32_t * cnt)
{
    int32_t const q = lb + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= le) { return; }
    int64_t const c = order[q];
    int n = 0;
    int32_t seen[CLMFC];
    int const nfc = clfcs_b[c * 7];
    for (int f = 0; f < nfc; ++f)
    {
        int32_t const j = d_neighbour(clfcs_b, fccls_b, c, f);
        if (j < 0 || pos[j] >= 0 || claim[j] != q) { continue; }
        bool dup = false;
        for (int r = 0; r < n; ++r) { dup = dup || seen[r] == j; }
        if (!dup) { seen[n++] = j; }
    }
    cnt[q - lb] = n;
}

// pos[] of the children is written by a second kernel: rcm_children of OTHER cells of the level still reads pos[] < 0 here
__global__ void rcm_emit_kernel(int32_t lb, int32_t le, int32_t * order, int32_t const * clfcs_b, int32_t const * fccls_b,
                                int32_t const * pos, int32_t const * claim, int8_t const * degree, int32_t const * off)
{
    int32_t const q = lb + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= le) { return; }
    int64_t const c = order[q];
    int32_t nb[CLMFC];
    int const nnb = rcm_children(q, c, clfcs_b, fccls_b, pos, claim, degree, nb);
    int32_t const at = le + off[q - lb];
    for (int r = 0; r < nnb; ++r) { order[at + r] = nb[r]; }
}

__global__ void rcm_mark_kernel(int32_t lb, int32_t le, int32_t const * order, int32_t * pos)
{
    int32_t const q = lb + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= le) { return; }
    pos[order[q]] = q;
}

// by_key[r] = order[ncell-1-r] (reversed), and the sort class of every cell
__global__ void mesh_reverse_kernel(int64_t ncell, int32_t const * order, int32_t * by_key)
{
    int64_t const r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (r < ncell) { by_key[r] = order[ncell - 1 - r]; }
}

__global__ void mesh_iota_kernel(int64_t n, int32_t * id)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) { id[i] = static_cast<int32_t>(i); }
}

// class of the cell in slot r of the locality order: type * 3 + (0 interface, 1 owned interior, 2 halo copy)
__global__ void mesh_class_kernel(int64_t ncell, int32_t const * by_key, int32_t const * cltpn_b, int8_t const * owned, uint8_t * cls,
                                  unsigned long long * hist)
{
    int64_t const r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (r >= ncell) { return; }
    int32_t const c = by_key[r];
    int const h = !owned ? 1 : (2 == owned[c] ? 0 : (owned[c] ? 1 : 2));
    int const k = cltpn_b[c] * 3 + h;
    cls[r] = static_cast<uint8_t>(k);
    atomicAdd(&hist[k], 1ull);
}

__global__ void mesh_iperm_kernel(int64_t ncell, int32_t const * perm, int32_t * iperm)
{
    int64_t const k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k < ncell) { iperm[perm[k]] = static_cast<int32_t>(k); }
}

__global__ void mesh_nbr_kernel(MeshGroup g, int64_t ncell, int64_t nface, int64_t ngstcell, int32_t const * perm, int32_t const * iperm,
                                int32_t const * clfcs_b, int32_t const * fccls_b, int32_t * nbr, int * err)
{
    int64_t const i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= g.n) { return; }
    int64_t const icl = perm[g.beg + i];
    int const nfc = ct_nfc(g.type);
    for (int f = 0; f < nfc; ++f)
    {
        int32_t const ifc = clfcs_b[icl * 7 + 1 + f];
        if (ifc < 0 || ifc >= nface)
        {
            atomicMin(&err[1], static_cast<int>(icl));
            return;
        }
        int32_t const jcl = fccls_b[int64_t(ifc) * 4] + fccls_b[int64_t(ifc) * 4 + 1] - static_cast<int32_t>(icl);
        if (jcl >= ncell || jcl < -ngstcell)
        {
            atomicMin(&err[2], static_cast<int>(icl));
            return;
        }
        nbr[g.off_nbr + int64_t(f) * g.ld + i] = (jcl >= 0) ? iperm[jcl] : static_cast<int32_t>(ncell + (-jcl - 1));
    }
}

__global__ void mesh_rowmap_kernel(int64_t ncell, int64_t ngstcell, int32_t const * iperm, int32_t * rowmap)
{
    int64_t const s = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (s >= ncell + ngstcell) { return; }
    rowmap[s] = (s < ngstcell) ? static_cast<int32_t>(ncell + (ngstcell - 1 - s)) : iperm[s - ngstcell];
}

__global__ void mesh_fetch_faces_kernel(int64_t n, int ndim, int32_t const * faces, int32_t const * fccls_b, double const * fcnml_b,
                                        int32_t * fccls2, double * fcnml)
{
    int64_t const q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (q >= n) { return; }
    int64_t const f = faces[q];
    if (fccls2)
    {
        fccls2[2 * q] = fccls_b[4 * f];
        fccls2[2 * q + 1] = fccls_b[4 * f + 1];
    }
    if (fcnml)
    {
        for (int d = 0; d < ndim; ++d) { fcnml[q * ndim + d] = fcnml_b[f * ndim + d]; }
    }
}

inline unsigned grid_of(int64_t n, int tb = 256) { return static_cast<unsigned>(n > 0 ? (n + tb - 1) / tb : 1); }

constexpr int RCM_TOO_DEEP = 77;

// RCM order of the body cells on the device: d_order[r] = r-th cell of the Cuthill-McKee order (not yet reversed)
int device_rcm(int64_t ncell, int32_t const * clfcs_b, int32_t const * fccls_b, DevBuf<int32_t> & d_order)
{
    DevBuf<int8_t> degree;
    DevBuf<int32_t> pos, claim, cnt, off;
    DevBuf<unsigned long long> best;
    DevBuf<unsigned char> tmp;
    SCG_CUDA(degree.alloc(ncell));
    SCG_CUDA(pos.alloc(ncell));
    SCG_CUDA(claim.alloc(ncell));
    SCG_CUDA(d_order.alloc(ncell));
    SCG_CUDA(best.alloc(1));
    int64_t cap = 1 << 16; // level width the scan buffers are sized for; grown on demand
    size_t tmp_bytes = 0;
    auto size_scan = [&](int64_t n) -> int
    {
        SCG_CUDA(cnt.alloc(n + 1));
        SCG_CUDA(off.alloc(n + 1));
        tmp_bytes = 0;
        SCG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, off.p, static_cast<int>(n + 1), stream()));
        SCG_CUDA(tmp.alloc(static_cast<int64_t>(tmp_bytes)));
        return 0;
    };
    if (int rc = size_scan(cap)) { return rc; }
    cudaStream_t const st = stream();
    rcm_degree_kernel<<<grid_of(ncell), 256, 0, st>>>(ncell, clfcs_b, fccls_b, degree.p, pos.p, claim.p);
    int64_t done = 0, levels = 0;
    while (done < ncell)
    {
        // next component
        unsigned long long const init = ULLONG_MAX;
        SCG_CUDA(cudaMemcpyAsync(best.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));
        rcm_seed_kernel<<<grid_of(ncell), 256, 0, st>>>(ncell, degree.p, pos.p, best.p);
        rcm_start_kernel<<<1, 1, 0, st>>>(best.p, static_cast<int32_t>(done), d_order.p, pos.p);
        int64_t lb = done, le = done + 1;
        while (lb < le)
        {
            int64_t const w = le - lb;
            if (w > cap)
            {
                SCG_CUDA(cudaStreamSynchronize(st));
                cap = w * 2;
                if (int rc = size_scan(cap)) { return rc; }
            }
            rcm_claim_kernel<<<grid_of(w), 256, 0, st>>>(static_cast<int32_t>(lb), static_cast<int32_t>(le), d_order.p, clfcs_b, fccls_b, pos.p, claim.p);
            rcm_count_kernel<<<grid_of(w), 256, 0, st>>>(static_cast<int32_t>(lb), static_cast<int32_t>(le), d_order.p, clfcs_b, fccls_b, pos.p, claim.p, cnt.p);
            // w + 1 entries: off[w] = number of children of the whole level (cnt[w] is never read by the sum)
            SCG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, off.p, static_cast<int>(w + 1), st));
            rcm_emit_kernel<<<grid_of(w), 256, 0, st>>>(static_cast<int32_t>(lb), static_cast<int32_t>(le), d_order.p, clfcs_b, fccls_b, pos.p, claim.p, degree.p, off.p);
            int32_t total = 0;
            SCG_CUDA(cudaMemcpyAsync(&total, off.p + w, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            SCG_CUDA(cudaStreamSynchronize(st));
            if (total > 0)
            {
                rcm_mark_kernel<<<grid_of(total), 256, 0, st>>>(static_cast<int32_t>(le), static_cast<int32_t>(le + total), d_order.p, pos.p);
            }
            lb = le;
            le += total;
            // One host round trip per level: a graph that is a long chain (a strip of cells) has as many levels as cells.
            // Such graphs go back to the sequential host code, which is linear in the cells whatever the depth.
            if (++levels > 16384 && le < levels * 64) { return RCM_TOO_DEEP; }
        }
        done = le;
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}


// Upload the raw tables, then build ordering, groups, perm / iperm, neighbour table and row map on the device.
// `on_device`: the eight table pointers are device pointers (tables built by scg_hmesh_dev.cu); they are copied device
// to device, because the mesh object owns its tables.  `owned` is always a host pointer.
int layout_on_device(scg_mesh * m, int32_t const * cltpn, int32_t const * clfcs, int32_t const * fccls, int32_t const * fcnds,
                     double const * ndcrd, double const * fccnd, double const * fcnml, double const * clcnd, int8_t const * owned,
                     bool on_device = false)
{
    HostStageTimer tm("dmesh");
    int const ndim = m->ndim;
    int64_t const ncell = m->ncell, nface = m->nface, nnode = m->nnode;
    int64_t const ngstcell = m->ngstcell, ngstface = m->ngstface, ngstnode = m->ngstnode;
    cudaStream_t const st = stream();
    cudaMemcpyKind const kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    auto up_i = [&](DevBuf<int32_t> & b, int32_t const * h, int64_t n) -> int
    {
        SCG_CUDA(b.alloc(n));
        SCG_CUDA(cudaMemcpyAsync(b.p, h, sizeof(int32_t) * n, kind, st));
        return 0;
    };
    auto up_d = [&](DevBuf<double> & b, double const * h, int64_t n) -> int
    {
        SCG_CUDA(b.alloc(n));
        SCG_CUDA(cudaMemcpyAsync(b.p, h, sizeof(double) * n, kind, st));
        return 0;
    };
    DevBuf<int32_t> d_cltpn;
    DevBuf<int8_t> d_owned;
    if (up_i(m->d_clfcs, clfcs, m->T * 7) || up_i(m->d_fccls, fccls, (ngstface + nface) * 4) || up_i(d_cltpn, cltpn, m->T)
        || up_d(m->d_clcnd, clcnd, m->T * ndim))
    {
        return -1;
    }
    if (owned)
    {
        SCG_CUDA(d_owned.alloc(ncell));
        SCG_CUDA(cudaMemcpyAsync(d_owned.p, owned, static_cast<size_t>(ncell), cudaMemcpyHostToDevice, st));
    }
    int32_t const * cltpn_b = d_cltpn.p + ngstcell;
    int32_t const * clfcs_b = m->d_clfcs.p + ngstcell * 7;
    int32_t const * fccls_b = m->d_fccls.p + ngstface * 4;
    double const * clcnd_b = m->d_clcnd.p + ngstcell * ndim;

    DevBuf<int> d_err;
    SCG_CUDA(d_err.alloc(3));
    int h_err[3] = {INT_MAX, INT_MAX, INT_MAX};
    SCG_CUDA(cudaMemcpyAsync(d_err.p, h_err, sizeof(h_err), cudaMemcpyHostToDevice, st));
    mesh_validate_kernel<<<grid_of(ncell), 256, 0, st>>>(ndim, ncell, cltpn_b, clfcs_b, d_err.p);
    SCG_CUDA(cudaMemcpyAsync(h_err, d_err.p, sizeof(h_err), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    if (h_err[0] != INT_MAX)
    {
        int64_t const i = h_err[0];
        int32_t ht = 0, hn = 0;
        SCG_CUDA(cudaMemcpy(&ht, d_cltpn.p + ngstcell + i, sizeof(int32_t), cudaMemcpyDeviceToHost));
        SCG_CUDA(cudaMemcpy(&hn, m->d_clfcs.p + (ngstcell + i) * 7, sizeof(int32_t), cudaMemcpyDeviceToHost));
        return fail("scg_mesh_create: cell %lld has type %d / %d faces, not supported in %dD", static_cast<long long>(i), ht, hn, ndim);
    }
    tm.lap("upload connectivity");

    // ---- cells in ascending locality key (ties in ascending id)
    DevBuf<int32_t> by_key;
    SCG_CUDA(by_key.alloc(ncell));
    DevBuf<unsigned char> tmp;
    if (SCG_ORDER_RCM == m->order)
    {
        DevBuf<int32_t> order;
        int const rc = device_rcm(ncell, clfcs_b, fccls_b, order);
        if (RCM_TOO_DEEP == rc)
        {
            // the sequential code on host copies of the two connectivity tables, its rank uploaded as the cell list
            SCG_CUDA(cudaStreamSynchronize(st));
            std::vector<int32_t> h_clfcs(static_cast<size_t>(ncell) * 7), h_fccls(static_cast<size_t>(nface) * 4);
            SCG_CUDA(cudaMemcpy(h_clfcs.data(), clfcs_b, sizeof(int32_t) * h_clfcs.size(), cudaMemcpyDeviceToHost));
            SCG_CUDA(cudaMemcpy(h_fccls.data(), fccls_b, sizeof(int32_t) * h_fccls.size(), cudaMemcpyDeviceToHost));
            std::vector<uint64_t> key;
            rcm_rank(ncell, h_clfcs.data(), h_fccls.data(), key);
            std::vector<int32_t> h_by_key(static_cast<size_t>(ncell));
            for (int64_t i = 0; i < ncell; ++i) { h_by_key[key[i]] = static_cast<int32_t>(i); }
            SCG_CUDA(cudaMemcpy(by_key.p, h_by_key.data(), sizeof(int32_t) * ncell, cudaMemcpyHostToDevice));
        }
        else if (rc) { return rc; }
        else
        {
            mesh_reverse_kernel<<<grid_of(ncell), 256, 0, st>>>(ncell, order.p, by_key.p);
            SCG_CUDA(cudaStreamSynchronize(st));
        }
    }
    else if (SCG_ORDER_MORTON == m->order)
    {
        DevBuf<long long> lohi;
        SCG_CUDA(lohi.alloc(6));
        long long h_lohi[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
        SCG_CUDA(cudaMemcpyAsync(lohi.p, h_lohi, sizeof(h_lohi), cudaMemcpyHostToDevice, st));
        mesh_bbox_kernel<<<grid_of(ncell), 256, 0, st>>>(ndim, ncell, clcnd_b, lohi.p);
        SCG_CUDA(cudaMemcpyAsync(h_lohi, lohi.p, sizeof(h_lohi), cudaMemcpyDeviceToHost, st));
        SCG_CUDA(cudaStreamSynchronize(st));
        MortonBox B;
        double const bits = (3 == ndim) ? 2097151.0 : 2147483647.0;
        for (int d = 0; d < 3; ++d)
        {
            double const lo = d < ndim ? h_unordered(h_lohi[d]) : 0.0, hi = d < ndim ? h_unordered(h_lohi[3 + d]) : 0.0;
            B.lo[d] = lo;
            B.sc[d] = (hi > lo) ? bits / (hi - lo) : 0.0; // same expression as morton_keys
        }
        DevBuf<uint64_t> key, key2;
        DevBuf<int32_t> id;
        SCG_CUDA(key.alloc(ncell));
        SCG_CUDA(key2.alloc(ncell));
        SCG_CUDA(id.alloc(ncell));
        mesh_morton_kernel<<<grid_of(ncell), 256, 0, st>>>(ndim, ncell, clcnd_b, B, key.p, id.p);
        size_t bytes = 0;
        SCG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key.p, key2.p, id.p, by_key.p, static_cast<int>(ncell), 0, 64, st));
        SCG_CUDA(tmp.alloc(static_cast<int64_t>(bytes)));
        SCG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key.p, key2.p, id.p, by_key.p, static_cast<int>(ncell), 0, 64, st));
        SCG_CUDA(cudaStreamSynchronize(st));
    }
    else { mesh_iota_kernel<<<grid_of(ncell), 256, 0, st>>>(ncell, by_key.p); }
    tm.lap("ordering keys");

    // ---- stable sort by (type, class): perm
    {
        DevBuf<uint8_t> cls, cls2;
        DevBuf<unsigned long long> hist;
        SCG_CUDA(cls.alloc(ncell));
        SCG_CUDA(cls2.alloc(ncell));
        SCG_CUDA(hist.alloc_zero(27));
        SCG_CUDA(m->d_perm.alloc(ncell));
        mesh_class_kernel<<<grid_of(ncell), 256, 0, st>>>(ncell, by_key.p, cltpn_b, owned ? d_owned.p : nullptr, cls.p, hist.p);
        size_t bytes = 0;
        SCG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, cls.p, cls2.p, by_key.p, m->d_perm.p, static_cast<int>(ncell), 0, 5, st));
        SCG_CUDA(tmp.alloc(static_cast<int64_t>(bytes)));
        SCG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, cls.p, cls2.p, by_key.p, m->d_perm.p, static_cast<int>(ncell), 0, 5, st));
        unsigned long long h_hist[27];
        SCG_CUDA(cudaMemcpyAsync(h_hist, hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
        SCG_CUDA(cudaStreamSynchronize(st));
        // ---- groups: one per cell type present (ascending type id); inside: interface, owned interior, halo copies
        int64_t k = 0, off_nbr = 0, off_geo = 0;
        for (int t = 0; t < 9; ++t)
        {
            int64_t const n_t = static_cast<int64_t>(h_hist[t * 3] + h_hist[t * 3 + 1] + h_hist[t * 3 + 2]);
            if (0 == n_t) { continue; }
            MeshGroup g;
            g.type = t;
            g.beg = static_cast<int32_t>(k);
            g.n = static_cast<int32_t>(n_t);
            g.ld = (g.n + 31) / 32 * 32;
            g.nif = static_cast<int32_t>(h_hist[t * 3]);
            g.nact = static_cast<int32_t>(h_hist[t * 3] + h_hist[t * 3 + 1]);
            g.off_nbr = off_nbr;
            g.off_geo = off_geo;
            off_nbr += int64_t(ct_nfc(t)) * g.ld;
            off_geo += g.geo_size();
            m->groups.push_back(g);
            k += n_t;
        }
        m->nbr_total = off_nbr;
        m->geo_total = off_geo;
    }
    if (m->groups.size() > MAX_GROUPS) { return fail("scg_mesh_create: too many cell-type groups"); }
    by_key.release();
    tm.lap("sort into groups");

    // ---- iperm, neighbour table, row map
    DevBuf<int32_t> d_iperm;
    SCG_CUDA(d_iperm.alloc(ncell));
    SCG_CUDA(m->d_nbr.alloc(m->nbr_total));
    SCG_CUDA(m->d_rowmap.alloc(m->T));
    mesh_iperm_kernel<<<grid_of(ncell), 256, 0, st>>>(ncell, m->d_perm.p, d_iperm.p);
    for (auto const & g : m->groups)
    {
        mesh_nbr_kernel<<<grid_of(g.n), 256, 0, st>>>(g, ncell, nface, ngstcell, m->d_perm.p, d_iperm.p, clfcs_b, fccls_b, m->d_nbr.p, d_err.p);
    }
    mesh_rowmap_kernel<<<grid_of(m->T), 256, 0, st>>>(ncell, ngstcell, d_iperm.p, m->d_rowmap.p);
    m->perm.resize(static_cast<size_t>(ncell));
    m->iperm.resize(static_cast<size_t>(ncell));
    SCG_CUDA(cudaMemcpyAsync(m->perm.data(), m->d_perm.p, sizeof(int32_t) * ncell, cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaMemcpyAsync(m->iperm.data(), d_iperm.p, sizeof(int32_t) * ncell, cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaMemcpyAsync(h_err, d_err.p, sizeof(h_err), cudaMemcpyDeviceToHost, st));
    SCG_CUDA(cudaStreamSynchronize(st));
    if (h_err[1] != INT_MAX) { return fail("scg_mesh_create: a face index in clfcs[%lld] is out of range", static_cast<long long>(h_err[1])); }
    if (h_err[2] != INT_MAX)
    {
        return fail("scg_mesh_create: a neighbour of cell %lld is out of range (build_ghost() not run?)", static_cast<long long>(h_err[2]));
    }
    tm.lap("groups + neighbour table");

    // ---- the remaining raw tables (prepare_ce, boundary-condition setup)
    if (up_i(m->d_fcnds, fcnds, (ngstface + nface) * 5) || up_d(m->d_ndcrd, ndcrd, (ngstnode + nnode) * ndim)
        || up_d(m->d_fccnd, fccnd, (ngstface + nface) * ndim) || up_d(m->d_fcnml, fcnml + ngstface * ndim, nface * ndim))
    {
        return -1;
    }
    SCG_CUDA(cudaStreamSynchronize(st));
    SCG_CUDA(cudaGetLastError());
    tm.lap("upload geometry tables");
    return 0;
}

} // namespace

// fccls[f][0..1] (owner, neighbour; reference ids) and fcnml[f][:] of the listed body faces, gathered on the device.
// The boundary-condition setup reads a few per cent of the faces; no host copy of the tables is kept for it.
int mesh_fetch_faces(scg_mesh * m, int32_t const * faces, int64_t n, int32_t * fccls2, double * fcnml)
{
    if (n <= 0) { return 0; }
    cudaStream_t const st = stream();
    DevBuf<int32_t> d_f, d_c;
    DevBuf<double> d_n;
    SCG_CUDA(d_f.alloc(n));
    SCG_CUDA(cudaMemcpyAsync(d_f.p, faces, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    if (fccls2) { SCG_CUDA(d_c.alloc(2 * n)); }
    if (fcnml) { SCG_CUDA(d_n.alloc(m->ndim * n)); }
    mesh_fetch_faces_kernel<<<grid_of(n), 256, 0, st>>>(n, m->ndim, d_f.p, m->d_fccls.p + m->ngstface * 4, m->d_fcnml.p, fccls2 ? d_c.p : nullptr,
                                                       fcnml ? d_n.p : nullptr);
    if (fccls2) { SCG_CUDA(cudaMemcpyAsync(fccls2, d_c.p, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, st)); }
    if (fcnml) { SCG_CUDA(cudaMemcpyAsync(fcnml, d_n.p, sizeof(double) * m->ndim * n, cudaMemcpyDeviceToHost, st)); }
    SCG_CUDA(cudaStreamSynchronize(st));
    SCG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

int scg_mesh_create(scg_mesh ** out, int ndim, int64_t nnode, int64_t nface, int64_t ncell, int64_t ngstnode,
                    int64_t ngstface, int64_t ngstcell, const int32_t * cltpn, const int32_t * clfcs,
                    const int32_t * fccls, const int32_t * fcnds, const double * ndcrd, const double * fccnd,
                    const double * fcnml, const double * clcnd, int order)
{
    return scg_mesh_create_part(out, ndim, nnode, nface, ncell, ngstnode, ngstface, ngstcell, cltpn, clfcs, fccls, fcnds, ndcrd,
                                fccnd, fcnml, clcnd, order, nullptr);
}

int scg_mesh_create_part(scg_mesh ** out, int ndim, int64_t nnode, int64_t nface, int64_t ncell, int64_t ngstnode,
                         int64_t ngstface, int64_t ngstcell, const int32_t * cltpn, const int32_t * clfcs,
                         const int32_t * fccls, const int32_t * fcnds, const double * ndcrd, const double * fccnd,
                         const double * fcnml, const double * clcnd, int order, const int8_t * owned)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_REQUIRE(out != nullptr, "scg_mesh_create: null output");
    SCG_REQUIRE(2 == ndim || 3 == ndim, "scg_mesh_create: ndim must be 2 or 3");
    SCG_REQUIRE(ncell > 0 && nface > 0 && nnode > 0, "scg_mesh_create: empty mesh");
    SCG_REQUIRE(ncell + ngstcell < (int64_t(1) << 31) - 64, "scg_mesh_create: too many cells for int32 indices");
    SCG_REQUIRE(cltpn && clfcs && fccls && fcnds && ndcrd && fccnd && fcnml && clcnd, "scg_mesh_create: null table");
    SCG_REQUIRE(order >= SCG_ORDER_NONE && order <= SCG_ORDER_RCM, "scg_mesh_create: unknown ordering %d", order);

    auto * m = new scg_mesh;
    m->ndim = ndim;
    m->order = order;
    m->nnode = nnode;
    m->nface = nface;
    m->ncell = ncell;
    m->ngstnode = ngstnode;
    m->ngstface = ngstface;
    m->ngstcell = ngstcell;
    m->T = ncell + ngstcell;
    m->stride = (m->T + 31) / 32 * 32;

    // ---- default: the layout stage runs on the device (layout_on_device); the host code below is its checker
    {
        char const * const env = std::getenv("SCG_MESH_HOST_LAYOUT");
        if (!(env && '1' == env[0]))
        {
            if (int rc = layout_on_device(m, cltpn, clfcs, fccls, fcnds, ndcrd, fccnd, fcnml, clcnd, owned))
            {
                delete m;
                return rc;
            }
            *out = m;
            return 0;
        }
    }

    int32_t const * cltpn_b = cltpn + ngstcell;
    int32_t const * clfcs_b = clfcs + ngstcell * 7;
    int32_t const * fccls_b = fccls + ngstface * 4;
    double const * clcnd_b = clcnd + ngstcell * ndim;

    // ---- validate cell types and neighbour ranges
    for (int64_t i = 0; i < ncell; ++i)
    {
        int const t = cltpn_b[i];
        bool const ok = (2 == ndim) ? (t == CT_QUAD || t == CT_TRI) : (t >= CT_HEX && t <= CT_PYRAMID);
        if (!ok || clfcs_b[i * 7] != ct_nfc(t))
        {
            delete m;
            return fail("scg_mesh_create: cell %lld has type %d / %d faces, not supported in %dD",
                        static_cast<long long>(i), t, clfcs_b[i * 7], ndim);
        }
    }

    HostStageTimer tm("dmesh");
    // ---- cell ordering: type groups (ascending type id), inside a group by key
    std::vector<uint64_t> key;
    if (SCG_ORDER_MORTON == order) { morton_keys(ndim, ncell, clcnd_b, key); }
    else if (SCG_ORDER_RCM == order) { rcm_rank(ncell, clfcs_b, fccls_b, key); }
    tm.lap("ordering keys");
    m->perm.resize(static_cast<size_t>(ncell));
    // primary key: cell type; then owned interface cells (owned == 2: their state is sent to another partition, so
    // they are marched first and their halo message travels while the interior cells are marched), owned
    // interior cells, halo copies; then the locality key.  (type, class) has at most 27 values: a stable counting
    // sort, fed in the order of the locality key.
    auto halo = [&](int32_t c) -> int { return !owned ? 1 : (2 == owned[c] ? 0 : (owned[c] ? 1 : 2)); };
    {
        std::vector<int32_t> by_key; // cells in ascending locality key (ties in ascending id); empty = identity
        if (SCG_ORDER_RCM == order)
        {
            // the keys are a permutation of 0 .. ncell-1
            by_key.resize(static_cast<size_t>(ncell));
            host_parallel_for(ncell, [&](int64_t b, int64_t e)
                              {
                                  for (int64_t i = b; i < e; ++i) { by_key[key[i]] = static_cast<int32_t>(i); }
                              });
        }
        else if (!key.empty())
        {
            // Morton: sort (key, id) pairs in one chunk per thread, then merge the chunks pairwise
            std::vector<std::pair<uint64_t, int32_t>> kv(static_cast<size_t>(ncell));
            host_parallel_for(ncell, [&](int64_t b, int64_t e)
                              {
                                  for (int64_t i = b; i < e; ++i) { kv[i] = {key[i], static_cast<int32_t>(i)}; }
                              });
            int64_t const nchunk = (ncell < 200000) ? 1 : host_threads();
            int64_t const clen = (ncell + nchunk - 1) / nchunk;
            host_parallel_for(nchunk, [&](int64_t cb, int64_t ce)
                              {
                                  for (int64_t c = cb; c < ce; ++c)
                                  {
                                      std::sort(kv.begin() + std::min(ncell, c * clen), kv.begin() + std::min(ncell, (c + 1) * clen));
                                  }
                              }, 2);
            for (int64_t w = 1; w < nchunk; w *= 2)
            {
                int64_t const npair = (nchunk + 2 * w - 1) / (2 * w);
                host_parallel_for(npair, [&](int64_t pb, int64_t pe)
                                  {
                                      for (int64_t p = pb; p < pe; ++p)
                                      {
                                          int64_t const lo = std::min(ncell, 2 * p * w * clen), mid = std::min(ncell, (2 * p + 1) * w * clen),
                                                        hi = std::min(ncell, (2 * p + 2) * w * clen);
                                          std::inplace_merge(kv.begin() + lo, kv.begin() + mid, kv.begin() + hi);
                                      }
                                  }, 2);
            }
            by_key.resize(static_cast<size_t>(ncell));
            host_parallel_for(ncell, [&](int64_t b, int64_t e)
                              {
                                  for (int64_t i = b; i < e; ++i) { by_key[i] = kv[i].second; }
                              });
        }
        int64_t cnt[9 * 3 + 1] = {0};
        for (int64_t i = 0; i < ncell; ++i) { cnt[cltpn_b[i] * 3 + halo(static_cast<int32_t>(i)) + 1] += 1; }
        for (int b = 0; b < 27; ++b) { cnt[b + 1] += cnt[b]; }
        for (int64_t r = 0; r < ncell; ++r)
        {
            int32_t const c = by_key.empty() ? static_cast<int32_t>(r) : by_key[r];
            m->perm[cnt[cltpn_b[c] * 3 + halo(c)]++] = c;
        }
    }
    m->iperm.resize(static_cast<size_t>(ncell));
    host_parallel_for(ncell, [&](int64_t b, int64_t e)
                      {
                          for (int64_t k = b; k < e; ++k) { m->iperm[m->perm[k]] = static_cast<int32_t>(k); }
                      });

    tm.lap("sort into groups");
    // ---- groups
    {
        int64_t k = 0;
        int64_t off_nbr = 0, off_geo = 0;
        while (k < ncell)
        {
            int const t = cltpn_b[m->perm[k]];
            int64_t e = k;
            while (e < ncell && cltpn_b[m->perm[e]] == t) { ++e; }
            MeshGroup g;
            g.type = t;
            g.beg = static_cast<int32_t>(k);
            g.n = static_cast<int32_t>(e - k);
            g.ld = (g.n + 31) / 32 * 32;
            g.nact = 0;
            g.nif = 0;
            for (int64_t q = k; q < e; ++q)
            {
                g.nact += halo(m->perm[q]) < 2 ? 1 : 0;
                g.nif += halo(m->perm[q]) == 0 ? 1 : 0;
            }
            g.off_nbr = off_nbr;
            g.off_geo = off_geo;
            off_nbr += int64_t(ct_nfc(t)) * g.ld;
            off_geo += g.geo_size();
            m->groups.push_back(g);
            k = e;
        }
        m->nbr_total = off_nbr;
        m->geo_total = off_geo;
    }
    if (m->groups.size() > MAX_GROUPS)
    {
        delete m;
        return fail("scg_mesh_create: too many cell-type groups");
    }

    // ---- neighbour table in device numbering
    HostBuf<int32_t> nbr;
    nbr.alloc(static_cast<size_t>(m->nbr_total));
    {
        std::atomic<int64_t> bad_cell{-1};
        std::atomic<int> bad_kind{0};
        for (auto const & g : m->groups)
        {
            int const nfc = ct_nfc(g.type);
            host_parallel_for(g.n, [&](int64_t ib, int64_t ie)
                              {
                                  for (int64_t i = ib; i < ie; ++i)
                                  {
                                      int64_t const icl = m->perm[g.beg + i];
                                      for (int f = 0; f < nfc; ++f)
                                      {
                                          int32_t const ifc = clfcs_b[icl * 7 + 1 + f];
                                          if (ifc < 0 || ifc >= nface)
                                          {
                                              bad_cell = icl;
                                              bad_kind = 1;
                                              return;
                                          }
                                          int32_t const jcl = fccls_b[int64_t(ifc) * 4] + fccls_b[int64_t(ifc) * 4 + 1] - static_cast<int32_t>(icl);
                                          if (jcl >= ncell || jcl < -ngstcell)
                                          {
                                              bad_cell = icl;
                                              bad_kind = 2;
                                              return;
                                          }
                                          nbr[g.off_nbr + int64_t(f) * g.ld + i] = (jcl >= 0) ? m->iperm[jcl] : static_cast<int32_t>(ncell + (-jcl - 1));
                                      }
                                  }
                              });
        }
        if (bad_cell >= 0)
        {
            long long const icl = bad_cell;
            int const kind = bad_kind;
            delete m;
            return (1 == kind) ? fail("scg_mesh_create: a face index in clfcs[%lld] is out of range", icl)
                               : fail("scg_mesh_create: a neighbour of cell %lld is out of range (build_ghost() not run?)", icl);
        }
    }

    tm.lap("groups + neighbour table");
    // ---- stored-row -> device-index map (ghost rows first on the host side)
    HostBuf<int32_t> rowmap;
    rowmap.alloc(static_cast<size_t>(m->T));
    for (int64_t s = 0; s < ngstcell; ++s) { rowmap[s] = static_cast<int32_t>(ncell + (ngstcell - 1 - s)); }
    host_parallel_for(ncell, [&](int64_t b, int64_t e)
                      {
                          for (int64_t i = b; i < e; ++i) { rowmap[ngstcell + i] = m->iperm[i]; }
                      });

    tm.lap("rowmap + host copies");
    // ---- device uploads
    bool ok = true;
    auto up_i = [&](DevBuf<int32_t> & b, int32_t const * h, int64_t n)
    {
        ok = ok && cudaSuccess == b.alloc(n)
             && cudaSuccess == cudaMemcpyAsync(b.p, h, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream());
    };
    auto up_d = [&](DevBuf<double> & b, double const * h, int64_t n)
    {
        ok = ok && cudaSuccess == b.alloc(n)
             && cudaSuccess == cudaMemcpyAsync(b.p, h, sizeof(double) * n, cudaMemcpyHostToDevice, stream());
    };
    up_i(m->d_perm, m->perm.data(), ncell);
    up_i(m->d_rowmap, rowmap.data(), m->T);
    up_i(m->d_nbr, nbr.data(), m->nbr_total);
    up_i(m->d_clfcs, clfcs, m->T * 7);
    up_i(m->d_fccls, fccls, (ngstface + nface) * 4);
    up_i(m->d_fcnds, fcnds, (ngstface + nface) * 5);
    up_d(m->d_ndcrd, ndcrd, (ngstnode + nnode) * ndim);
    up_d(m->d_fccnd, fccnd, (ngstface + nface) * ndim);
    up_d(m->d_clcnd, clcnd, m->T * ndim);
    up_d(m->d_fcnml, fcnml + ngstface * ndim, nface * ndim);
    ok = ok && cudaSuccess == cudaStreamSynchronize(stream());
    tm.lap("uploads");
    if (!ok)
    {
        cudaError_t const e = cudaGetLastError();
        delete m;
        return fail("scg_mesh_create: device allocation/upload failed: %s", cudaGetErrorString(e));
    }
    *out = m;
    return 0;
}

// The device mesh of a StaticMesh whose tables were built on the device (scg_hmesh_dev.cu): no table crosses the bus.
int scg_mesh_create_from_hmesh(scg_mesh ** out, scg_hmesh * hm, int order, const int8_t * owned)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_REQUIRE(out && hm, "scg_mesh_create_from_hmesh: null argument");
    scg::hdev::Tables * const t = scg::hdev::tables_of(hm);
    int where = 0;
    if (int rc = scg_hmesh_on_device(hm, &where)) { return rc; }
    SCG_REQUIRE(t && 2 == where, "scg_mesh_create_from_hmesh: the mesh has no current device tables (not built on the device, or a table "
                                 "was handed out to the host since: use scg_mesh_create over the host tables)");
    SCG_REQUIRE(2 == t->ndim || 3 == t->ndim, "scg_mesh_create: ndim must be 2 or 3");
    SCG_REQUIRE(t->ncell > 0 && t->nface > 0 && t->nnode > 0, "scg_mesh_create: empty mesh");
    SCG_REQUIRE(t->ncell + t->ngstcell < (int64_t(1) << 31) - 64, "scg_mesh_create: too many cells for int32 indices");
    SCG_REQUIRE(order >= SCG_ORDER_NONE && order <= SCG_ORDER_RCM, "scg_mesh_create: unknown ordering %d", order);
    auto * m = new scg_mesh;
    m->ndim = t->ndim;
    m->order = order;
    m->nnode = t->nnode;
    m->nface = t->nface;
    m->ncell = t->ncell;
    m->ngstnode = t->ngstnode;
    m->ngstface = t->ngstface;
    m->ngstcell = t->ngstcell;
    m->T = m->ncell + m->ngstcell;
    m->stride = (m->T + 31) / 32 * 32;
    if (int rc = layout_on_device(m, t->cltpn.p, t->clfcs.p, t->fccls.p, t->fcnds.p, t->ndcrd.p, t->fccnd.p, t->fcnml.p, t->clcnd.p, owned, true))
    {
        delete m;
        return rc;
    }
    *out = m;
    return 0;
}

int scg_mesh_destroy(scg_mesh * m)
{
    if (m)
    {
        if (m->refcount > 0) { return fail("scg_mesh_destroy: mesh still used by %d solver(s)", m->refcount); }
        cudaStreamSynchronize(stream());
        delete m;
    }
    return 0;
}

int scg_mesh_perm(scg_mesh * m, int32_t * perm)
{
    SCG_REQUIRE(m && perm, "scg_mesh_perm: null argument");
    std::memcpy(perm, m->perm.data(), sizeof(int32_t) * m->perm.size());
    return 0;
}

} // extern "C"
