// scg_mesh.cu -- device mesh for the multidim march: the mesh reorderer (type grouping + Morton /
// reverse Cuthill-McKee cell renumbering) and the group-compact neighbour table.
//
// Input is the reference's StaticMesh tables (cpp/solvcon/mesh/StaticMesh.hpp:425-447), ghost rows
// first exactly as SimpleArray stores them (StaticMesh_boundary.cpp:219-234).  The march kernels
// never touch clfcs/fccls/fcnds: the neighbour of cell icl across local face f,
// fcrcl = fccls[ifc,0] + fccls[ifc,1] - icl (StaticMesh.hpp:355-358), is resolved here once into
// nbr[f][cell] in device numbering.
#include "scg_md.cuh"
#include "scg_par.hpp"

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

} // namespace

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
            g.nact = 0;
            g.nif = 0;
            for (int64_t q = k; q < e; ++q)
            {
                g.nact += halo(m->perm[q]) < 2 ? 1 : 0;
                g.nif += halo(m->perm[q]) == 0 ? 1 : 0;
            }
            g.off_nbr = off_nbr;
            g.off_geo = off_geo;
            off_nbr += int64_t(ct_nfc(t)) * g.n;
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
                                          nbr[g.off_nbr + int64_t(f) * g.n + i] = (jcl >= 0) ? m->iperm[jcl] : static_cast<int32_t>(ncell + (-jcl - 1));
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

    // ---- host copies for the boundary-condition setup
    m->h_fccls.alloc(static_cast<size_t>(nface) * 2);
    host_parallel_for(nface, [&](int64_t b, int64_t e)
                      {
                          for (int64_t f = b; f < e; ++f)
                          {
                              m->h_fccls[2 * f] = fccls_b[4 * f];
                              m->h_fccls[2 * f + 1] = fccls_b[4 * f + 1];
                          }
                      });
    m->h_fcnml.copy_from(fcnml + ngstface * ndim, static_cast<size_t>(nface) * ndim);

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
