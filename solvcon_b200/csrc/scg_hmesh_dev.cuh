// scg_hmesh_dev.cuh -- the StaticMesh table builder on the device (interface between scg_hmesh.cpp and scg_hmesh_dev.cu).
//
// Row f1 of SURVEY.md 8(f): faces-from-cells dedup, face / cell metrics, boundary list and ghost mirroring as GPU
// kernels (cpp/solvcon/mesh/StaticMesh_interior.cpp:351-548, 623-976, StaticMesh_boundary.cpp:103-627).  Tables are
// kept in the reference's stored layout -- ghost rows first (StaticMesh_boundary.cpp:219-234) -- so they can be handed
// to the device-mesh layout stage (scg_mesh.cu) without touching the host, and downloaded table by table when a
// caller asks for one.
#pragma once

#include "scg_common.cuh"

#include <vector>

namespace scg
{
namespace hdev
{

struct Tables
{
    int ndim = 0;
    int64_t nnode = 0, nface = 0, ncell = 0, nbound = 0, ngstnode = 0, ngstface = 0, ngstcell = 0;
    // stored arrays: [ngst* + n*][cols]
    DevBuf<double> ndcrd, fccnd, fcnml, fcara, clcnd, clvol;
    DevBuf<int32_t> fctpn, cltpn, clgrp, fcnds, fccls, clnds, clfcs;
    bool interior = false, ghost = false;
};

// Upload nodes + cells (body rows) and build the interior tables; `do_metric` as StaticMesh::build_interior.
int build_interior(Tables & t, int ndim, int64_t nnode, int64_t ncell, double const * ndcrd, int32_t const * cltpn, int32_t const * clgrp,
                   int32_t const * clnds, bool do_metric);
// Boundary faces (fccls[f][1] < 0) in ascending face id.
int boundary_faces(Tables & t, std::vector<int32_t> & faces);
// Ghost cells / faces / nodes for the boundary faces listed in bnd[nbound] (ghost cell -(k+1) mirrors the cell on bnd[k]).
int build_ghost(Tables & t, int32_t const * bnd, int64_t nbound);
// rows idx[0..n) (entity indices, negative = ghost) of a stored device table with `ngh` ghost rows -> host out[n][cols]
int gather_rows(double const * tab, int64_t ngh, int cols, int32_t const * idx, int64_t n, double * out);
int gather_rows(int32_t const * tab, int64_t ngh, int cols, int32_t const * idx, int64_t n, int32_t * out);

} // namespace hdev
} // namespace scg

struct scg_hmesh;
namespace scg
{
namespace hdev
{
// the device tables of a host mesh object, or null (scg_hmesh.cpp)
Tables * tables_of(scg_hmesh * m);
} // namespace hdev
} // namespace scg
