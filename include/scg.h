/*
 * scg.h -- C-ABI of libscg_cuda.so: the B200 (sm_100a) device twin of
 * solvcon's space-time CESE march.
 *
 * The reference has no FFI of its own for this path: its boundary is the
 * pybind11 class surface of the `_solvcon` extension.
 * Each entry point below is what the C++ method named beside it forwards to
 * when a maintainer puts the device library under the unchanged classes
 * (INTEGRATION.md shows the binding).  Conventions:
 *
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from scg_last_error() (thread-local).  No exception and no
 *     C++/torch type crosses this boundary.
 *   - handles are opaque; plain pointers + sizes only.
 *   - host arrays use the reference's own SimpleArray layout (row-major,
 *     time planes interleaved in 1D, ghost rows FIRST in the multidim tables)
 *     so a SimpleArray's data() pointer can be passed as is.  The device keeps
 *     its own SoA / reordered layout; upload/download convert.
 *   - all reals are fp64, all indices int32 (reference: cpp/solvcon/base.hpp:71-81).
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef SCG_H
#define SCG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCG_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ runtime */

SCG_API const char * scg_last_error(void);
/* 1 if the library was built without FMA contraction (bit-parity build), 0 for the speed variant */
SCG_API int scg_build_is_strict(void);
SCG_API int scg_device_count(int * count);
/* bind this process/thread to one device and create the library's work stream */
SCG_API int scg_init(int device);
SCG_API int scg_shutdown(void);
SCG_API int scg_sync(void);
/* device memory info in bytes */
SCG_API int scg_mem_info(int64_t * free_bytes, int64_t * total_bytes);
/* kernels launched by this library since the last reset (for bench.py "gpu_launches") */
SCG_API int64_t scg_launch_count(int reset);
/* the library's work stream (a cudaStream_t) so that a caller's communication library can order its own
 * streams against it with events; valid after scg_init */
SCG_API void * scg_stream_handle(void);
/* diagnostics: with SCG_TIMING=1 in the environment every multidim kernel launch is bracketed by events;
 * this prints launches / average / share per kernel to stderr and clears the records */
SCG_API int scg_timing_dump(void);
/* CUDA-event timer on the library's work stream: start, stop -> elapsed ms */
SCG_API int scg_timer_start(void);
SCG_API int scg_timer_stop(double * elapsed_ms);
/* write `bytes` of device scratch (>= L2 size) to evict L2 between timed iterations */
SCG_API int scg_flush_l2(int64_t bytes);
/* pinned host memory for the end-to-end path */
SCG_API int scg_host_alloc(void ** ptr, int64_t bytes);
SCG_API int scg_host_free(void * ptr);
/* raw device scratch and an async copy on the library stream (halo staging between partitions) */
SCG_API int scg_dev_alloc(void ** ptr, int64_t bytes);
SCG_API int scg_dev_free(void * ptr);

SCG_API int scg_dev_copy(void * dst, const void * src, int64_t bytes);

/* --------------------------------------------------------------- 1D scalar
 * replaces spacetime::SolverBase<...> for LinearScalar / InviscidBurgers
 * (cpp/solvcon/spacetime/core.hpp:946-1036, kernel/linear_scalar.hpp:45-87,
 * kernel/inviscid_burgers.hpp:51-109).
 */
typedef struct scg_st scg_st;

#define SCG_EQ_LINEAR 0
#define SCG_EQ_BURGERS 1

#define SCG_PH_SO0 0 /* march_half_so0(odd)        core.hpp:946-956  */
#define SCG_PH_CFL 1 /* update_cfl(odd)            core.hpp:958-967  */
#define SCG_PH_SO1 2 /* march_half_so1_alpha<A>    core.hpp:969-980  */
#define SCG_PH_BC0 3 /* treat_boundary_so0()       core.hpp:982-992  */
#define SCG_PH_BC1 4 /* treat_boundary_so1()       core.hpp:994-1004 */

/* ncelm conservation elements -> nx = 2*ncelm+5 coordinate slots (core.cpp:74) */
SCG_API int scg_st_create(scg_st ** out, int eq, int64_t ncelm, double time_increment);
SCG_API int scg_st_destroy(scg_st * s);
SCG_API int scg_st_set_time_increment(scg_st * s, double time_increment); /* Field::set_time_increment core.cpp:111-116 */
/* fill the grid on the device exactly as Grid::Grid(xmin,xmax,ncelm) does (core.cpp:14-99) */
SCG_API int scg_st_grid_uniform(scg_st * s, double xmin, double xmax);
/* host -> device; any pointer may be NULL (skipped).  Arrays have nx entries. */
SCG_API int scg_st_upload(scg_st * s, const double * x, const double * so0, const double * so1, const double * cfl);
/* device -> host; any pointer may be NULL */
SCG_API int scg_st_download(scg_st * s, double * x, double * so0, double * so1, double * cfl);
SCG_API int scg_st_phase(scg_st * s, int phase, int odd_plane, int alpha);
/* march_alpha<A>(steps), core.hpp:1027-1036; alpha in {0,1,2} */
SCG_API int scg_st_march(scg_st * s, int alpha, int64_t steps);
/* min / max of cfl over the plane points the reference fills, and sum(so0) over the even plane */
SCG_API int scg_st_reduce(scg_st * s, double * cfl_min, double * cfl_max, double * so0_sum);
/* slab decomposition: this object owns CEs [ce_begin, ce_end) of a global ncelm_global grid;
 * halo points are exchanged by the caller through scg_st_halo_pack/unpack */
/* slab mode on/off: when on, march / march_half leave the odd-plane ghosts to scg_st_halo_unpack instead of the
 * periodic copy of treat_boundary_so0/so1 (core.hpp:982-1004) */
SCG_API int scg_st_set_halo_mode(scg_st * s, int on);
/* one half step: 1 = march_half1_alpha (even -> odd plane, core.hpp:1006-1015), 2 = march_half2_alpha (:1017-1025).
 * A slab marches half 1, exchanges {so0, so1} of its first / last odd-plane point with its ring neighbours
 * (halo_pack -> send/recv -> halo_unpack), then marches half 2. */
SCG_API int scg_st_march_half(scg_st * s, int half, int alpha);
/* dev_send_left / right: 2 doubles each on the device ({so0, so1} of the first / last odd-plane point) */
SCG_API int scg_st_halo_pack(scg_st * s, int odd_plane, double * dev_send_left, double * dev_send_right);
SCG_API int scg_st_halo_unpack(scg_st * s, int odd_plane, const double * dev_recv_left, const double * dev_recv_right);
/* march_alpha<A>(steps) of one slab of the periodic ring over the library's communicator (scg_comm_init): per step
 * half 1, the two boundary points to / from the ring neighbours by ncclSend/ncclRecv on the communication stream,
 * half 2.  core.hpp:1006-1036 with treat_boundary_so0/so1 (core.hpp:982-1004) done across ranks. */
SCG_API int scg_st_march_ring(scg_st * s, int alpha, int64_t steps, int left_rank, int right_rank);

/* ---------------------------------------------------------------- 1D Euler
 * replaces onedim::Euler1DCore (cpp/solvcon/onedim/Euler1DCore.hpp:41-145,
 * Euler1DCore.cpp:70-172).
 */
typedef struct scg_e1d scg_e1d;

SCG_API int scg_e1d_create(scg_e1d ** out, int64_t ncoord, double time_increment);
SCG_API int scg_e1d_destroy(scg_e1d * s);
/* arrays: coord[ncoord], gamma[ncoord], so0[ncoord][3], so1[ncoord][3], cfl[ncoord]; NULL = skip */
SCG_API int scg_e1d_upload(scg_e1d * s, const double * coord, const double * gamma, const double * so0,
                           const double * so1, const double * cfl);
SCG_API int scg_e1d_download(scg_e1d * s, double * so0, double * so1, double * cfl);
/* phases as SCG_PH_*: Euler1DCore::{march_half_so0,update_cfl,march_half_so1_alpha<A>,
 * treat_boundary_so0,treat_boundary_so1}; alpha in {1,2} (wrap_onedim.cpp:113-115) */
SCG_API int scg_e1d_phase(scg_e1d * s, int phase, int odd_plane, int alpha);
/* march_alpha<A>(steps), Euler1DCore.hpp:339-349 */
SCG_API int scg_e1d_march(scg_e1d * s, int alpha, int64_t steps);
SCG_API int scg_e1d_reduce(scg_e1d * s, double * cfl_min, double * cfl_max, double * mass);

/* ---- synthetic stepped-channel lattice (solvcon_b200/meshgen.py: mixed_3d), connectivity only, on all host threads.
 * No counterpart in the reference (its only mesher is the Python obsrefl one, _driver.py:57-258); used by bench.py's
 * mesh generator.  See csrc/scg_meshgen.cpp. */
SCG_API int scg_meshgen_mixed3d_count(int64_t nx, int64_t ny, int64_t nz, int64_t x1, int64_t z1, int64_t i0, int64_t i1,
                                      int64_t j0, int64_t j1, int64_t k0, int64_t k1, int64_t * nnode_lattice, int64_t * ncell);
SCG_API int scg_meshgen_mixed3d_cells(int64_t nx, int64_t ny, int64_t nz, int64_t x1, int64_t z1, int64_t i0, int64_t i1,
                                      int64_t j0, int64_t j1, int64_t k0, int64_t k1, const double * xyz_lattice, int32_t * cltpn,
                                      int32_t * clnds, uint8_t * used, int64_t * box_or_null, int64_t * nnode_used);

/* ------------------------------------------------------------- host mesh
 * native twin of solvcon::StaticMesh (cpp/solvcon/mesh/StaticMesh.hpp:270-455): the table builder
 * behind the unchanged `StaticMesh(ndim, nnode, nface, ncell)` / build_interior / add_bc /
 * build_boundary / build_ghost sequence (mesh/pymod/wrap_StaticMesh.cpp:61-175).  Host code only.
 */
typedef struct scg_hmesh scg_hmesh;

SCG_API int scg_hmesh_create(scg_hmesh ** out, int ndim, int64_t nnode, int64_t nface, int64_t ncell);
SCG_API int scg_hmesh_destroy(scg_hmesh * m);
/* which: 0 ndim 1 nnode 2 nface 3 ncell 4 nbound 5 ngstnode 6 ngstface 7 ngstcell 8 nedge 9 nbcs */
SCG_API int scg_hmesh_count(scg_hmesh * m, int which, int64_t * out);
/* zero-copy access to a table by name (ndcrd fccnd fcnml fcara clcnd clvol fctpn cltpn clgrp fcnds fccls
 * clnds clfcs ednds bndfcs): pointer to the first STORED row (ghost rows first), shape, ghost count.
 * Pointers are invalidated by build_interior / build_boundary / build_ghost. */
SCG_API int scg_hmesh_table(scg_hmesh * m, const char * name, void ** ptr, int64_t * rows, int64_t * cols,
                            int64_t * nghost, int * is_real, int * one_d);
SCG_API int scg_hmesh_build_interior(scg_hmesh * m, int do_metric, int do_edge); /* StaticMesh.hpp:363-374 */
SCG_API int scg_hmesh_build_edge(scg_hmesh * m);                                  /* StaticMesh_interior.cpp:566-609 */
SCG_API int scg_hmesh_add_bc(scg_hmesh * m, const char * name, const int32_t * faces, int64_t n); /* StaticMesh_boundary.cpp:14-78 */
SCG_API int scg_hmesh_bc(scg_hmesh * m, int64_t ibc, const char ** name, int32_t ** facn, int64_t * nbound);
SCG_API int scg_hmesh_bc_rename(scg_hmesh * m, int64_t ibc, const char * name); /* StaticMeshBc::set_name, StaticMesh.hpp:243 */
/* Tables handed out by scg_hmesh_table stay valid across build_* calls: a re-made table retires its old buffer
 * instead of freeing it (the reference's views keep the old SimpleArray buffer alive, python/common.hpp:209-215).
 * scg_hmesh_drop_retired frees those old generations; only for callers that hold no earlier view. */
SCG_API int scg_hmesh_drop_retired(scg_hmesh * m);
SCG_API int scg_hmesh_build_boundary(scg_hmesh * m); /* StaticMesh_boundary.cpp:103-182 */
SCG_API int scg_hmesh_build_ghost(scg_hmesh * m);    /* StaticMesh_boundary.cpp:185-627 */
/* Where build_interior / build_boundary / build_ghost run: 1 = GPU kernels (csrc/scg_hmesh_dev.cu: faces-from-cells
 * dedup, metrics, ghost mirroring; tables stay on the device and are downloaded table by table when scg_hmesh_table
 * asks for one), 0 = host threads, -1 = automatic (device when one is present and the mesh has >= 2e5 cells; the
 * environment variable SCG_HMESH_DEVICE=0|1 overrides).  Same tables either way, bit for bit.
 * scg_hmesh_on_device: 0 = host tables, 1 = interior tables on the device, 2 = + ghost entities and no table has been
 * handed out to the host since (a view is writable: from then on the host copies are the mesh). */
SCG_API int scg_hmesh_set_device(scg_hmesh * m, int mode);
/* rows idx[0..n) (entity indices: >= 0 body, < 0 ghost) of the named table -> out[n][cols] (double or int32 like the
 * table), gathered on the device when the table lives there: a caller that needs the boundary faces' centroids of a
 * 1e8-face mesh does not bring the whole table to the host. */
SCG_API int scg_hmesh_gather(scg_hmesh * m, const char * name, const int32_t * idx, int64_t n, void * out);
SCG_API int scg_hmesh_on_device(scg_hmesh * m, int * out);

/* ---------------------------------------------------------- multidim Euler
 * replaces EulerCore over StaticMesh (cpp/solvcon/multidim/euler.hpp:56-211,
 * euler_{ce,march,solution,gradient,bc}.cpp, GradientElement.cpp).
 */
typedef struct scg_mesh scg_mesh;
typedef struct scg_md scg_md;

#define SCG_ORDER_NONE 0   /* reference cell order, grouped by cell type (stable) */
#define SCG_ORDER_MORTON 1 /* Morton (Z-curve) key of the cell centroid within each type group */
#define SCG_ORDER_RCM 2    /* reverse Cuthill-McKee rank within each type group */

/*
 * Upload the StaticMesh tables the march reads (StaticMesh.hpp:425-447).  Every
 * pointer addresses the first STORED row (ghost rows first, as SimpleArray::data()):
 *   cltpn[ngstcell+ncell], clfcs[ngstcell+ncell][7], clcnd[ngstcell+ncell][ndim],
 *   fccls[ngstface+nface][4], fcnds[ngstface+nface][5], fccnd/fcnml[ngstface+nface][ndim],
 *   ndcrd[ngstnode+nnode][ndim].
 * `order` selects the device cell ordering (results are returned in reference order).
 */
SCG_API int scg_mesh_create(scg_mesh ** out, int ndim,
                            int64_t nnode, int64_t nface, int64_t ncell,
                            int64_t ngstnode, int64_t ngstface, int64_t ngstcell,
                            const int32_t * cltpn, const int32_t * clfcs, const int32_t * fccls,
                            const int32_t * fcnds, const double * ndcrd, const double * fccnd,
                            const double * fcnml, const double * clcnd, int order);
/* The mesh of one PARTITION (domain decomposition, one process per GPU): as above plus owned[ncell]
 * (1 = this partition updates the cell, 0 = halo copy of a cell owned by another partition, refreshed by
 * scg_md_halo_unpack).  Halo copies are ordinary body cells of the local StaticMesh; they are stored
 * after the owned cells of their type group and skipped by the march kernels.  owned may be NULL. */
SCG_API int scg_mesh_create_part(scg_mesh ** out, int ndim,
                                 int64_t nnode, int64_t nface, int64_t ncell,
                                 int64_t ngstnode, int64_t ngstface, int64_t ngstcell,
                                 const int32_t * cltpn, const int32_t * clfcs, const int32_t * fccls,
                                 const int32_t * fcnds, const double * ndcrd, const double * fccnd,
                                 const double * fcnml, const double * clcnd, int order, const int8_t * owned);

/* The device mesh of a StaticMesh object (scg_hmesh, above) whose tables were built ON THE DEVICE
 * (scg_hmesh_set_device): the tables go from the builder's kernels to the layout stage without crossing the bus.
 * Same result as scg_mesh_create_part over the downloaded tables. */
struct scg_hmesh;
SCG_API int scg_mesh_create_from_hmesh(scg_mesh ** out, struct scg_hmesh * hmesh, int order, const int8_t * owned);
SCG_API int scg_mesh_destroy(scg_mesh * m);
/* device order <-> reference order permutation: perm[k] = reference cell id stored at device slot k */
SCG_API int scg_mesh_perm(scg_mesh * m, int32_t * perm /* ncell */);

#define SCG_MD_UPDATE 0     /* EulerCore::update()        euler.hpp:123-127         */
#define SCG_MD_SOLT 1       /* calc_solt()                euler_march.cpp:161-191   */
#define SCG_MD_SOLN 2       /* calc_soln()                euler_march.cpp:197-303   */
#define SCG_MD_BC_SOLN 3    /* bc_soln()                  euler_bc.cpp:143-215      */
#define SCG_MD_CFL 4        /* calc_cfl()                 euler_solution.cpp:106-160*/
#define SCG_MD_DSOLN 5      /* calc_dsoln()               euler_gradient.cpp:23-130 */
#define SCG_MD_BC_DSOLN 6   /* bc_dsoln()                 euler_bc.cpp:219-320      */
#define SCG_MD_PREPARE_CE 7 /* prepare_ce()               euler_ce.cpp:62-347       */
#define SCG_MD_SUBSTEP 8    /* march_substep()            euler_march.cpp:339-348   */
/* march_substep of one partition of a decomposed mesh in two stages (additive API): IFACE = update(), calc_solt,
 * then soln/cfl/dsoln of the owned INTERFACE cells (owned[c] == 2 in scg_mesh_create_part) -- their state can be
 * packed and sent right after; REST = the interior cells and the boundary kernels.  IFACE + REST == SUBSTEP. */
#define SCG_MD_SUBSTEP_IFACE 9
#define SCG_MD_SUBSTEP_REST 10

#define SCG_BC_NONREFL 0  /* EulerBC::NonReflective euler.hpp:27-32 */
#define SCG_BC_SLIPWALL 1
#define SCG_BC_INLET 2

/* array ids for upload/download (EulerCore accessors, euler.hpp:105-118) */
#define SCG_A_CEVOL 0
#define SCG_A_CECND 1
#define SCG_A_SFCND 2
#define SCG_A_SFNML 3
#define SCG_A_SO0C 4
#define SCG_A_SO0N 5
#define SCG_A_SO0T 6
#define SCG_A_SO1C 7
#define SCG_A_SO1N 8
#define SCG_A_CFLO 9
#define SCG_A_CFLC 10
#define SCG_A_GAMMA 11
#define SCG_A_COUNT 12

/* EulerCore(mesh, time_increment): allocates the device store and runs prepare_ce on the device */
SCG_API int scg_md_create(scg_md ** out, scg_mesh * mesh, double time_increment);
SCG_API int scg_md_destroy(scg_md * s);
SCG_API int scg_md_set_params(scg_md * s, double sigma0, double taumin, double tauscale); /* euler.hpp:99-104 */
SCG_API int scg_md_set_time_increment(scg_md * s, double time_increment);
/* march kernels: 3 (default) cp.async-staged soln+cfl kernel + equation-outer dsoln kernel;
 * 2 soln+cfl kernel + cooperative (lane-group per cell) dsoln kernel; 1 one fused thread-per-cell kernel;
 * 0 one kernel per reference phase.  All four are bit-identical. */
SCG_API int scg_md_set_fused(scg_md * s, int fused);
/* init_solution(gamma, rho, v, p), euler_solution.cpp:54-99 (v has 3 entries) */
SCG_API int scg_md_init_solution(scg_md * s, double gamma, double rho, const double * v, double p);
/* add_bc(kind, faces, value), euler_bc.cpp:324-389; same validation, same order semantics */
SCG_API int scg_md_add_bc(scg_md * s, int kind, const int32_t * faces, int64_t nface, const double * value, int nvalue);
SCG_API int scg_md_clear_bc(scg_md * s);
/* host array (reference layout, ghost rows first) <-> device store */
SCG_API int scg_md_upload(scg_md * s, int array_id, const double * host);
/* as scg_md_upload for the stored rows [row_begin, row_end) only (host still points at row 0; ghost rows are stored
 * first, so [0, ngstcell) are the ghost rows): march() needs the whole of so0n / so1n / gamma but only the ghost rows
 * of the arrays it overwrites */
SCG_API int scg_md_upload_rows(scg_md * s, int array_id, const double * host, int64_t row_begin, int64_t row_end);
SCG_API int scg_md_download(scg_md * s, int array_id, double * host);
SCG_API int scg_md_phase(scg_md * s, int phase);
/* march(steps) = steps x 2 fused substeps, euler_march.cpp:350-359 */
SCG_API int scg_md_march(scg_md * s, int64_t steps);
/* caller-side reductions of solvcon/multidim/euler/_field.py:138-148 done on the device:
 * min/max of cflo over body cells and sum(density * clvol) */
SCG_API int scg_md_reduce(scg_md * s, const double * clvol_host_or_null, double * cfl_min, double * cfl_max, double * mass);
/* derived body-cell fields of EulerField (solvcon/multidim/euler/_field.py:56-126: density, velx, vely, speed,
 * pressure, mach, total_energy, cfl; velz added) evaluated on the device from the newest solution (so0n) with
 * numpy's operation order, written to out[ncell] in reference cell order */
#define SCG_F_DENSITY 0
#define SCG_F_VELX 1
#define SCG_F_VELY 2
#define SCG_F_VELZ 3
#define SCG_F_SPEED 4
#define SCG_F_PRESSURE 5
#define SCG_F_MACH 6
#define SCG_F_TOTAL_ENERGY 7
#define SCG_F_CFL 8
SCG_API int scg_md_derived(scg_md * s, int field, double * out);
/* algorithmic bytes one substep moves for this mesh, for the roofline */
SCG_API int scg_md_algorithmic_bytes(scg_md * s, int64_t * bytes_per_substep);
/* get_normal_matrix(ifc), euler_bc.cpp:391-425 (host-side helper, row-major ndim x ndim) */
SCG_API int scg_md_normal_matrix(scg_md * s, int32_t ifc, double * out);

/* self test of the shared-reciprocal division used inside the kernels (scg_md_kernels.cuh, DivBy): for every
 * pair, out_fast = the shared-reciprocal quotient, ok = its range test, out_ref = the plain a / b.  Wherever
 * ok is 1 the two must be bit-equal (tests/test_gpu_md.py). */
SCG_API int scg_selftest_div(const double * a, const double * b, double * out_fast, double * out_ref, int32_t * ok, int64_t n);

/* self test of the deferred-range-test square root and reciprocal (scg_md_kernels.cuh, sqrt_fast / rcp_fast):
 * out_fast[0:n] = sqrt_fast(x), out_fast[n:2n] = rcp_fast(x), ok likewise, out_ref = sqrt(x), 1.0 / x.  Wherever
 * ok is 1 fast and ref must be bit-equal (tests/test_gpu_md.py). */
SCG_API int scg_selftest_rsq(const double * x, double * out_fast, double * out_ref, int32_t * ok, int64_t n);

/* halo exchange support for domain decomposition (one process per GPU; the caller moves the packed
 * buffers with NCCL send/recv, see solvcon_b200/decomp.py).  A partition is a StaticMesh of the owned cells
 * plus one layer of HALO copies of the remote face neighbours (scg_mesh_create_part).  send_cells /
 * recv_cells are local reference cell ids, in the order both sides agreed on.  what = 0: the new state
 * so0n+so1n, neq*(1+ndim) doubles per cell, after every substep;  what = 1: solution point + gamma,
 * ndim+1 doubles per cell, once at setup (a halo copy's CE centroid cannot be computed locally: its own
 * outer neighbours are missing).  Buffers are cell-major, buf[i * width + c], so the cells of one neighbour
 * rank (a contiguous range of the lists) form one contiguous message. */
SCG_API int scg_md_halo_set(scg_md * s, const int32_t * send_cells, int64_t nsend, const int32_t * recv_cells, int64_t nrecv);
SCG_API int scg_md_halo_pack(scg_md * s, int what, double ** dev_buffer, int64_t * ndouble);
SCG_API int scg_md_halo_unpack(scg_md * s, int what, const double * dev_buffer);

/* ---- the library's own communicator (csrc/scg_comm.cu) ----------------------------------------------------------
 * No counterpart in the reference (it is single-process).  NCCL point-to-point calls are issued by the
 * library on its own communication stream; libnccl is dlopen'ed (`libnccl_path` or the default soname; a copy already
 * mapped into the process is reused), so the library links neither NCCL nor torch.  The caller only carries the
 * 128-byte unique id from rank 0 to the other ranks (file, socket, MPI, torch.distributed ...). */
SCG_API int scg_comm_unique_id(void * id128, int64_t nbytes, const char * libnccl_path); /* rank 0: ncclGetUniqueId */
SCG_API int scg_comm_init(int rank, int nranks, const void * id128, const char * libnccl_path); /* ncclCommInitRank */
SCG_API int scg_comm_info(int * rank, int * nranks, int * nccl_version);
SCG_API int scg_comm_finalize(void);
/* Neighbour table of a partition: for peer q its rank and its segment [begin, end) of the lists given to
 * scg_md_halo_set (cells).  scg_md_halo_exchange = pack -> one ncclGroup of ncclSend/ncclRecv with every peer ->
 * unpack, stream-ordered (no host synchronisation).  scg_md_march_dist(steps) = EulerCore::march of one partition:
 * per substep the interface cells, the exchange on the communication stream overlapped with the interior cells and
 * the boundary kernels, the unpack (euler_march.cpp:339-359 per partition; bit-identical for any number of parts). */
SCG_API int scg_md_halo_peers(scg_md * s, int npeer, const int32_t * peer_rank, const int64_t * send_begin,
                              const int64_t * send_end, const int64_t * recv_begin, const int64_t * recv_end);
SCG_API int scg_md_halo_exchange(scg_md * s, int what);
SCG_API int scg_md_march_dist(scg_md * s, int64_t steps);

#ifdef __cplusplus
}
#endif

#endif /* SCG_H */
