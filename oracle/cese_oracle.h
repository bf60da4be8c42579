/*
 * cese_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of solvcon's space-time CESE march, used as the
 * parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (solvcon_b200) never links, imports or calls anything declared here.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against
 * the unmodified reference compiled from /root/reference (oracle/_ref, see
 * oracle/Makefile and tests/test_oracle_vs_ref.py) and against the golden
 * vectors the reference's own tests hold (tests/test_oracle_golden.py).
 *
 * All arrays use the reference's host layout (row-major, planes interleaved,
 * ghost rows in front for the multidim tables).  Arithmetic is IEEE double,
 * one operation per source operator, in the reference's expression order;
 * build with -ffp-contract=off (no FMA), like the reference's x86-64 build.
 */
#ifndef CESE_ORACLE_H
#define CESE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------- 1D scalar ---
 * Follows cpp/solvcon/spacetime/core.{hpp,cpp} and kernel/{linear_scalar,
 * inviscid_burgers}.hpp.  nx = 2*ncelm + 5; arrays x, so0, so1, cfl have nx
 * entries (nvar == 1).  eq: 0 linear scalar, 1 inviscid Burgers.
 * phase: 0 march_half_so0, 1 update_cfl, 2 march_half_so1_alpha,
 *        3 treat_boundary_so0, 4 treat_boundary_so1.
 */
int orc_grid_fill(double xmin, double xmax, int64_t ncelm, double * x /* nx */);
int orc_grid_fill_xloc(double const * xloc, int64_t nloc, double * x /* 2*(nloc-1)+5 */);
int orc_st_phase(int eq, int64_t ncelm, double const * x, double * so0, double * so1, double * cfl,
                 double dt, int phase, int odd, int alpha);
int orc_st_march(int eq, int64_t ncelm, double const * x, double * so0, double * so1, double * cfl,
                 double dt, int alpha, int64_t steps);

/* -------------------------------------------------------------- 1D Euler ---
 * Follows cpp/solvcon/onedim/Euler1DCore.{hpp,cpp}.  coord, cfl, gamma have
 * ncoord entries; so0, so1 are [ncoord][3].  Phases as above.
 */
int orc_e1d_phase(int64_t ncoord, double const * coord, double const * gamma, double * so0, double * so1,
                  double * cfl, double dt, int phase, int odd, int alpha);
int orc_e1d_march(int64_t ncoord, double const * coord, double const * gamma, double * so0, double * so1,
                  double * cfl, double dt, int alpha, int64_t steps);

/* -------------------------------------------------------- multidim Euler ---
 * Follows cpp/solvcon/multidim/{euler_ce,euler_march,euler_solution,
 * euler_gradient,euler_bc}.cpp and GradientElement.cpp over the StaticMesh
 * tables of cpp/solvcon/mesh/StaticMesh.hpp.
 *
 * Every table pointer addresses BODY row 0; ghost row g (reference index
 * -(g+1)) lives at negative offsets, exactly like SimpleArray::body().
 */
typedef struct orc_mesh
{
    int32_t ndim;
    int32_t ncell;
    int32_t ngstcell;
    int32_t nface;
    int32_t const * cltpn; /* [T]        cell type                         */
    int32_t const * clfcs; /* [T][7]     col 0 = nfaces                    */
    int32_t const * fccls; /* [F][4]     owner, neighbour                  */
    int32_t const * fcnds; /* [F][5]     col 0 = nnodes                    */
    double const * ndcrd;  /* [N][ndim]                                    */
    double const * fccnd;  /* [F][ndim]                                    */
    double const * fcnml;  /* [F][ndim]                                    */
    double const * clcnd;  /* [T][ndim]                                    */
} orc_mesh;

typedef struct orc_md
{
    orc_mesh const * mesh;
    double dt;
    double sigma0, taumin, tauscale;
    /* geometry (body pointers; cevol/cecnd carry ghost rows, sf* do not) */
    double * cevol; /* [T][7]            */
    double * cecnd; /* [T][7*ndim]       */
    double * sfcnd; /* [ncell][24][ndim] */
    double * sfnml; /* [ncell][24][ndim] */
    /* solution (body pointers, ghost rows in front) */
    double * so0c; /* [T][neq]       */
    double * so0n;
    double * so0t;
    double * so1c; /* [T][neq][ndim] */
    double * so1n;
    double * cflo; /* [T] */
    double * cflc;
    double * gamma;
    /* boundary conditions, in registration order */
    int32_t nbnd;
    int32_t const * bnd_kind;        /* [nbnd] 0 nonrefl 1 slipwall 2 inlet */
    int32_t const * bnd_nface;       /* [nbnd]                              */
    int32_t const * const * bnd_faces; /* [nbnd][nface_i]                   */
    double const * const * bnd_value;  /* [nbnd][ndim+3] (inlet only)       */
} orc_md;

/* phase: 0 update(swap) 1 solt 2 soln 3 bc_soln 4 cfl 5 dsoln 6 bc_dsoln 7 prepare_ce 8 march_substep */
int orc_md_phase(orc_md * md, int phase);
int orc_md_march(orc_md * md, int64_t steps);
/* normal_matrix<NDIM> of euler_bc.cpp:90-128; out is [ndim][ndim] row-major */
void orc_normal_matrix(int ndim, double const * n, double * out);

#ifdef __cplusplus
}
#endif

#endif /* CESE_ORACLE_H */
