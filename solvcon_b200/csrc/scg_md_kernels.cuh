// scg_md_kernels.cuh -- device code of the multidim Euler CESE march.
//
// Device twin of EulerCore's phases (cpp/solvcon/multidim/euler_{ce,march,solution,gradient,bc}.cpp,
// GradientElement.cpp).  Every kernel is a pure gather: one thread owns one output row (a body cell,
// a stored row, or a boundary face) and nothing is accumulated across threads, so results do not
// depend on the launch shape or the cell ordering.  Arithmetic keeps the reference's expression
// order operation by operation; with -fmad=false (the strict build) the results are bit-equal to the
// x86-64 reference.  Work that is provably bit-neutral is removed:
//   * structural zeros / ones of the flux Jacobian are skipped (adding +-0 to a partial sum that
//     started from +0 never changes it; 1*x == x),
//   * adj(D) and det(D) of a gradient element are computed once, not once per equation
//     (euler_gradient.cpp:74-77 recomputes them),
//   * the per-face Taylor difference u_j + hdt*ut_j - u_i + jdis.grad(u_j) is evaluated once per face,
//     not once per gradient element that uses the face.
#pragma once

#include "scg_fastmath.cuh"
#include "scg_md.cuh"

#include <cfloat>
#include <math_constants.h>
#include <type_traits>

namespace scg
{

constexpr double MD_TINY = 1.e-60;         // euler_march.cpp:30,89; euler_solution.cpp:108
constexpr double MD_ALMOST_ZERO = 1.e-200; // euler_gradient.cpp:28

struct MdField // by value to every kernel
{
    double * so0c;
    double * so0n;
    double * so0t;
    double * so0t_next; // fused calc_solt: so0t of the NEXT substep, written by the dsoln / bc_dsoln kernels
    double * sw;        // [neq][stride] so0c + hdt * so0t of every row (SCG_SW_SHARE): what calc_dsoln needs of a neighbour's so0c and
    double * sw_next;   // so0t, formed once by the producer of so0t instead of being gathered as two values per face and equation
    double2 * ct;       // [neq][stride] {so0c, so0t} of every row side by side (SCG_CT_PAIR): md_solncfl_staged gathers both with
    double2 * ct_next;  // ONE 16-byte copy per equation instead of two 8-byte ones; produced wherever so0t / so0t_next is
    double * js;        // [2*ndim+3][stride] the quotients of EulerJacobian::update of every row's so0c (see JacoS::pack);
    double * js_next;   // produced wherever so0t / so0t_next is, consumed by md_solncfl_staged instead of redoing them per face
    double * so1c;
    double * so1n;
    double * gamma;
    double * cflo;
    double * cflc;
    double * solpt; // [ndim][stride] solution point: cecnd[icl,0] for body cells, clcnd for ghosts
    int64_t S;      // component stride
    int32_t ncell;
    int32_t T;
    double dt, sigma0, taumin, tauscale;
    int fuse_solt; // 1: the gradient kernels also evaluate calc_solt for the rows they finish
};

struct MdGeo // one cell-type group
{
    int32_t const * nbr; // [nfc][n]
    double * cevol;      // [nfc+1][n]
    double * cecnd;      // [nfc*ndim][n]  (faces 1..nfc)
    double * sfcnd;      // [nsf*ndim][n]
    double * sfnml;      // [nsf*ndim][n]
    int32_t beg;
    int32_t n;    // array stride of the group (its cell count padded to a multiple of 32)
    int32_t cnt;  // cells in the group
    int32_t nact; // owned cells, stored first: the march kernels stop here
    int32_t ibeg, iend; // cell range of this launch (mode-3 kernels): [0, nact) or the interface / interior part of it
};

// L1 / L2 prefetch of one address per lane (no register is tied up while the line is in flight)
__device__ __forceinline__ void prefetch_l1(void const * p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(void const * p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// zero-numerator mode of the gradient quotients in eqo_solve: the sign of a zero gradient cannot reach so1n
// (grad*grad is +0 either way; o[d] starts at +0 and +0 + (+-0) is +0, x + (+-0) is x)
#ifndef SCG_EQO_ZERO
#define SCG_EQO_ZERO 2
#endif

#ifndef SCG_SW_SHARE
#define SCG_SW_SHARE 1 // md_dsoln_eqo gathers so0c + hdt*so0t (MdField::sw) as one value instead of its two terms
#endif

#ifndef SCG_CT_PAIR
#define SCG_CT_PAIR 0 // 1: md_solncfl_staged gathers {so0c, so0t} of the neighbour as 16-byte pairs (MdField::ct).  Bit-equal; soln 1-4 %
                      // faster, but the producers' extra 16-byte stores take it back: 3D +0.4 %, 2D -1.7 % (profiles/README.md)
#endif

#ifndef SCG_JS_SHARE
#define SCG_JS_SHARE 0 // 1: md_solncfl_staged gathers the Jacobian quotients of the neighbour state (MdField::js) instead of redoing them per
                       // face.  Bit-equal and 140 instructions per face less, but the 9 extra gathered values cost more: hex soln -2 %,
                       // tet soln +22 %, dsoln +3-5 % (the producer), 2D soln +17 % (profiles/README.md)
#endif

#ifndef SCG_SKIP_ZERO_ADD
#define SCG_SKIP_ZERO_ADD 1 // drop the leading "0.0 +" of two ordered sums whose zero sign provably cannot reach a result
#endif

#ifndef SCG_SHARED_RCP
#define SCG_SHARED_RCP 1
#endif

#ifndef SCG_SOLN_PREFETCH
#define SCG_SOLN_PREFETCH 0
#endif

#ifndef SCG_SOLN_EARLY
#define SCG_SOLN_EARLY 0x34 // cell types (bit t - CT_QUAD; default hex, prism, pyramid) whose md_solncfl_staged copies the two values only the
                            // epilogue needs (cevol[0] and gamma of the cell) with the first staged group instead of loading them -- and
                            // waiting one exposed DRAM round trip -- after the face loop
#endif
__host__ __device__ constexpr bool soln_early(int type) { return 0 != ((SCG_SOLN_EARLY >> (type - CT_QUAD)) & 1); }

#ifndef SCG_EQO_SOLT_PF
#define SCG_EQO_SOLT_PF 0 // md_dsoln_eqo with register prefetch: the inputs of the fused calc_solt are loaded at the top of the LAST equation
                          // into the registers of the (then idle) gather prefetch instead of after the loop
#endif

#ifndef SCG_EQO_RCL_SMEM
#define SCG_EQO_RCL_SMEM 0x0f // cell types (bit t - CT_QUAD; all but prisms / pyramids, whose gathers want the L1 more) whose md_dsoln_eqo keeps the neighbour indices in shared memory during the equation
                              // loop.  In registers ptxas hoists the loop-invariant gather offsets out of the loop and spills them: every
                              // equation then starts with a local-memory reload whose L1 lines the gathers have evicted, and only then can
                              // the gathers be issued (profiles/r02g: 12-25 % of the samples of the 3D kernels sit on those two waits)
#endif
__host__ __device__ constexpr int eqo_rcl_smem(int type) { return (SCG_EQO_RCL_SMEM >> (type - CT_QUAD)) & 1; }

// Two more ways of taking loop-invariant values of md_dsoln_eqo's equation loop out of the register file (what does not fit is
// spilled by ptxas and reloaded through L1 at the top of every equation): sigma0 / |cfl| in one shared-memory row, and the
// displacement vectors jdis of the LAST faces in three rows per face.
#ifndef SCG_EQO_SGM_SMEM
#define SCG_EQO_SGM_SMEM 0x08 // cell types (bit t - CT_QUAD): tetrahedra
#endif
#ifndef SCG_EQO_JF_HEX
#define SCG_EQO_JF_HEX 0
#endif
#ifndef SCG_EQO_JF_TET
#define SCG_EQO_JF_TET 0
#endif
#ifndef SCG_EQO_JF_PRISM
#define SCG_EQO_JF_PRISM 0
#endif
__host__ __device__ constexpr int eqo_sgm_smem(int type) { return (SCG_EQO_SGM_SMEM >> (type - CT_QUAD)) & 1; }
__host__ __device__ constexpr int eqo_jdis_faces(int type)
{
    return (type == CT_HEX) ? SCG_EQO_JF_HEX : (type == CT_TET) ? SCG_EQO_JF_TET : (type == CT_PRISM || type == CT_PYRAMID) ? SCG_EQO_JF_PRISM : 0;
}

#ifndef SCG_EQO_OPAQUE_K
#define SCG_EQO_OPAQUE_K 0x0f // cell types (bit t - CT_QUAD; all but prisms / pyramids) whose md_dsoln_eqo makes the row index opaque once per
                              // equation: addresses derived from it are formed again (one IMAD each) instead of being hoisted out of the
                              // loop as 64-bit loop invariants, which ptxas then spills
#endif
__host__ __device__ constexpr bool eqo_opaque_k(int type) { return 0 != ((SCG_EQO_OPAQUE_K >> (type - CT_QUAD)) & 1); }

#ifndef SCG_EQO_PF_AFTER_UDF
#define SCG_EQO_PF_AFTER_UDF 0 // 1: measured equal (hex dsoln 540 vs 543 us): the wait at the top of an equation is not scoreboard sharing
#endif

#ifndef SCG_EQO_PF3_PIN
#define SCG_EQO_PF3_PIN 0
#endif

#ifndef SCG_EQO_GAMMA_PF
#define SCG_EQO_GAMMA_PF 1 // md_dsoln_eqo without register prefetch: gamma of the cell (read by the fused calc_solt only) is requested into L2
                           // at the top of the cell
#endif

#ifndef SCG_XCTA_PF
#define SCG_XCTA_PF 0 // > 0: every CTA requests into L2 the streamed rows (neighbour indices, geometry) that the CTA SCG_XCTA_PF
                      // blocks ahead will read first -- the blocks that replace the resident ones
#endif

// Cross-CTA L2 prefetch (SCG_XCTA_PF = D > 0).  A march kernel starts every cell with two dependent round trips: the streamed rows
// of the cell itself (neighbour indices, geometry, its own state) and then the gathers through the indices.  The first one always
// misses L2 (every row is read once per launch); the CTA that will take this one's place -- about one wave of resident CTAs, D
// blocks, further on -- is known, so its rows are requested now: one lane per 128-byte line, no register held.
__device__ __forceinline__ void xcta_pf_row8(double const * row, int64_t i2) // 8-byte rows: lanes 0 and 16 of a warp cover its 32 cells
{
    if (0 == (threadIdx.x & 15)) { asm volatile("prefetch.global.L2 [%0];" ::"l"(row + i2)); }
}
__device__ __forceinline__ void xcta_pf_row4(int32_t const * row, int64_t i2) // 4-byte rows: lane 0
{
    if (0 == (threadIdx.x & 31)) { asm volatile("prefetch.global.L2 [%0];" ::"l"(row + i2)); }
}

// ------------------------------------------------------------------------------------------------
// flux function and Jacobian, detail::EulerJacobian<2>/<3>::update (euler_march.cpp:35-81, 94-157)
// ------------------------------------------------------------------------------------------------
template <int ND>
struct Jaco
{
    double fcn[ND + 2][ND];
    double jac[ND + 2][ND + 2][ND];
};

// structural kind of jac[i][j][d]: 0 = zero, 1 = one, 2 = general value
template <int ND>
__host__ __device__ constexpr int jkind(int i, int j, int d)
{
    if (0 == i) { return (j == d + 1) ? 1 : 0; }
    if (ND + 1 == i) { return 2; }
    if (0 == j) { return 2; }
    if (ND + 1 == j) { return (d == i - 1) ? 2 : 0; }
    return (i == j || d == i - 1 || d == j - 1) ? 2 : 0;
}

template <int ND, bool FAST>
__device__ __forceinline__ bool jaco_update_impl(Jaco<ND> & J, double ga, double const (&u)[ND + 2]);

template <int ND>
__device__ __forceinline__ void jaco_update(Jaco<ND> & J, double ga, double const (&u)[ND + 2])
{
#if SCG_SHARED_RCP
    if (jaco_update_impl<ND, true>(J, ga, u)) { return; }
#endif
    jaco_update_impl<ND, false>(J, ga, u);
}

template <bool FAST>
__device__ __forceinline__ bool jaco_update_2d(Jaco<2> & J, double ga, double const (&u)[4])
{
    double const ga1 = ga - 1;
    double const ga3 = ga - 3;
    double const ga1h = ga1 / 2;
    double const u1 = u[0] + MD_TINY;
    double const u2 = u[1];
    double const u3 = u[2];
    double const u4 = u[3];
    double const rho2 = u1 * u1;
    DivBy<FAST> by_u1(u1);
    DivBy<FAST> by_rho2(rho2);
    double const v1 = by_u1(u2);
    double const v1o2 = v1 * v1;
    double const v2 = by_u1(u3);
    double const v2o2 = v2 * v2;
    double const ke2 = by_u1(u2 * u2 + u3 * u3);
    double const g1ke2 = ga1 * ke2;
    double const vs = by_u1(ke2);
    double const gretot = ga * u4;
    double const getot = by_u1(gretot);
    double const pr = ga1 * u4 - ga1h * ke2;

    J.fcn[0][0] = u2;              J.fcn[0][1] = u3;
    J.fcn[1][0] = pr + u2 * v1;    J.fcn[1][1] = u2 * v2;
    J.fcn[2][0] = u3 * v1;         J.fcn[2][1] = pr + u3 * v2;
    J.fcn[3][0] = (pr + u4) * v1;  J.fcn[3][1] = (pr + u4) * v2;

    J.jac[1][0][0] = -v1o2 + ga1h * vs;  J.jac[1][0][1] = -v1 * v2;
    J.jac[1][1][0] = -ga3 * v1;          J.jac[1][1][1] = v2;
    J.jac[1][2][0] = -ga1 * v2;          J.jac[1][2][1] = v1;
    J.jac[1][3][0] = ga1;

    J.jac[2][0][0] = -v2 * v1;           J.jac[2][0][1] = -v2o2 + ga1h * vs;
    J.jac[2][1][0] = v2;                 J.jac[2][1][1] = -ga1 * v1;
    J.jac[2][2][0] = v1;                 J.jac[2][2][1] = -ga3 * v2;
    J.jac[2][3][1] = ga1;

    J.jac[3][0][0] = by_rho2((-gretot + g1ke2) * u2);
    J.jac[3][0][1] = by_rho2((-gretot + g1ke2) * u3);
    J.jac[3][1][0] = getot - ga1h * (vs + 2 * v1o2);
    J.jac[3][1][1] = -ga1 * v1 * v2;
    J.jac[3][2][0] = -ga1 * v2 * v1;
    J.jac[3][2][1] = getot - ga1h * (vs + 2 * v2o2);
    J.jac[3][3][0] = ga * v1;
    J.jac[3][3][1] = ga * v2;
    return by_u1.ok && by_rho2.ok;
}

template <bool FAST>
__device__ __forceinline__ bool jaco_update_3d(Jaco<3> & J, double ga, double const (&u)[5])
{
    double const ga1 = ga - 1;
    double const ga3 = ga - 3;
    double const ga1h = ga1 / 2;
    double const u1 = u[0] + MD_TINY;
    double const u2 = u[1];
    double const u3 = u[2];
    double const u4 = u[3];
    double const u5 = u[4];
    double const rho2 = u1 * u1;
    DivBy<FAST> by_u1(u1);
    DivBy<FAST> by_rho2(rho2);
    double const v1 = by_u1(u2);
    double const v1o2 = v1 * v1;
    double const v2 = by_u1(u3);
    double const v2o2 = v2 * v2;
    double const v3 = by_u1(u4);
    double const v3o2 = v3 * v3;
    double const ke2 = by_u1(u2 * u2 + u3 * u3 + u4 * u4);
    double const g1ke2 = ga1 * ke2;
    double const vs = by_u1(ke2);
    double const gretot = ga * u5;
    double const getot = by_u1(gretot);
    double const pr = ga1 * u5 - ga1h * ke2;

    J.fcn[0][0] = u2;             J.fcn[0][1] = u3;             J.fcn[0][2] = u4;
    J.fcn[1][0] = pr + u2 * v1;   J.fcn[1][1] = u2 * v2;        J.fcn[1][2] = u2 * v3;
    J.fcn[2][0] = u3 * v1;        J.fcn[2][1] = pr + u3 * v2;   J.fcn[2][2] = u3 * v3;
    J.fcn[3][0] = u4 * v1;        J.fcn[3][1] = u4 * v2;        J.fcn[3][2] = pr + u4 * v3;
    J.fcn[4][0] = (pr + u5) * v1; J.fcn[4][1] = (pr + u5) * v2; J.fcn[4][2] = (pr + u5) * v3;

    J.jac[1][0][0] = -v1o2 + ga1h * vs; J.jac[1][0][1] = -v1 * v2; J.jac[1][0][2] = -v1 * v3;
    J.jac[1][1][0] = -ga3 * v1;         J.jac[1][1][1] = v2;       J.jac[1][1][2] = v3;
    J.jac[1][2][0] = -ga1 * v2;         J.jac[1][2][1] = v1;
    J.jac[1][3][0] = -ga1 * v3;                                    J.jac[1][3][2] = v1;
    J.jac[1][4][0] = ga1;

    J.jac[2][0][0] = -v2 * v1; J.jac[2][0][1] = -v2o2 + ga1h * vs; J.jac[2][0][2] = -v2 * v3;
    J.jac[2][1][0] = v2;       J.jac[2][1][1] = -ga1 * v1;
    J.jac[2][2][0] = v1;       J.jac[2][2][1] = -ga3 * v2;         J.jac[2][2][2] = v3;
                               J.jac[2][3][1] = -ga1 * v3;         J.jac[2][3][2] = v2;
                               J.jac[2][4][1] = ga1;

    J.jac[3][0][0] = -v3 * v1; J.jac[3][0][1] = -v3 * v2; J.jac[3][0][2] = -v3o2 + ga1h * vs;
    J.jac[3][1][0] = v3;                                  J.jac[3][1][2] = -ga1 * v1;
                               J.jac[3][2][1] = v3;       J.jac[3][2][2] = -ga1 * v2;
    J.jac[3][3][0] = v1;       J.jac[3][3][1] = v2;       J.jac[3][3][2] = -ga3 * v3;
                                                          J.jac[3][4][2] = ga1;

    J.jac[4][0][0] = by_rho2((-gretot + g1ke2) * u2);
    J.jac[4][0][1] = by_rho2((-gretot + g1ke2) * u3);
    J.jac[4][0][2] = by_rho2((-gretot + g1ke2) * u4);
    J.jac[4][1][0] = getot - ga1h * (vs + 2 * v1o2); J.jac[4][1][1] = -ga1 * v1 * v2; J.jac[4][1][2] = -ga1 * v1 * v3;
    J.jac[4][2][0] = -ga1 * v2 * v1; J.jac[4][2][1] = getot - ga1h * (vs + 2 * v2o2); J.jac[4][2][2] = -ga1 * v2 * v3;
    J.jac[4][3][0] = -ga1 * v3 * v1; J.jac[4][3][1] = -ga1 * v3 * v2; J.jac[4][3][2] = getot - ga1h * (vs + 2 * v3o2);
    J.jac[4][4][0] = ga * v1; J.jac[4][4][1] = ga * v2; J.jac[4][4][2] = ga * v3;
    return by_u1.ok && by_rho2.ok;
}

template <> __device__ __forceinline__ bool jaco_update_impl<2, true>(Jaco<2> & J, double ga, double const (&u)[4]) { return jaco_update_2d<true>(J, ga, u); }
template <> __device__ __forceinline__ bool jaco_update_impl<2, false>(Jaco<2> & J, double ga, double const (&u)[4]) { return jaco_update_2d<false>(J, ga, u); }
template <> __device__ __forceinline__ bool jaco_update_impl<3, true>(Jaco<3> & J, double ga, double const (&u)[5]) { return jaco_update_3d<true>(J, ga, u); }
template <> __device__ __forceinline__ bool jaco_update_impl<3, false>(Jaco<3> & J, double ga, double const (&u)[5]) { return jaco_update_3d<false>(J, ga, u); }

// acc += jac[i][j][d] * x with the structural zeros skipped and the ones not multiplied
#define SCG_JTERM(ND, J, i, j, d, x, acc)                   \
    do                                                      \
    {                                                       \
        if (2 == jkind<ND>(i, j, d)) { acc += J.jac[i][j][d] * (x); } \
        else if (1 == jkind<ND>(i, j, d)) { acc += (x); }   \
    } while (0)

// calc_solt_impl body, euler_march.cpp:175-189: so0t[i] = -sum_d sum_j jac[i][j][d] * g[j][d]
template <int ND>
__device__ __forceinline__ void solt_row(Jaco<ND> const & J, double const (&g)[ND + 2][ND], double (&t)[ND + 2])
{
#pragma unroll
    for (int ieq = 0; ieq < ND + 2; ++ieq)
    {
        double val = 0.0;
#pragma unroll
        for (int idm = 0; idm < ND; ++idm)
        {
#pragma unroll
            for (int jeq = 0; jeq < ND + 2; ++jeq) { SCG_JTERM(ND, J, ieq, jeq, idm, g[jeq][idm], val); }
        }
        t[ieq] = -val;
    }
}

// ------------------------------------------------------------------------------------------------
// GradientElementTypeGroup tables, GradientElement.cpp:16-79 (0-based local faces here)
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fge_face(int type, int isub, int ivx)
{
    constexpr int quad[4][3] = {{0, 1, -1}, {1, 2, -1}, {2, 3, -1}, {3, 0, -1}};
    constexpr int tri[3][3] = {{0, 1, -1}, {1, 2, -1}, {2, 0, -1}};
    constexpr int hex[8][3] = {{1, 2, 4}, {5, 2, 1}, {3, 2, 5}, {4, 2, 3}, {4, 0, 1}, {1, 0, 5}, {5, 0, 3}, {3, 0, 4}};
    constexpr int tet[4][3] = {{2, 0, 1}, {1, 0, 3}, {3, 0, 2}, {1, 3, 2}};
    constexpr int prism[6][3] = {{4, 1, 3}, {2, 1, 4}, {3, 1, 2}, {3, 0, 4}, {4, 0, 2}, {2, 0, 3}};
    constexpr int pyramid[6][3] = {{0, 4, 1}, {1, 4, 2}, {2, 4, 3}, {3, 4, 0}, {0, 2, 3}, {2, 0, 1}};
    return type == CT_QUAD    ? quad[isub][ivx]
         : type == CT_TRI     ? tri[isub][ivx]
         : type == CT_HEX     ? hex[isub][ivx]
         : type == CT_TET     ? tet[isub][ivx]
         : type == CT_PRISM   ? prism[isub][ivx]
                              : pyramid[isub][ivx];
}

// GradientElement::determinant / adjugate, GradientElement.cpp:215-249 (rows of `a` are vertices)
template <int ND>
__device__ __forceinline__ double det_n(double const (&a)[ND][ND])
{
    if (2 == ND) { return a[0][0] * a[1][1] - a[0][1] * a[1][0]; }
    return a[0][0] * (a[1][1] * a[ND - 1][ND - 1] - a[1][ND - 1] * a[ND - 1][1])
         - a[0][1] * (a[1][0] * a[ND - 1][ND - 1] - a[1][ND - 1] * a[ND - 1][0])
         + a[0][ND - 1] * (a[1][0] * a[ND - 1][1] - a[1][1] * a[ND - 1][0]);
}

template <int ND>
__device__ __forceinline__ void adj_n(double const (&a)[ND][ND], double (&r)[ND][ND])
{
    if (2 == ND)
    {
        r[0][0] = a[1][1];
        r[0][1] = -a[0][1];
        r[1][0] = -a[1][0];
        r[1][1] = a[0][0];
        return;
    }
    constexpr int Z = ND - 1; // == 2; written so the 2D instantiation stays in bounds
    r[0][0] = a[1][1] * a[Z][Z] - a[1][Z] * a[Z][1];
    r[0][1] = a[0][Z] * a[Z][1] - a[0][1] * a[Z][Z];
    r[0][Z] = a[0][1] * a[1][Z] - a[0][Z] * a[1][1];
    r[1][0] = a[1][Z] * a[Z][0] - a[1][0] * a[Z][Z];
    r[1][1] = a[0][0] * a[Z][Z] - a[0][Z] * a[Z][0];
    r[1][Z] = a[0][Z] * a[1][0] - a[0][0] * a[1][Z];
    r[Z][0] = a[1][0] * a[Z][1] - a[1][1] * a[Z][0];
    r[Z][1] = a[0][1] * a[Z][0] - a[0][0] * a[Z][1];
    r[Z][Z] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
}

// ------------------------------------------------------------------------------------------------
// per-cell phases (device functions so the per-phase kernels and the fused substep share them)
// ------------------------------------------------------------------------------------------------

// calc_soln_impl for one cell, euler_march.cpp:197-303.  Returns so0n in `out`.
template <int TYPE>
__device__ __forceinline__ void cell_soln(MdField const & F, MdGeo const & G, int32_t i, double (&out)[ct_ndim(TYPE) + 2])
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    int64_t const n = G.n;
    int64_t const S = F.S;
    double const qdt = F.dt * 0.25;
    double const hdt = F.dt * 0.5;
    double acc[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { acc[e] = 0.0; }

    int64_t jnext = G.nbr[i];
#pragma unroll 1
    for (int f = 0; f < NFC; ++f)
    {
        int64_t const j = jnext;
#if SCG_SOLN_PREFETCH
        // The loop body is long (a Jacobian and up to four sub-face flux evaluations) and only a few
        // warps fit on an SM, so the next face's neighbour gather and geometry are pulled towards the
        // SM now and hit the cache when the next iteration asks for them.
        if (f + 1 < NFC)
        {
            jnext = G.nbr[(f + 1) * n + i];
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                prefetch_l1(&F.so0c[e * S + jnext]);
                prefetch_l1(&F.so0t[e * S + jnext]);
#pragma unroll
                for (int d = 0; d < ND; ++d) { prefetch_l1(&F.so1c[(e * ND + d) * S + jnext]); }
            }
            prefetch_l1(&F.gamma[jnext]);
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                prefetch_l1(&F.solpt[d * S + jnext]);
                prefetch_l1(&G.cecnd[((f + 1) * ND + d) * n + i]);
            }
            prefetch_l1(&G.cevol[(f + 2) * n + i]);
        }
#else
        if (f + 1 < NFC) { jnext = G.nbr[(f + 1) * n + i]; }
#endif
        double jc[ND], bc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            jc[d] = F.solpt[d * S + j];
            bc[d] = G.cecnd[(f * ND + d) * n + i];
        }
        double const bvol = G.cevol[(f + 1) * n + i];
        double ju[NEQ], jt[NEQ], jg[NEQ][ND];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            ju[e] = F.so0c[e * S + j];
            jt[e] = F.so0t[e * S + j];
#pragma unroll
            for (int d = 0; d < ND; ++d) { jg[e][d] = F.so1c[(e * ND + d) * S + j]; }
        }
        double const ga = F.gamma[j];
        // spatial part over the basic CE (euler_march.cpp:236-249)
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double fusp = ju[e];
#pragma unroll
            for (int d = 0; d < ND; ++d) { fusp += (bc[d] - jc[d]) * jg[e][d]; }
            acc[e] += fusp * bvol;
        }
        Jaco<ND> J;
        jaco_update<ND>(J, ga, ju);
        // temporal part over the sub-faces (euler_march.cpp:262-294)
        int const nsf = ct_fnnd(TYPE, f);
        int const q0 = ct_sfoff(TYPE, f);
#pragma unroll 1
        for (int s = 0; s < nsf; ++s)
        {
            int64_t const q = q0 + s;
            double sc[ND], sn[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                sc[d] = G.sfcnd[(q * ND + d) * n + i];
                sn[d] = G.sfnml[(q * ND + d) * n + i];
            }
            double usfc[NEQ];
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                usfc[e] = qdt * jt[e];
#pragma unroll
                for (int d = 0; d < ND; ++d) { usfc[e] += (sc[d] - jc[d]) * jg[e][d]; }
            }
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                double dot = 0.0;
#pragma unroll
                for (int d2 = 0; d2 < ND; ++d2)
                {
                    double dfcn = J.fcn[e][d2];
#pragma unroll
                    for (int je = 0; je < NEQ; ++je) { SCG_JTERM(ND, J, e, je, d2, usfc[je], dfcn); }
                    dot += dfcn * sn[d2];
                }
                acc[e] -= hdt * dot;
            }
        }
    }
    double const cvol = G.cevol[i];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = acc[e] / cvol; }
}

// calc_soln_impl for one cell, second formulation: identical terms and identical summation order per
// equation, but organised for the memory system.  Everything a face needs -- the neighbour gather AND the
// geometry of all of its sub-faces -- is requested in one batch at the top of the face iteration (one
// exposed memory latency per face instead of one per sub-face), and the flux evaluation runs equation-outer /
// sub-face-inner so only one Jacobian row is live at a time.  acc[e] still receives the spatial term and
// then the sub-face terms in sub-face order, so every rounding is the one of euler_march.cpp:236-294.
template <int TYPE>
__device__ __forceinline__ void cell_soln_rows(MdField const & F, MdGeo const & G, int32_t i, double (&out)[ct_ndim(TYPE) + 2])
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    constexpr int MSF = (ND == 2) ? 2 : ((TYPE == CT_TET) ? 3 : 4); // most sub-faces on one face
    int64_t const n = G.n;
    int64_t const S = F.S;
    double const qdt = F.dt * 0.25;
    double const hdt = F.dt * 0.5;
    double acc[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { acc[e] = 0.0; }

    int64_t jnext = G.nbr[i];
#pragma unroll 1
    for (int f = 0; f < NFC; ++f)
    {
        int64_t const j = jnext;
        if (f + 1 < NFC) { jnext = G.nbr[(f + 1) * n + i]; }
        int const nsf = ct_fnnd(TYPE, f);
        int const q0 = ct_sfoff(TYPE, f);
        // ---- one batch of loads
        double jc[ND], bc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            jc[d] = F.solpt[d * S + j];
            bc[d] = G.cecnd[(f * ND + d) * n + i];
        }
        double const bvol = G.cevol[(f + 1) * n + i];
        double ju[NEQ], jt[NEQ], jg[NEQ][ND];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            ju[e] = F.so0c[e * S + j];
            jt[e] = F.so0t[e * S + j];
#pragma unroll
            for (int d = 0; d < ND; ++d) { jg[e][d] = F.so1c[(e * ND + d) * S + j]; }
        }
        double const ga = F.gamma[j];
        double sc[MSF][ND], sn[MSF][ND];
#pragma unroll
        for (int s = 0; s < MSF; ++s)
        {
            int64_t const q = q0 + (s < nsf ? s : nsf - 1); // a triangle face of a prism/pyramid re-reads its last sub-face
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                sc[s][d] = G.sfcnd[(q * ND + d) * n + i];
                sn[s][d] = G.sfnml[(q * ND + d) * n + i];
            }
        }
        // ---- spatial part over the basic CE (euler_march.cpp:236-249)
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double fusp = ju[e];
#pragma unroll
            for (int d = 0; d < ND; ++d) { fusp += (bc[d] - jc[d]) * jg[e][d]; }
            acc[e] += fusp * bvol;
        }
        // ---- Taylor state on every sub-face (euler_march.cpp:268-280)
        double usfc[MSF][NEQ];
#pragma unroll
        for (int s = 0; s < MSF; ++s)
        {
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                usfc[s][e] = qdt * jt[e];
#pragma unroll
                for (int d = 0; d < ND; ++d) { usfc[s][e] += (sc[s][d] - jc[d]) * jg[e][d]; }
            }
        }
        Jaco<ND> J;
        jaco_update<ND>(J, ga, ju);
        // ---- temporal part, one equation (Jacobian row) at a time (euler_march.cpp:281-294)
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
#pragma unroll
            for (int s = 0; s < MSF; ++s)
            {
                double dot = 0.0;
#pragma unroll
                for (int d2 = 0; d2 < ND; ++d2)
                {
                    double dfcn = J.fcn[e][d2];
#pragma unroll
                    for (int je = 0; je < NEQ; ++je) { SCG_JTERM(ND, J, e, je, d2, usfc[s][je], dfcn); }
                    dot += dfcn * sn[s][d2];
                }
                if (s < nsf) { acc[e] -= hdt * dot; }
            }
        }
    }
    double const cvol = G.cevol[i];
#if SCG_SHARED_RCP
    DivBy<true> by_cvol(cvol);
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = by_cvol(acc[e]); }
    if (by_cvol.ok) { return; }
#endif
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = acc[e] / cvol; }
}

// EulerCore::calc_cfl for one cell, euler_solution.cpp:106-160.  `u` is so0n of the cell; its energy
// is rewritten in place; returns cflo / cflc.
template <int TYPE>
__device__ __forceinline__ void cell_cfl(MdField const & F, MdGeo const & G, int32_t i, double (&u)[ct_ndim(TYPE) + 2],
                                         double & cflo, double & cflc)
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    int64_t const n = G.n;
    int64_t const k = G.beg + i;
    double const hdt = F.dt / 2.0;
    double cc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { cc[d] = F.solpt[d * F.S + k]; }
    double dist = DBL_MAX;
#pragma unroll
    for (int f = 0; f < NFC; ++f)
    {
        double d2 = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            double const diff = G.cecnd[(f * ND + d) * n + i] - cc[d];
            d2 += diff * diff;
        }
        double const r = sqrt(d2);
        dist = (r < dist) ? r : dist; // std::min(dist, r)
    }
    double const ga = F.gamma[k];
    double const ga1 = ga - 1.0;
    double const density = u[0];
    double momsq = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) { momsq += u[1 + d] * u[1 + d]; }
    double const ke = momsq / (2.0 * density);
    double const energy = u[NEQ - 1];
    double const pr = ga1 * (energy - ke);
    double const pr_adj = (pr + fabs(pr)) / 2.0;
    double const wspd = sqrt(ga * pr_adj / density) + sqrt(momsq) / density;
    cflo = hdt * wspd / dist;
    cflc = (cflo - 1.0) * pr_adj / (pr_adj + MD_TINY) + 1.0;
    u[NEQ - 1] = pr_adj / ga1 + ke + MD_TINY;
}

// The thermodynamic half of calc_cfl (euler_solution.cpp:133-158) once `dist` is known.  FAST: the six quotients and
// two square roots run their inlined IEEE sequences with one deferred range flag (scg_fastmath.cuh); returns it.
template <int ND, bool FAST>
__device__ __forceinline__ bool cfl_from_dist_impl(double hdt, double ga, double dist, double (&u)[ND + 2], double & cflo, double & cflc)
{
    constexpr int NEQ = ND + 2;
    bool ok = true;
    double const ga1 = ga - 1.0;
    double const density = u[0];
    double momsq = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) { momsq += u[1 + d] * u[1 + d]; }
    DivBy<FAST> by_2rho(2.0 * density);
    DivBy<FAST> by_rho(density);
    DivBy<FAST> by_dist(dist);
    DivBy<FAST> by_ga1(ga1);
    double const ke = by_2rho(momsq);
    double const energy = u[NEQ - 1];
    double const pr = ga1 * (energy - ke);
    double const pr_adj = (pr + fabs(pr)) / 2.0;
    DivBy<FAST> by_pr(pr_adj + MD_TINY);
    double const c2 = by_rho(ga * pr_adj);
    double const s1 = FAST ? sqrt_fast_z(c2, ok) : sqrt(c2);
    double const s2 = FAST ? sqrt_fast_z(momsq, ok) : sqrt(momsq);
    double const wspd = s1 + by_rho(s2);
    cflo = by_dist(hdt * wspd);
    cflc = by_pr((cflo - 1.0) * pr_adj) + 1.0;
    double const e_new = by_ga1(pr_adj) + ke + MD_TINY;
    ok = ok && by_2rho.ok && by_rho.ok && by_dist.ok && by_ga1.ok && by_pr.ok;
    if (ok || !FAST) { u[NEQ - 1] = e_new; }
    return ok;
}

template <int ND>
__device__ __forceinline__ void cfl_from_dist(double hdt, double ga, double dist, double (&u)[ND + 2], double & cflo, double & cflc)
{
#if SCG_SHARED_RCP
    if (cfl_from_dist_impl<ND, true>(hdt, ga, dist, u, cflo, cflc)) { return; }
#endif
    cfl_from_dist_impl<ND, false>(hdt, ga, dist, u, cflo, cflc);
}

// ------------------------------------------------------------------------------------------------
// Asynchronous global -> shared copies (LDGSTS).  Every thread copies the elements it will read itself,
// so cp.async.wait_group is the only synchronisation needed -- no CTA barrier.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double * smem_dst, double const * gsrc)
{
    unsigned const d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// 16 bytes, L2 only (.cg): for rows a warp reads once -- they do not displace the neighbour gathers from L1
__device__ __forceinline__ void cp_async16(double * smem_dst, double const * gsrc)
{
    unsigned const d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// 16 bytes through L1 (.ca): a gathered pair
__device__ __forceinline__ void cp_async16_ca(double * smem_dst, void const * gsrc)
{
    unsigned const d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// Ask L2 for the line holding *p: no register, no scoreboard.  Used for data THIS thread reads a few microseconds
// later (later faces / later equations of its own cell), so that the staged copies and gathers meet L2 instead of HBM.
__device__ __forceinline__ void l2_prefetch(void const * p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#ifndef SCG_SOLN_L2PF
#define SCG_SOLN_L2PF 0 // 1: geometry rows of faces >= 2 at the top of the cell; 2: + the neighbour rows of faces >= 2
#endif
#ifndef SCG_EQO_GSMEM_MASK
#define SCG_EQO_GSMEM_MASK 0 // 3D cell types whose new gradient is parked in shared memory for the fused calc_solt: 1 hex, 2 tet, 4 prism
#endif
#ifndef SCG_EQO_L2PF
#define SCG_EQO_L2PF 0 // 1: the neighbour rows of equations >= 1 (>= 2 with register prefetch) right after the geometry
#endif
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Bulk asynchronous copies (cp.async.bulk, the TMA's 1D mode) completing on an mbarrier.  The geometry rows of a
// cell group are contiguous over its cells, so the 32 doubles a warp needs of one row are ONE 256-byte copy issued by
// one lane instead of 32 8-byte LDGSTS: no per-lane address arithmetic, nothing through the LSU / L1 (which the
// neighbour gathers keep for themselves).  One mbarrier per warp and buffer; a wait flips with the buffer's use count.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t * bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(double * smem_dst, double const * gsrc, unsigned bytes, uint64_t * bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gsrc), "r"(bytes), "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar)))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "SCG_MBAR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra SCG_MBAR_DONE;\n"
                 "bra SCG_MBAR_WAIT;\n"
                 "SCG_MBAR_DONE:\n"
                 "}" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))),
                 "r"(parity)
                 : "memory");
}

// How md_solncfl_staged stages the geometry rows, per cell type (one decimal digit each, in the order quad, tri, hex,
// tet, prism, pyramid from the lowest digit): 0 = 8-byte cp.async.ca per thread, 1 = cp.async.bulk + mbarrier (one lane
// per warp, 256-byte copies), 2 = 16-byte cp.async.cg (two rows per warp instruction, nothing allocated in L1).
// Measured (profiles/r02f_*): mode 2 is 1-3 % faster for hexahedra, prisms, quadrilaterals and triangles and 15 % slower
// for tetrahedra (odd row counts: 13 and 9); mode 1 is slower everywhere.
#ifndef SCG_SOLN_STAGE_MODES
#define SCG_SOLN_STAGE_MODES 220222
#endif
__host__ __device__ constexpr int soln_stage_mode(int type)
{
    int m = SCG_SOLN_STAGE_MODES;
    for (int t = CT_QUAD; t < type; ++t) { m /= 10; }
    return m % 10;
}

// Jacobian in scalar form: the quantities every entry of EulerJacobian<ND>::update is built from
// (euler_march.cpp:35-81, 94-157), and row<E>() which forms fcn[E][:] and jac[E][:][:] from them with the
// very expressions of jaco_update_2d/3d.  Evaluating a row right where it is consumed keeps one row in
// registers instead of the whole tensor (63 doubles in 3D).
template <int ND>
struct JacoS
{
    static constexpr int NEQ = ND + 2;
    double ga, ga1, ga3, ga1h;
    double m[ND];   // momenta u2..
    double en;      // total energy
    double v[ND];   // velocities
    double vo2[ND]; // squared velocities
    double vs, getot, pr;
    double e0[ND];  // jac[NEQ-1][0][:]

    template <bool FAST>
    __device__ __forceinline__ bool set_impl(double ga_in, double const (&u)[NEQ])
    {
        ga = ga_in;
        ga1 = ga - 1;
        ga3 = ga - 3;
        ga1h = ga1 / 2;
        double const u1 = u[0] + MD_TINY;
        double const rho2 = u1 * u1;
        DivBy<FAST> by_u1(u1);
        DivBy<FAST> by_rho2(rho2);
        double msq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            m[d] = u[1 + d];
            v[d] = by_u1(m[d]);
            vo2[d] = v[d] * v[d];
            msq = (0 == d) ? m[d] * m[d] : msq + m[d] * m[d];
        }
        en = u[NEQ - 1];
        double const ke2 = by_u1(msq);
        double const g1ke2 = ga1 * ke2;
        vs = by_u1(ke2);
        double const gretot = ga * en;
        getot = by_u1(gretot);
        pr = ga1 * en - ga1h * ke2;
#pragma unroll
        for (int d = 0; d < ND; ++d) { e0[d] = by_rho2((-gretot + g1ke2) * m[d]); }
        return by_u1.ok && by_rho2.ok;
    }
    __device__ __forceinline__ void set(double ga_in, double const (&u)[NEQ])
    {
#if SCG_SHARED_RCP
        if (set_impl<true>(ga_in, u)) { return; }
#endif
        set_impl<false>(ga_in, u);
    }
    // The members of set() that cost divisions -- v[], vs, getot, pr, e0[] -- as NJS values per row.  A cell's state is
    // the neighbour state of each of its faces' other cell, so calc_soln evaluates set() of the same state once per
    // face that sees it (6 times for a hexahedron's); pack() is evaluated once, by the kernel that produces so0t of the
    // row, and load() rebuilds the object from the stored values: the same numbers, since set() is deterministic.
    static constexpr int NJS = 2 * ND + 3;
    __device__ __forceinline__ void pack(double * js, int64_t S, int64_t k) const
    {
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            js[d * S + k] = v[d];
            js[(ND + 3 + d) * S + k] = e0[d];
        }
        js[ND * S + k] = vs;
        js[(ND + 1) * S + k] = getot;
        js[(ND + 2) * S + k] = pr;
    }
    __device__ __forceinline__ void load(double ga_in, double const (&u)[NEQ], double const (&q)[NJS])
    {
        ga = ga_in;
        ga1 = ga - 1;
        ga3 = ga - 3;
        ga1h = ga1 / 2;
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            m[d] = u[1 + d];
            v[d] = q[d];
            vo2[d] = v[d] * v[d];
            e0[d] = q[ND + 3 + d];
        }
        en = u[NEQ - 1];
        vs = q[ND];
        getot = q[ND + 1];
        pr = q[ND + 2];
    }
    // fcn[E][d] and the general entries of jac[E][j][d] (structural zeros / ones are never read)
    template <int E>
    __device__ __forceinline__ void row(double (&fc)[ND], double (&jr)[NEQ][ND]) const
    {
        if (0 == E)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { fc[d] = m[d]; }
        }
        else if (NEQ - 1 == E)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                fc[d] = (pr + en) * v[d];
                jr[0][d] = e0[d];
                jr[NEQ - 1][d] = ga * v[d];
#pragma unroll
                for (int j = 1; j <= ND; ++j)
                {
                    jr[j][d] = (j - 1 == d) ? getot - ga1h * (vs + 2 * vo2[d]) : -ga1 * v[j - 1] * v[d];
                }
            }
        }
        else
        {
            constexpr int a = (E >= 1 && E <= ND) ? E - 1 : 0; // this row's momentum component
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                fc[d] = (d == a) ? pr + m[a] * v[d] : m[a] * v[d];
                jr[0][d] = (d == a) ? -vo2[a] + ga1h * vs : -v[a] * v[d];
#pragma unroll
                for (int j = 1; j <= ND; ++j)
                {
                    if (j - 1 == a) { jr[j][d] = (d == a) ? -ga3 * v[a] : v[d]; } // diagonal block
                    else if (d == a) { jr[j][d] = -ga1 * v[j - 1]; }
                    else if (d == j - 1) { jr[j][d] = v[a]; }
                }
                if (d == a) { jr[NEQ - 1][d] = ga1; }
            }
        }
    }
};

// acc_e -= hdt * sum_d (fcn[E][d] + sum_j jac[E][j][d] * usfc[s][j]) * sn[s][d] over the sub-faces of one
// face, in sub-face order (euler_march.cpp:281-294)
template <int ND, int E, int MSF>
__device__ __forceinline__ void flux_row(JacoS<ND> const & JS, double const (&usfc)[MSF][ND + 2], double const (&sn)[MSF][ND],
                                         double hdt, int nsf, double & acc_e)
{
    constexpr int NEQ = ND + 2;
    double fc[ND], jr[NEQ][ND];
    JS.template row<E>(fc, jr);
#pragma unroll
    for (int s = 0; s < MSF; ++s)
    {
        // dot = 0.0 + t0 + t1 + ..: without the leading "0.0 +" (SCG_SKIP_ZERO_ADD) dot can be -0 where the reference has
        // +0; hdt * dot is then -0 instead of +0 and acc - (-0) == acc - (+0) for every acc but -0, which acc never is (it
        // starts from +0 and x - y is -0 only for x == -0)
        double dot = 0.0;
#pragma unroll
        for (int d2 = 0; d2 < ND; ++d2)
        {
            double dfcn = fc[d2];
#pragma unroll
            for (int je = 0; je < NEQ; ++je)
            {
                if (2 == jkind<ND>(E, je, d2)) { dfcn += jr[je][d2] * usfc[s][je]; }
                else if (1 == jkind<ND>(E, je, d2)) { dfcn += usfc[s][je]; }
            }
            double const t = dfcn * sn[s][d2];
            dot = (SCG_SKIP_ZERO_ADD && 0 == d2) ? t : dot + t;
        }
        if (s < nsf) { acc_e -= hdt * dot; }
    }
}

// calc_soln + the distance part of calc_cfl for one cell, third formulation ("staged"): the same terms
// in the same order as cell_soln_rows, but no load of the face being evaluated is ever waited for.
// Everything face f+1 needs travels global -> shared by cp.async while face f is being evaluated, into
// three single buffers that are refilled as soon as their last reader of face f is done:
//   N  the neighbour gather (solpt, so0c, so0t, so1c, gamma of the cell across the face)      29 rows in 3D
//   A  cecnd, cevol of the face and sfcnd of its sub-faces                                    16 rows (hex)
//   B  sfnml of its sub-faces                                                                 12 rows (hex)
// N and A are consumed by the spatial term, the Taylor states and the Jacobian; they are refilled for
// face f+1 before the flux evaluation of face f starts.  B is consumed by the flux evaluation and refilled
// right after it.  `stage` points at this thread's column of the CTA buffer [ROWS][TB]; every thread reads
// only what it copied itself, so cp.async.wait_group is the only synchronisation.
template <int TYPE, int TB>
struct SolnStage
{
    static constexpr int ND = ct_ndim(TYPE);
    static constexpr int NEQ = ND + 2;
    static constexpr int MSF = (ND == 2) ? 2 : ((TYPE == CT_TET) ? 3 : 4);
    // N
    static constexpr int R_JC = 0;                   // solpt[ND]
    static constexpr int R_JU = R_JC + ND;           // so0c[NEQ]
    static constexpr int R_JT = R_JU + NEQ;          // so0t[NEQ]
    static constexpr int R_JG = R_JT + NEQ;          // so1c[NEQ][ND]
    static constexpr int R_GA = R_JG + NEQ * ND;     // gamma
    static constexpr int NJS = SCG_JS_SHARE ? 2 * ND + 3 : 0;
    static constexpr int R_JS = R_GA + 1;            // js[NJS] (SCG_JS_SHARE)
    // A
    static constexpr int R_BC = R_JS + NJS;          // cecnd[f][ND]
    static constexpr int R_VOL = R_BC + ND;          // cevol[f+1]
    static constexpr int R_SC = R_VOL + 1;           // sfcnd[s][ND]
    // B
    static constexpr int R_SN = R_SC + MSF * ND;     // sfnml[s][ND]
    // E (SCG_SOLN_EARLY): what only the epilogue reads, copied once with the first group
    static constexpr int R_CV = R_SN + MSF * ND;     // cevol[0] of the cell
    static constexpr int R_GK = R_CV + 1;            // gamma of the cell
    static constexpr int ROWS = R_SN + MSF * ND + (soln_early(TYPE) ? 2 : 0);
    static constexpr int BYTES = ROWS * TB * int(sizeof(double));
};

// the epilogue's two values; joins the group the caller commits next
template <int TYPE, int TB>
__device__ __forceinline__ void soln_stage_issue_e(MdField const & F, MdGeo const & G, int64_t i, double * st)
{
    if constexpr (soln_early(TYPE))
    {
        using L = SolnStage<TYPE, TB>;
        cp_async8(st + L::R_CV * TB, &G.cevol[i]);
        cp_async8(st + L::R_GK * TB, &F.gamma[G.beg + i]);
    }
}

template <int TYPE, int TB>
__device__ __forceinline__ void soln_stage_issue_an(MdField const & F, MdGeo const & G, int64_t i, int64_t j, int f, double * st)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = L::ND, NEQ = L::NEQ;
    int64_t const n = G.n;
    int64_t const S = F.S;
#pragma unroll
    for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_JC + d) * TB, &F.solpt[d * S + j]); }
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
#if SCG_CT_PAIR
        cp_async16_ca(st + (L::R_JU + 2 * e) * TB + threadIdx.x, &F.ct[e * S + j]); // this thread's pair e of the 2*NEQ rows R_JU..
#else
        cp_async8(st + (L::R_JU + e) * TB, &F.so0c[e * S + j]);
        cp_async8(st + (L::R_JT + e) * TB, &F.so0t[e * S + j]);
#endif
#pragma unroll
        for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_JG + e * ND + d) * TB, &F.so1c[(e * ND + d) * S + j]); }
    }
    cp_async8(st + L::R_GA * TB, &F.gamma[j]);
#pragma unroll
    for (int q = 0; q < L::NJS; ++q) { cp_async8(st + (L::R_JS + q) * TB, &F.js[q * S + j]); }
    int const nsf = ct_fnnd(TYPE, f);
    int const q0 = ct_sfoff(TYPE, f);
#pragma unroll
    for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_BC + d) * TB, &G.cecnd[(f * ND + d) * n + i]); }
    cp_async8(st + L::R_VOL * TB, &G.cevol[(f + 1) * n + i]);
#pragma unroll
    for (int s = 0; s < L::MSF; ++s)
    {
        int64_t const q = q0 + (s < nsf ? s : nsf - 1); // a triangle face of a prism/pyramid re-reads its last sub-face
#pragma unroll
        for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_SC + s * ND + d) * TB, &G.sfcnd[(q * ND + d) * n + i]); }
    }
    cp_async_commit();
}

template <int TYPE, int TB>
__device__ __forceinline__ void soln_stage_issue_b(MdGeo const & G, int64_t i, int f, double * st)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = L::ND;
    int64_t const n = G.n;
    int const nsf = ct_fnnd(TYPE, f);
    int const q0 = ct_sfoff(TYPE, f);
#pragma unroll
    for (int s = 0; s < L::MSF; ++s)
    {
        int64_t const q = q0 + (s < nsf ? s : nsf - 1);
#pragma unroll
        for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_SN + s * ND + d) * TB, &G.sfnml[(q * ND + d) * n + i]); }
    }
    cp_async_commit();
}

template <int TYPE, int TB>
__device__ __forceinline__ void cell_soln_staged(MdField const & F, MdGeo const & G, int32_t i, double * st,
                                                 double (&out)[ct_ndim(TYPE) + 2], double & dist, bool & dist_ok, int32_t & jmax)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    constexpr int MSF = L::MSF;
    int64_t const n = G.n;
    double const qdt = F.dt * 0.25;
    double const hdt = F.dt * 0.5;
    double acc[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { acc[e] = 0.0; }

    int32_t jn = G.nbr[i];
    jmax = jn; // the largest neighbour index tells whether the cell touches a ghost
    soln_stage_issue_e<TYPE, TB>(F, G, i, st);
    soln_stage_issue_an<TYPE, TB>(F, G, i, jn, 0, st);
    soln_stage_issue_b<TYPE, TB>(G, i, 0, st);
    jn = G.nbr[n + i]; // NFC >= 3
    double cc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { cc[d] = F.solpt[d * F.S + G.beg + i]; }
    dist = DBL_MAX;
    dist_ok = true;
    double d2min = CUDART_INF;
#if SCG_SOLN_L2PF
    // the staged copies run ONE face ahead; faces >= 2 are requested from HBM now so that their copies meet L2
#pragma unroll
    for (int f = 2; f < NFC; ++f)
    {
        int const q0 = ct_sfoff(TYPE, f);
#pragma unroll
        for (int d = 0; d < ND; ++d) { l2_prefetch(&G.cecnd[(f * ND + d) * n + i]); }
        l2_prefetch(&G.cevol[(f + 1) * n + i]);
#pragma unroll
        for (int s = 0; s < ct_fnnd(TYPE, f); ++s)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                l2_prefetch(&G.sfcnd[((q0 + s) * ND + d) * n + i]);
                l2_prefetch(&G.sfnml[((q0 + s) * ND + d) * n + i]);
            }
        }
#if SCG_SOLN_L2PF >= 2
        {
            int64_t const j = G.nbr[f * n + i];
            int64_t const S = F.S;
#pragma unroll
            for (int d = 0; d < ND; ++d) { l2_prefetch(&F.solpt[d * S + j]); }
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                l2_prefetch(&F.so0c[e * S + j]);
                l2_prefetch(&F.so0t[e * S + j]);
#pragma unroll
                for (int d = 0; d < ND; ++d) { l2_prefetch(&F.so1c[(e * ND + d) * S + j]); }
            }
            l2_prefetch(&F.gamma[j]);
        }
#endif
    }
#endif

#pragma unroll 1
    for (int f = 0; f < NFC; ++f)
    {
        int const nsf = ct_fnnd(TYPE, f);
        double usfc[MSF][NEQ];
        JacoS<ND> JS;
        cp_async_wait<1>(); // N and A of this face have landed (B may still be in flight)
        {
            double jc[ND], ju[NEQ], jt[NEQ], jg[NEQ][ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { jc[d] = st[(L::R_JC + d) * TB]; }
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
#if SCG_CT_PAIR
                double2 const p = *reinterpret_cast<double2 const *>(st + (L::R_JU + 2 * e) * TB + threadIdx.x);
                ju[e] = p.x;
                jt[e] = p.y;
#else
                ju[e] = st[(L::R_JU + e) * TB];
                jt[e] = st[(L::R_JT + e) * TB];
#endif
#pragma unroll
                for (int d = 0; d < ND; ++d) { jg[e][d] = st[(L::R_JG + e * ND + d) * TB]; }
            }
            double const ga = st[L::R_GA * TB];
            // ---- spatial part over the basic CE (euler_march.cpp:236-249) and calc_cfl's distance
            {
                double bc[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { bc[d] = st[(L::R_BC + d) * TB]; }
                double const bvol = st[L::R_VOL * TB];
#pragma unroll
                for (int e = 0; e < NEQ; ++e)
                {
                    double fusp = ju[e];
#pragma unroll
                    for (int d = 0; d < ND; ++d) { fusp += (bc[d] - jc[d]) * jg[e][d]; }
                    acc[e] += fusp * bvol;
                }
                double d2 = 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    double const diff = bc[d] - cc[d];
                    d2 += diff * diff;
                }
                // euler_solution.cpp:129 takes std::min over the faces of sqrt(d2); the square root is monotone and
                // correctly rounded, so the minimum of the roots is the root of the minimum: one square root per cell
                d2min = (d2 < d2min) ? d2 : d2min;
            }
            // ---- Taylor state on every sub-face (euler_march.cpp:268-280)
#pragma unroll
            for (int s = 0; s < MSF; ++s)
            {
                double sc[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { sc[d] = st[(L::R_SC + s * ND + d) * TB] - jc[d]; }
#pragma unroll
                for (int e = 0; e < NEQ; ++e)
                {
                    usfc[s][e] = qdt * jt[e];
#pragma unroll
                    for (int d = 0; d < ND; ++d) { usfc[s][e] += sc[d] * jg[e][d]; }
                }
            }
#if SCG_JS_SHARE
            {
                double jq[L::NJS];
#pragma unroll
                for (int q = 0; q < L::NJS; ++q) { jq[q] = st[(L::R_JS + q) * TB]; }
                JS.load(ga, ju, jq);
            }
#else
            JS.set(ga, ju);
#endif
        }
        // ---- N and A are free: start the copies of the next face
        if (f + 1 < NFC)
        {
            soln_stage_issue_an<TYPE, TB>(F, G, i, jn, f + 1, st);
            jmax = jn > jmax ? jn : jmax;
            if (f + 2 < NFC) { jn = G.nbr[(f + 2) * n + i]; }
            cp_async_wait<1>(); // B of this face has landed
        }
        else { cp_async_wait<0>(); }
        double sn[MSF][ND];
#pragma unroll
        for (int s = 0; s < MSF; ++s)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { sn[s][d] = st[(L::R_SN + s * ND + d) * TB]; }
        }
        if (f + 1 < NFC) { soln_stage_issue_b<TYPE, TB>(G, i, f + 1, st); }
        // ---- temporal part, one equation (Jacobian row) at a time (euler_march.cpp:281-294)
        flux_row<ND, 0, MSF>(JS, usfc, sn, hdt, nsf, acc[0]);
        flux_row<ND, 1, MSF>(JS, usfc, sn, hdt, nsf, acc[1]);
        flux_row<ND, 2, MSF>(JS, usfc, sn, hdt, nsf, acc[2]);
        flux_row<ND, 3, MSF>(JS, usfc, sn, hdt, nsf, acc[3]);
        if (NEQ > 4) { flux_row<ND, (NEQ > 4 ? 4 : 0), MSF>(JS, usfc, sn, hdt, nsf, acc[NEQ - 1]); }
    }
    if (d2min < CUDART_INF) { dist = sqrt_fast(d2min, dist_ok); } // no finite face distance: dist stays DBL_MAX like in the reference
    double const cvol = soln_early(TYPE) ? st[L::R_CV * TB] : G.cevol[i];
#if SCG_SHARED_RCP
    DivBy<true> by_cvol(cvol);
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = by_cvol(acc[e]); }
    if (by_cvol.ok) { return; }
#endif
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = acc[e] / cvol; }
}


// The neighbour gather alone (the N rows of SolnStage): every thread copies its own column
template <int TYPE, int TB>
__device__ __forceinline__ void soln_stage_issue_n(MdField const & F, int64_t j, double * st)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = L::ND, NEQ = L::NEQ;
    int64_t const S = F.S;
#pragma unroll
    for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_JC + d) * TB, &F.solpt[d * S + j]); }
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
#if SCG_CT_PAIR
        cp_async16_ca(st + (L::R_JU + 2 * e) * TB + threadIdx.x, &F.ct[e * S + j]); // this thread's pair e of the 2*NEQ rows R_JU..
#else
        cp_async8(st + (L::R_JU + e) * TB, &F.so0c[e * S + j]);
        cp_async8(st + (L::R_JT + e) * TB, &F.so0t[e * S + j]);
#endif
#pragma unroll
        for (int d = 0; d < ND; ++d) { cp_async8(st + (L::R_JG + e * ND + d) * TB, &F.so1c[(e * ND + d) * S + j]); }
    }
    cp_async8(st + L::R_GA * TB, &F.gamma[j]);
#pragma unroll
    for (int q = 0; q < L::NJS; ++q) { cp_async8(st + (L::R_JS + q) * TB, &F.js[q * S + j]); }
    cp_async_commit();
}

// The geometry rows of face f for the 32 cells [iw0, iw0 + 32) of a warp: one 256-byte bulk copy per row, issued by
// one lane, completing on the warp's mbarrier.  Group strides are multiples of 32 and iw0 is one too, so every source
// segment is 256-byte aligned and inside its row.
template <int TYPE, int TB>
__device__ __forceinline__ void soln_bulk_issue_a(MdGeo const & G, int64_t iw0, int f, double * wst, uint64_t * bar)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = L::ND;
    constexpr unsigned ROWB = 32 * sizeof(double);
    int64_t const n = G.n;
    int const nsf = ct_fnnd(TYPE, f);
    int const q0 = ct_sfoff(TYPE, f);
    mbar_expect_tx(bar, (ND + 1 + L::MSF * ND) * ROWB);
#pragma unroll
    for (int d = 0; d < ND; ++d) { bulk_g2s(wst + (L::R_BC + d) * TB, &G.cecnd[(f * ND + d) * n + iw0], ROWB, bar); }
    bulk_g2s(wst + L::R_VOL * TB, &G.cevol[(f + 1) * n + iw0], ROWB, bar);
#pragma unroll
    for (int s = 0; s < L::MSF; ++s)
    {
        int64_t const q = q0 + (s < nsf ? s : nsf - 1); // a triangle face of a prism/pyramid re-reads its last sub-face
#pragma unroll
        for (int d = 0; d < ND; ++d) { bulk_g2s(wst + (L::R_SC + s * ND + d) * TB, &G.sfcnd[(q * ND + d) * n + iw0], ROWB, bar); }
    }
}

template <int TYPE, int TB>
__device__ __forceinline__ void soln_bulk_issue_b(MdGeo const & G, int64_t iw0, int f, double * wst, uint64_t * bar)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = L::ND;
    constexpr unsigned ROWB = 32 * sizeof(double);
    int64_t const n = G.n;
    int const nsf = ct_fnnd(TYPE, f);
    int const q0 = ct_sfoff(TYPE, f);
    mbar_expect_tx(bar, L::MSF * ND * ROWB);
#pragma unroll
    for (int s = 0; s < L::MSF; ++s)
    {
        int64_t const q = q0 + (s < nsf ? s : nsf - 1);
#pragma unroll
        for (int d = 0; d < ND; ++d) { bulk_g2s(wst + (L::R_SN + s * ND + d) * TB, &G.sfnml[(q * ND + d) * n + iw0], ROWB, bar); }
    }
}

// The same rows by 16-byte cp.async.cg (staging mode 2): a warp's 32 doubles of one row are 16 chunks, so one
// instruction of the warp moves TWO rows (lanes 0-15 the first, lanes 16-31 the second): half the LDGSTS of the 8-byte
// scheme, 128-bit accesses, and nothing allocated in L1.  Rows are addressed inside the group's geometry block
// [cevol | cecnd | sfcnd | sfnml], all with stride n.
template <int TYPE, int TB>
struct SolnRows
{
    using L = SolnStage<TYPE, TB>;
    static constexpr int ND = L::ND;
    static constexpr int NFC = ct_nfc(TYPE);
    static constexpr int NSF = ct_nsf(TYPE);
    // geometry-block row of staging row r of buffer A (r = 0 .. ND + MSF*ND) for face f
    __device__ static __forceinline__ int row_a(int r, int f, int q0, int nsf)
    {
        if (r < ND) { return (NFC + 1) + f * ND + r; }                // cecnd[f][d]
        if (r == ND) { return f + 1; }                                // cevol[f + 1]
        int const s = (r - ND - 1) / ND, d = (r - ND - 1) % ND;
        int const q = q0 + (s < nsf ? s : nsf - 1);
        return (NFC + 1) + NFC * ND + q * ND + d;                     // sfcnd[q][d]
    }
    __device__ static __forceinline__ int row_b(int r, int q0, int nsf)
    {
        int const s = r / ND, d = r % ND;
        int const q = q0 + (s < nsf ? s : nsf - 1);
        return (NFC + 1) + NFC * ND + NSF * ND + q * ND + d;          // sfnml[q][d]
    }
};

template <int TYPE, int TB>
__device__ __forceinline__ void soln_vec_issue_a(MdGeo const & G, int64_t iw0, int f, int lane, double * wst)
{
    using L = SolnStage<TYPE, TB>;
    using R = SolnRows<TYPE, TB>;
    constexpr int NR = L::ND + 1 + L::MSF * L::ND;
    int64_t const n = G.n;
    int const nsf = ct_fnnd(TYPE, f);
    int const q0 = ct_sfoff(TYPE, f);
    int const hi = lane >> 4, ch = (lane & 15) * 2;
#pragma unroll
    for (int p = 0; p < (NR + 1) / 2; ++p)
    {
        int const r = 2 * p + hi;
        if (r < NR) { cp_async16(wst + (L::R_BC + r) * TB + ch, G.cevol + R::row_a(r, f, q0, nsf) * n + iw0 + ch); }
    }
}

template <int TYPE, int TB>
__device__ __forceinline__ void soln_vec_issue_b(MdGeo const & G, int64_t iw0, int f, int lane, double * wst)
{
    using L = SolnStage<TYPE, TB>;
    using R = SolnRows<TYPE, TB>;
    constexpr int NR = L::MSF * L::ND;
    int64_t const n = G.n;
    int const nsf = ct_fnnd(TYPE, f);
    int const q0 = ct_sfoff(TYPE, f);
    int const hi = lane >> 4, ch = (lane & 15) * 2;
#pragma unroll
    for (int p = 0; p < (NR + 1) / 2; ++p)
    {
        int const r = 2 * p + hi;
        if (r < NR) { cp_async16(wst + (L::R_SN + r) * TB + ch, G.cevol + R::row_b(r, q0, nsf) * n + iw0 + ch); }
    }
    cp_async_commit();
}

// cell_soln_staged with the geometry rows staged by bulk copies (MODE 1) or 16-byte copies (MODE 2): same terms, same order.  `i` is the
// cell this lane evaluates (clamped into the launch range for the lanes that have none: they keep the warp's
// barriers and __syncwarp()s company and their results are dropped by the caller), `iw0` the first cell of the warp.
template <int TYPE, int TB, int MODE>
__device__ __forceinline__ void cell_soln_staged_bulk(MdField const & F, MdGeo const & G, int32_t i, int32_t iw0, int lane, double * st,
                                                      uint64_t * bars, double (&out)[ct_ndim(TYPE) + 2], double & dist, bool & dist_ok,
                                                      int32_t & jmax)
{
    using L = SolnStage<TYPE, TB>;
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    constexpr int MSF = L::MSF;
    int64_t const n = G.n;
    double const qdt = F.dt * 0.25;
    double const hdt = F.dt * 0.5;
    double acc[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { acc[e] = 0.0; }

    double * const wst = st - lane; // the warp's 32 columns of the staging buffer
    int32_t jn = G.nbr[i];
    jmax = jn; // the largest neighbour index tells whether the cell touches a ghost
    soln_stage_issue_e<TYPE, TB>(F, G, i, st);
    if constexpr (2 == MODE)
    {
        soln_vec_issue_a<TYPE, TB>(G, iw0, 0, lane, wst);
        soln_stage_issue_n<TYPE, TB>(F, jn, st); // commits the group: A and N of face 0
        soln_vec_issue_b<TYPE, TB>(G, iw0, 0, lane, wst);
    }
    else
    {
        soln_stage_issue_n<TYPE, TB>(F, jn, st);
        if (0 == lane)
        {
            soln_bulk_issue_a<TYPE, TB>(G, iw0, 0, wst, &bars[0]);
            soln_bulk_issue_b<TYPE, TB>(G, iw0, 0, wst, &bars[1]);
        }
    }
    jn = G.nbr[n + i]; // NFC >= 3
    double cc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { cc[d] = F.solpt[d * F.S + G.beg + i]; }
    dist = DBL_MAX;
    dist_ok = true;
    double d2min = CUDART_INF;

#pragma unroll 1
    for (int f = 0; f < NFC; ++f)
    {
        int const nsf = ct_fnnd(TYPE, f);
        double usfc[MSF][NEQ];
        JacoS<ND> JS;
        if constexpr (2 == MODE)
        {
            cp_async_wait<1>(); // this lane's share of A and N of this face has landed (B may still be in flight) ...
            __syncwarp();       // ... and so has every other lane's
        }
        else
        {
            cp_async_wait<0>();          // the neighbour gather of this face (this thread's own copies)
            mbar_wait(&bars[0], f & 1);  // cecnd / cevol / sfcnd rows of this face (the warp's bulk copies)
        }
        {
            double jc[ND], ju[NEQ], jt[NEQ], jg[NEQ][ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { jc[d] = st[(L::R_JC + d) * TB]; }
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
#if SCG_CT_PAIR
                double2 const p = *reinterpret_cast<double2 const *>(st + (L::R_JU + 2 * e) * TB + threadIdx.x);
                ju[e] = p.x;
                jt[e] = p.y;
#else
                ju[e] = st[(L::R_JU + e) * TB];
                jt[e] = st[(L::R_JT + e) * TB];
#endif
#pragma unroll
                for (int d = 0; d < ND; ++d) { jg[e][d] = st[(L::R_JG + e * ND + d) * TB]; }
            }
            double const ga = st[L::R_GA * TB];
            // ---- spatial part over the basic CE (euler_march.cpp:236-249) and calc_cfl's distance
            {
                double bc[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { bc[d] = st[(L::R_BC + d) * TB]; }
                double const bvol = st[L::R_VOL * TB];
#pragma unroll
                for (int e = 0; e < NEQ; ++e)
                {
                    double fusp = ju[e];
#pragma unroll
                    for (int d = 0; d < ND; ++d) { fusp += (bc[d] - jc[d]) * jg[e][d]; }
                    acc[e] += fusp * bvol;
                }
                double d2 = 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    double const diff = bc[d] - cc[d];
                    d2 += diff * diff;
                }
                // euler_solution.cpp:129 takes std::min over the faces of sqrt(d2); the square root is monotone and
                // correctly rounded, so the minimum of the roots is the root of the minimum: one square root per cell
                d2min = (d2 < d2min) ? d2 : d2min;
            }
            // ---- Taylor state on every sub-face (euler_march.cpp:268-280)
#pragma unroll
            for (int s = 0; s < MSF; ++s)
            {
                double sc[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { sc[d] = st[(L::R_SC + s * ND + d) * TB] - jc[d]; }
#pragma unroll
                for (int e = 0; e < NEQ; ++e)
                {
                    usfc[s][e] = qdt * jt[e];
#pragma unroll
                    for (int d = 0; d < ND; ++d) { usfc[s][e] += sc[d] * jg[e][d]; }
                }
            }
#if SCG_JS_SHARE
            {
                double jq[L::NJS];
#pragma unroll
                for (int q = 0; q < L::NJS; ++q) { jq[q] = st[(L::R_JS + q) * TB]; }
                JS.load(ga, ju, jq);
            }
#else
            JS.set(ga, ju);
#endif
        }
        // ---- N and A are free (every lane of the warp has read its column): start the copies of the next face
        if constexpr (2 == MODE)
        {
            if (f + 1 < NFC)
            {
                __syncwarp();
                soln_vec_issue_a<TYPE, TB>(G, iw0, f + 1, lane, wst);
                soln_stage_issue_n<TYPE, TB>(F, jn, st);
                jmax = jn > jmax ? jn : jmax;
                if (f + 2 < NFC) { jn = G.nbr[(f + 2) * n + i]; }
                cp_async_wait<1>(); // B of this face
            }
            else { cp_async_wait<0>(); }
            __syncwarp();
        }
        else
        {
            if (f + 1 < NFC)
            {
                soln_stage_issue_n<TYPE, TB>(F, jn, st);
                __syncwarp();
                if (0 == lane) { soln_bulk_issue_a<TYPE, TB>(G, iw0, f + 1, wst, &bars[0]); }
                jmax = jn > jmax ? jn : jmax;
                if (f + 2 < NFC) { jn = G.nbr[(f + 2) * n + i]; }
            }
            mbar_wait(&bars[1], f & 1); // sfnml rows of this face
        }
        double sn[MSF][ND];
#pragma unroll
        for (int s = 0; s < MSF; ++s)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { sn[s][d] = st[(L::R_SN + s * ND + d) * TB]; }
        }
        if (f + 1 < NFC)
        {
            __syncwarp();
            if constexpr (2 == MODE) { soln_vec_issue_b<TYPE, TB>(G, iw0, f + 1, lane, wst); }
            else if (0 == lane) { soln_bulk_issue_b<TYPE, TB>(G, iw0, f + 1, wst, &bars[1]); }
        }
        // ---- temporal part, one equation (Jacobian row) at a time (euler_march.cpp:281-294)
        flux_row<ND, 0, MSF>(JS, usfc, sn, hdt, nsf, acc[0]);
        flux_row<ND, 1, MSF>(JS, usfc, sn, hdt, nsf, acc[1]);
        flux_row<ND, 2, MSF>(JS, usfc, sn, hdt, nsf, acc[2]);
        flux_row<ND, 3, MSF>(JS, usfc, sn, hdt, nsf, acc[3]);
        if (NEQ > 4) { flux_row<ND, (NEQ > 4 ? 4 : 0), MSF>(JS, usfc, sn, hdt, nsf, acc[NEQ - 1]); }
    }
    if (d2min < CUDART_INF) { dist = sqrt_fast(d2min, dist_ok); } // no finite face distance: dist stays DBL_MAX like in the reference
    double const cvol = soln_early(TYPE) ? st[L::R_CV * TB] : G.cevol[i];
#if SCG_SHARED_RCP
    DivBy<true> by_cvol(cvol);
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = by_cvol(acc[e]); }
    if (by_cvol.ok) { return; }
#endif
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { out[e] = acc[e] / cvol; }
}



// EulerCore::calc_dsoln for one cell, euler_gradient.cpp:23-130, with the GradientElement constructor
// (GradientElement.cpp:137-213) and calc_gge_centroid (:81-133) expanded in place.  `un` is the cell's
// own so0n (energy already fixed by calc_cfl), `acfl_in` is cflc of the cell.
template <int TYPE>
__device__ __forceinline__ void cell_dsoln(MdField const & F, MdGeo const & G, int32_t i, double const (&un)[ct_ndim(TYPE) + 2],
                                           double cflc, double (&out)[ct_ndim(TYPE) + 2][ct_ndim(TYPE)])
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    constexpr int NFGE = ct_nfge(TYPE);
    int64_t const n = G.n;
    int64_t const S = F.S;
    int64_t const k = G.beg + i;
    double const hdt = F.dt * 0.5;
    double const acfl = fabs(cflc);
    double const sgm0 = F.sigma0 / acfl;
    double const tau = F.taumin + acfl * F.tauscale;
    double const ofg1 = 1.0 / NFGE;

    // ---- gradient element geometry
    double icecnd[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { icecnd[d] = F.solpt[d * S + k]; }
    double gp[NFC][ND], jd[NFC][ND];
    int32_t rcl[NFC];
#pragma unroll
    for (int f = 0; f < NFC; ++f)
    {
        int32_t const j = G.nbr[f * n + i];
        rcl[f] = j;
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            double const mid = G.cecnd[(f * ND + d) * n + i];
            double const jce = F.solpt[d * S + j];
            gp[f][d] = mid + tau * (jce - mid);
            jd[f][d] = gp[f][d] - jce;
        }
    }
    double avg[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { avg[d] = 0.0; }
#pragma unroll
    for (int f = 0; f < NFC; ++f)
    {
#pragma unroll
        for (int d = 0; d < ND; ++d) { avg[d] += gp[f][d]; }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) { avg[d] /= NFC; }
    double cnd[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { cnd[d] = 0.0; }
    double voc = 0.0;
#pragma unroll
    for (int isub = 0; isub < NFGE; ++isub)
    {
        double subcnd[ND];
        double dst[ND][ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { subcnd[d] = avg[d]; }
#pragma unroll
        for (int ivx = 0; ivx < ND; ++ivx)
        {
            int const fl = fge_face(TYPE, isub, ivx);
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                subcnd[d] += gp[fl][d];
                dst[ivx][d] = gp[fl][d] - avg[d];
            }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) { subcnd[d] /= static_cast<double>(ND + 1); }
        double const vob = det_n<ND>(dst);
        voc += vob;
#pragma unroll
        for (int d = 0; d < ND; ++d) { cnd[d] += subcnd[d] * vob; }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) { cnd[d] /= voc; }

    // ---- per-face displacement and Taylor difference (shared by the FGEs that use the face)
    double idis[NFC][ND];
    double udf[NFC][NEQ];
#pragma unroll
    for (int f = 0; f < NFC; ++f)
    {
        double jdis[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            idis[f][d] = gp[f][d] - cnd[d];
            jdis[d] = jd[f][d] + icecnd[d] - cnd[d];
        }
        int64_t const j = rcl[f];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double val = F.so0c[e * S + j] + hdt * F.so0t[e * S + j] - un[e];
#pragma unroll
            for (int d = 0; d < ND; ++d) { val += jdis[d] * F.so1c[(e * ND + d) * S + j]; }
            udf[f][e] = val;
        }
    }

    // ---- per-FGE gradients and W-1/2 weights
    double grad[NFGE][NEQ][ND];
    double widv[NFGE][NEQ];
    double wacc[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { wacc[e] = 0.0; }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        double dst[ND][ND], adj[ND][ND];
#pragma unroll
        for (int ivx = 0; ivx < ND; ++ivx)
        {
            int const fl = fge_face(TYPE, ifge, ivx);
#pragma unroll
            for (int d = 0; d < ND; ++d) { dst[ivx][d] = idis[fl][d]; }
        }
        adj_n<ND>(dst, adj);
        double const det = det_n<ND>(dst);
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double sq = 0.0;
#pragma unroll
            for (int a = 0; a < ND; ++a)
            {
                double r = 0.0;
#pragma unroll
                for (int b = 0; b < ND; ++b) { r += adj[a][b] * udf[fge_face(TYPE, ifge, b)][e]; }
                grad[ifge][e][a] = r / det;
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) { sq += grad[ifge][e][d] * grad[ifge][e][d]; }
            double const wgt = 1.0 / sqrt(sq + MD_ALMOST_ZERO);
            wacc[e] += wgt;
            widv[ifge][e] = wgt;
        }
    }

    // ---- W-3/4 limiter (euler_gradient.cpp:90-128)
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
        double wmax = 0.0, wmin = 0.0;
#pragma unroll
        for (int ifge = 0; ifge < NFGE; ++ifge)
        {
            double const wgt = widv[ifge][e] / wacc[e] - ofg1;
            widv[ifge][e] = wgt;
            wmax = fmax(wmax, wgt);
            wmin = fmin(wmin, wgt);
        }
        double const sm = fmin((1.0 - ofg1) / (wmax + MD_ALMOST_ZERO), -ofg1 / (wmin - MD_ALMOST_ZERO));
        double const sigma_max = fmin(sm, sgm0);
        double o[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { o[d] = 0.0; }
#pragma unroll
        for (int isub = 0; isub < NFGE; ++isub)
        {
            double const wgt = ofg1 + sigma_max * widv[isub][e];
#pragma unroll
            for (int d = 0; d < ND; ++d) { o[d] += wgt * grad[isub][e][d]; }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) { out[e][d] = o[d]; }
    }
}

// EulerCore::calc_dsoln for one cell, equation-outer formulation (march mode 3).
//
// cell_dsoln keeps every FGE gradient of every equation alive until the limiter has run (about 200
// doubles per cell).  The limiter only couples the gradients of ONE equation, so here the equation-
// independent part of every gradient element -- adj(D) and det(D), euler_gradient.cpp:74-77 -- is formed
// once and parked in shared memory (this thread's column of `sadj`, [NFGE][ND*ND+1][TB]); then one
// equation at a time takes its Taylor differences (euler_gradient.cpp:60-72), solves the NFGE systems,
// weighs and limits them (:79-128) and stores its ND gradient components.  The neighbour gather of
// equation e+1 is issued into a second register set before the arithmetic of equation e starts.  Same
// operations on the same operands in the same order as cell_dsoln; only the schedule differs.
// Cell types (bit t - CT_QUAD) for which the refined reciprocal of det(D) is parked in shared memory next to det(D)
// instead of being refined again by every equation (5 DFMA + 1 MUFU per gradient element and equation).  Measured
// (profiles/r02f_dsoln_yrow_ab.txt): tetrahedra 6 % faster, hexahedra / quadrilaterals / triangles 1-2 %, prisms 16 %
// slower (the extra rows cost them L1).
#ifndef SCG_EQO_YROW_MASK
#define SCG_EQO_YROW_MASK 0x0f // quad, tri, hex, tet
#endif
__host__ __device__ constexpr int eqo_yrow(int type) { return (SCG_EQO_YROW_MASK >> (type - CT_QUAD)) & 1; }
template <int TYPE, int TB>
struct DsolnEqo
{
    static constexpr int ND = ct_ndim(TYPE);
    static constexpr int NFGE = ct_nfge(TYPE);
    static constexpr int RPF = ND * ND + 1 + eqo_yrow(TYPE); // rows per gradient element: adj, det (, refined 1/det)
    // the cell's new gradient so1n[NEQ][ND] for the fused calc_solt: parked in shared memory in 2D (8 rows, 5 % faster
    // than reading it back); in 3D 15 more rows per thread would shrink L1 by as much and the gathers of the
    // equation loop live on L1 (tetrahedra: 318 -> 452 us), so there it is read back from L2
    static constexpr bool G_IN_SMEM = (2 == ND) || (0 != (SCG_EQO_GSMEM_MASK & (1 << (TYPE - CT_HEX))));
    static constexpr int R_G = NFGE * RPF;
    static constexpr int R_RCL = R_G + (G_IN_SMEM ? (ND + 2) * ND : 0); // neighbour indices: ct_nfc int32 rows (SCG_EQO_RCL_SMEM)
    static constexpr int R_SG = R_RCL + (eqo_rcl_smem(TYPE) ? (ct_nfc(TYPE) + 1) / 2 : 0); // sigma0 / |cfl| (SCG_EQO_SGM_SMEM)
    static constexpr int JF = eqo_jdis_faces(TYPE);                                        // faces whose jdis live in shared memory
    static constexpr int R_JD = R_SG + eqo_sgm_smem(TYPE);                                 // jdis[NFC - JF ..][ND]
    static constexpr int ROWS = R_JD + JF * ND;
    static constexpr int BYTES = ROWS * TB * int(sizeof(double));
};

// The arithmetic of one equation of calc_dsoln after the Taylor differences (euler_gradient.cpp:74-110):
// the NFGE gradients, their W-1/2 weights normalised and shifted, and sigma_max.  FAST: see above;
// returns whether every range test passed.
struct EqoNoMid
{
    __device__ __forceinline__ void operator()() const {}
};
// `mid` runs right after the products adj(D) * udf: from there on the equation needs no gathered value, so the caller may
// start the next equation's gather into the very registers the Taylor differences were formed from
template <int TYPE, int TB, bool FAST, typename MID = EqoNoMid>
__device__ __forceinline__ bool eqo_solve(double const * sadj, double const (&udf)[ct_nfc(TYPE)], double ofg1, double sgm0,
                                          double (&grad)[ct_nfge(TYPE)][ct_ndim(TYPE)], double (&w)[ct_nfge(TYPE)], double & sigma_max,
                                          MID && mid = MID())
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NFGE = ct_nfge(TYPE);
    constexpr int RPF = ND * ND + 1 + eqo_yrow(TYPE);
    bool ok = true;
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
#pragma unroll
        for (int a = 0; a < ND; ++a)
        {
            // r = 0.0 + p0 + p1 + ..: the leading "0.0 +" is dropped when SCG_SKIP_ZERO_ADD: it can only turn a -0 into a
            // +0, and the sign of a zero gradient cannot reach so1n (see SCG_EQO_ZERO)
            double r = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b)
            {
                double const p = sadj[(ifge * RPF + a * ND + b) * TB] * udf[fge_face(TYPE, ifge, b)];
                r = (SCG_SKIP_ZERO_ADD && 0 == b) ? p : r + p;
            }
            grad[ifge][a] = r;
        }
    }
    mid();
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        DivBy<FAST> by_det = (FAST && eqo_yrow(TYPE)) ? DivBy<FAST>(sadj[(ifge * RPF + ND * ND) * TB], sadj[(ifge * RPF + RPF - 1) * TB])
                                                      : DivBy<FAST>(sadj[(ifge * RPF + ND * ND) * TB]);
#pragma unroll
        for (int a = 0; a < ND; ++a) { grad[ifge][a] = by_det.template quot<SCG_EQO_ZERO>(grad[ifge][a]); }
        ok = ok && by_det.ok;
    }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        double sq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) { sq += grad[ifge][d] * grad[ifge][d]; }
        w[ifge] = sq + MD_ALMOST_ZERO;
    }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge) { w[ifge] = FAST ? sqrt_fast(w[ifge], ok) : sqrt(w[ifge]); }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge) { w[ifge] = FAST ? rcp_fast(w[ifge], ok) : 1.0 / w[ifge]; }
    double wacc = 0.0;
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge) { wacc += w[ifge]; }
    // ---- W-3/4 limiter (euler_gradient.cpp:90-110)
    {
        DivBy<FAST> by_wacc(wacc);
#pragma unroll
        for (int ifge = 0; ifge < NFGE; ++ifge) { w[ifge] = by_wacc.template quot<0>(w[ifge]) - ofg1; }
        ok = ok && by_wacc.ok;
    }
    double wmax = 0.0, wmin = 0.0;
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        wmax = fmax(wmax, w[ifge]);
        wmin = fmin(wmin, w[ifge]);
    }
    DivBy<FAST> by_hi(wmax + MD_ALMOST_ZERO);
    DivBy<FAST> by_lo(wmin - MD_ALMOST_ZERO);
    double const sm = fmin(by_hi.template quot<0>(1.0 - ofg1), by_lo.template quot<0>(-ofg1));
    ok = ok && by_hi.ok && by_lo.ok;
    sigma_max = fmin(sm, sgm0);
    return ok;
}

// what one equation of calc_dsoln gathers through the neighbour indices (euler_gradient.cpp:66-69)
template <int TYPE>
struct EqoGather
{
    static constexpr int ND = ct_ndim(TYPE);
    static constexpr int NFC = ct_nfc(TYPE);
    double un;
#if SCG_SW_SHARE
    double u[NFC], g[NFC][ND];
#else
    double u[NFC], t[NFC], g[NFC][ND];
#endif
    __device__ __forceinline__ void load(MdField const & F, int32_t const (&rcl)[NFC], int64_t k, int e)
    {
        int64_t const S = F.S;
        un = F.so0n[e * S + k];
#pragma unroll
        for (int f = 0; f < NFC; ++f)
        {
            int64_t const j = rcl[f];
#if SCG_SW_SHARE
            u[f] = F.sw[e * S + j]; // so0c + hdt * so0t, the first operation of the Taylor difference (euler_gradient.cpp:66)
#else
            u[f] = F.so0c[e * S + j];
            t[f] = F.so0t[e * S + j];
#endif
#pragma unroll
            for (int d = 0; d < ND; ++d) { g[f][d] = F.so1c[(e * ND + d) * S + j]; }
        }
    }
};

// PFM: how the neighbour gather of an equation is overlapped with arithmetic.  0: not at all (loaded at the top of its equation);
// 1: loaded at the top of the PREVIOUS equation into a second register set; 2: loaded in the MIDDLE of the previous equation, right
// after its products adj(D) * udf, into the same registers (no second set; ~60 % of an equation's arithmetic covers the latency);
// 3: like 2 but right after the Taylor differences are formed, i.e. as soon as the registers are free
template <int TYPE, int TB, int PFM>
__device__ __forceinline__ void cell_dsoln_eqo(MdField const & F, MdGeo const & G, int32_t i, double * sadj)
{
    constexpr bool PF = (1 == PFM);
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    constexpr int NFGE = ct_nfge(TYPE);
    constexpr int RPF = ND * ND + 1 + eqo_yrow(TYPE);
    int64_t const n = G.n;
    int64_t const S = F.S;
    int64_t const k = G.beg + i;
    int64_t const k_cell = k;
    double const hdt = F.dt * 0.5;
    double const acfl = fabs(F.cflc[k]);
    double const sgm0 = F.sigma0 / acfl;
    double const tau = F.taumin + acfl * F.tauscale;
    double const ofg1 = 1.0 / NFGE;

    int32_t rcl[NFC];
    // this thread's column of the int32 rows behind the double rows (SCG_EQO_RCL_SMEM)
    int32_t * const srcl = reinterpret_cast<int32_t *>(sadj - threadIdx.x + DsolnEqo<TYPE, TB>::R_RCL * TB) + threadIdx.x;
    double jdis[NFC][ND];
    EqoGather<TYPE> cur;
#if SCG_EQO_GAMMA_PF
    // the fused calc_solt reads the cell's gamma after the equation loop -- the one line of its inputs nobody touched before
    if (!(PF && SCG_EQO_SOLT_PF) && F.fuse_solt) { l2_prefetch(&F.gamma[k]); }
#endif
    {
        // ---- gradient element geometry (GradientElement.cpp:137-213, :81-133)
        double icecnd[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { icecnd[d] = F.solpt[d * S + k]; }
        double gp[NFC][ND];
#pragma unroll
        for (int f = 0; f < NFC; ++f) { rcl[f] = G.nbr[f * n + i]; }
        if (eqo_rcl_smem(TYPE))
        {
#pragma unroll
            for (int f = 0; f < NFC; ++f) { srcl[f * TB] = rcl[f]; }
        }
        if (0 != PFM) { cur.load(F, rcl, k, 0); } // the first equation's gather travels while the geometry is worked out
#if SCG_EQO_L2PF
        // every later equation's gather is requested from HBM now; it is consumed e equations (microseconds) later
#pragma unroll
        for (int e = (PF ? 2 : 1); e < NEQ; ++e)
        {
            l2_prefetch(&F.so0n[e * S + k]);
#pragma unroll
            for (int f = 0; f < NFC; ++f)
            {
                int64_t const j = rcl[f];
                l2_prefetch(&F.so0c[e * S + j]);
                l2_prefetch(&F.so0t[e * S + j]);
#pragma unroll
                for (int d = 0; d < ND; ++d) { l2_prefetch(&F.so1c[(e * ND + d) * S + j]); }
            }
        }
#endif
#pragma unroll
        for (int f = 0; f < NFC; ++f)
        {
            int64_t const j = rcl[f];
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                double const mid = G.cecnd[(f * ND + d) * n + i];
                double const jce = F.solpt[d * S + j];
                gp[f][d] = mid + tau * (jce - mid);
                jdis[f][d] = gp[f][d] - jce; // completed below: + icecnd - cnd
            }
        }
        double avg[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { avg[d] = 0.0; }
#pragma unroll
        for (int f = 0; f < NFC; ++f)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { avg[d] += gp[f][d]; }
        }
        {
            double a2[ND];
            DivBy<true> by_nfc(static_cast<double>(NFC));
#pragma unroll
            for (int d = 0; d < ND; ++d) { a2[d] = by_nfc.template quot<0>(avg[d]); }
            if (!by_nfc.ok)
            {
#pragma unroll
                for (int d = 0; d < ND; ++d) { a2[d] = avg[d] / static_cast<double>(NFC); }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) { avg[d] = a2[d]; }
        }
        double cnd[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { cnd[d] = 0.0; }
        double voc = 0.0;
#pragma unroll
        for (int isub = 0; isub < NFGE; ++isub)
        {
            double subcnd[ND];
            double dst[ND][ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { subcnd[d] = avg[d]; }
#pragma unroll
            for (int ivx = 0; ivx < ND; ++ivx)
            {
                int const fl = fge_face(TYPE, isub, ivx);
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    subcnd[d] += gp[fl][d];
                    dst[ivx][d] = gp[fl][d] - avg[d];
                }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) { subcnd[d] /= static_cast<double>(ND + 1); }
            double const vob = det_n<ND>(dst);
            voc += vob;
#pragma unroll
            for (int d = 0; d < ND; ++d) { cnd[d] += subcnd[d] * vob; }
        }
        {
            double c2[ND];
            DivBy<true> by_voc(voc);
#pragma unroll
            for (int d = 0; d < ND; ++d) { c2[d] = by_voc.template quot<0>(cnd[d]); }
            if (!by_voc.ok)
            {
#pragma unroll
                for (int d = 0; d < ND; ++d) { c2[d] = cnd[d] / voc; }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) { cnd[d] = c2[d]; }
        }
        // ---- displacements (GradientElement.cpp:196-211); idis replaces gp
#pragma unroll
        for (int f = 0; f < NFC; ++f)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                jdis[f][d] = jdis[f][d] + icecnd[d] - cnd[d];
                gp[f][d] = gp[f][d] - cnd[d];
            }
        }
        if (eqo_sgm_smem(TYPE)) { sadj[DsolnEqo<TYPE, TB>::R_SG * TB] = sgm0; }
#pragma unroll
        for (int f = NFC - DsolnEqo<TYPE, TB>::JF; f < NFC; ++f)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { sadj[(DsolnEqo<TYPE, TB>::R_JD + (f - (NFC - DsolnEqo<TYPE, TB>::JF)) * ND + d) * TB] = jdis[f][d]; }
        }
        // ---- adj(D), det(D) of every gradient element -> shared memory
#pragma unroll
        for (int ifge = 0; ifge < NFGE; ++ifge)
        {
            double dst[ND][ND], adj[ND][ND];
#pragma unroll
            for (int ivx = 0; ivx < ND; ++ivx)
            {
                int const fl = fge_face(TYPE, ifge, ivx);
#pragma unroll
                for (int d = 0; d < ND; ++d) { dst[ivx][d] = gp[fl][d]; }
            }
            adj_n<ND>(dst, adj);
            double const det = det_n<ND>(dst);
#pragma unroll
            for (int a = 0; a < ND; ++a)
            {
#pragma unroll
                for (int b = 0; b < ND; ++b) { sadj[(ifge * RPF + a * ND + b) * TB] = adj[a][b]; }
            }
            sadj[(ifge * RPF + ND * ND) * TB] = det;
            if (eqo_yrow(TYPE)) { sadj[(ifge * RPF + RPF - 1) * TB] = DivBy<true>(det).y; }
        }
    }

    // the fused calc_solt's inputs ride in the prefetch registers of the last equation: needs a slot per equation in u[] / g[]
    constexpr bool SOLT_PF = PF && SCG_EQO_SOLT_PF && !DsolnEqo<TYPE, TB>::G_IN_SMEM && NFC >= NEQ;
    double olast[ND]; // the last equation's gradient (the others are read back)
    // ---- one equation at a time
#pragma unroll 1
    for (int e = 0; e < NEQ; ++e)
    {
        // adj/det are re-read from shared memory by every equation: without this barrier the compiler hoists
        // the NFGE*(ND*ND+1) loop-invariant loads out of the loop and spills them
        asm volatile("" ::: "memory");
        if (eqo_rcl_smem(TYPE))
        {
#pragma unroll
            for (int f = 0; f < NFC; ++f) { rcl[f] = srcl[f * TB]; }
        }
        int64_t k = k_cell; // see SCG_EQO_OPAQUE_K
        if (eqo_opaque_k(TYPE)) { asm volatile("" : "+l"(k)); }
        EqoGather<TYPE> nxt;
        if (PF && !SCG_EQO_PF_AFTER_UDF)
        {
            if (SOLT_PF && e + 1 == NEQ)
            {
                // nothing is left to gather: the prefetch registers take the inputs of the fused calc_solt -- the cell's so0n and
                // gamma and the gradient rows this thread stored in the earlier iterations -- which travel during the last equation
                nxt.un = F.gamma[k];
#pragma unroll
                for (int q = 0; q < NEQ; ++q) { nxt.u[q] = F.so0n[q * S + k]; }
#pragma unroll
                for (int q = 0; q < NEQ - 1; ++q)
                {
#pragma unroll
                    for (int d = 0; d < ND; ++d) { nxt.g[q][d] = __ldcg(&F.so1n[(q * ND + d) * S + k]); }
                }
            }
            else { nxt.load(F, rcl, k, e + 1 < NEQ ? e + 1 : e); }
        }
        else if (0 == PFM) { cur.load(F, rcl, k, e); }
        double udf[NFC];
#pragma unroll
        for (int f = 0; f < NFC; ++f)
        {
#if SCG_SW_SHARE
            double val = cur.u[f] - cur.un;
#else
            double val = cur.u[f] + hdt * cur.t[f] - cur.un;
#endif
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                double const jd = (f >= NFC - DsolnEqo<TYPE, TB>::JF) ? sadj[(DsolnEqo<TYPE, TB>::R_JD + (f - (NFC - DsolnEqo<TYPE, TB>::JF)) * ND + d) * TB]
                                                                      : jdis[f][d];
                val += jd * cur.g[f][d];
            }
            udf[f] = val;
        }
        if (PF && SCG_EQO_PF_AFTER_UDF)
        {
            // The second register set is loaded AFTER the last use of the first: ptxas shares its few scoreboards among all loads in
            // flight, so a gather issued before `cur` is consumed makes that consumer wait for the new loads too (profiles/r02h: 11.6 %
            // of the samples of hex dsoln on that one DADD).  The gather addresses are made to depend on the last Taylor difference,
            // which keeps the scheduler from hoisting the loads back to the top of the equation.
#pragma unroll
            for (int f = 0; f < NFC; ++f) { asm volatile("" : "+r"(rcl[f]) : "d"(udf[NFC - 1])); }
            nxt.load(F, rcl, k, e + 1 < NEQ ? e + 1 : e);
        }
        if (PF) { cur = nxt; }
        if (3 == PFM)
        {
            // the Taylor differences are complete (the asm consumes them, so the loads below cannot be scheduled above their last
            // use of `cur`): the next equation's gather goes into the same registers and travels during the whole solve.  Issued
            // AFTER the last use because ptxas shares its few scoreboards among all loads in flight: a gather issued before `cur`
            // is consumed makes that consumer wait for the new loads too (profiles/r02h: 11.6 % of hex dsoln on that one DADD)
#if SCG_EQO_PF3_PIN
#pragma unroll
            for (int f = 0; f < NFC; ++f) { asm volatile("" : "+d"(udf[f])::"memory"); }
#else
            asm volatile("" ::: "memory");
#endif
            if (e + 1 < NEQ)
            {
                if (eqo_rcl_smem(TYPE))
                {
#pragma unroll
                    for (int f = 0; f < NFC; ++f) { rcl[f] = srcl[f * TB]; }
                }
                cur.load(F, rcl, k, e + 1);
            }
        }
        double grad[NFGE][ND], w[NFGE];
        double sigma_max;
        // every division, square root and reciprocal of the equation runs its inlined fast sequence with
        // the range tests collected in one flag, so the hot path is one basic block; when any test fails the
        // equation is redone with the plain operators (cold)
        double const sg = eqo_sgm_smem(TYPE) ? sadj[DsolnEqo<TYPE, TB>::R_SG * TB] : sgm0;
        auto mid = [&]() {
            if (2 == PFM)
            {
                asm volatile("" ::: "memory"); // the gather is issued HERE, not hoisted to the top of the equation
                if (e + 1 < NEQ)
                {
                    if (eqo_rcl_smem(TYPE))
                    {
#pragma unroll
                        for (int f = 0; f < NFC; ++f) { rcl[f] = srcl[f * TB]; }
                    }
                    cur.load(F, rcl, k, e + 1);
                }
            }
        };
        if (!eqo_solve<TYPE, TB, true>(sadj, udf, ofg1, sg, grad, w, sigma_max, mid))
        {
            eqo_solve<TYPE, TB, false>(sadj, udf, ofg1, sg, grad, w, sigma_max);
        }
        double o[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { o[d] = 0.0; }
#pragma unroll
        for (int isub = 0; isub < NFGE; ++isub)
        {
            double const wgt = ofg1 + sigma_max * w[isub];
#pragma unroll
            for (int d = 0; d < ND; ++d) { o[d] += wgt * grad[isub][d]; }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            F.so1n[(e * ND + d) * S + k] = o[d];
            if (DsolnEqo<TYPE, TB>::G_IN_SMEM) { sadj[(DsolnEqo<TYPE, TB>::R_G + e * ND + d) * TB] = o[d]; }
            olast[d] = o[d];
        }
    }
    // ---- fused calc_solt of the NEXT substep (euler_march.cpp:161-191): this cell's so0n and so1n are final, and
    //      after update() they are its so0c and so1c (see DsolnEqo::G_IN_SMEM for where the gradient comes from)
    if (F.fuse_solt)
    {
        if (DsolnEqo<TYPE, TB>::G_IN_SMEM) { asm volatile("" ::: "memory"); }
        double u[NEQ], g[NEQ][ND], t[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            u[e] = SOLT_PF ? cur.u[e < NFC ? e : 0] : F.so0n[e * S + k];
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                g[e][d] = SOLT_PF ? (e + 1 < NEQ ? cur.g[e < NFC ? e : 0][d] : olast[d])
                          : DsolnEqo<TYPE, TB>::G_IN_SMEM ? sadj[(DsolnEqo<TYPE, TB>::R_G + e * ND + d) * TB]
                                                          : __ldcg(&F.so1n[(e * ND + d) * S + k]);
            }
        }
        Jaco<ND> J;
        jaco_update<ND>(J, SOLT_PF ? cur.un : F.gamma[k], u);
        solt_row<ND>(J, g, t);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.so0t_next[e * S + k] = t[e]; }
#if SCG_SW_SHARE
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.sw_next[e * S + k] = u[e] + (F.dt * 0.5) * t[e]; }
#endif
#if SCG_CT_PAIR
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.ct_next[e * S + k] = make_double2(u[e], t[e]); }
#endif
#if SCG_JS_SHARE
        {
            JacoS<ND> Q;
            Q.set(F.gamma[k], u);
            Q.pack(F.js_next, S, k);
        }
#endif
    }
}

// ------------------------------------------------------------------------------------------------
// calc_dsoln, equation-outer, second generation (3D): every global load happens before the equation loop.
//
// Three facts about cell_dsoln_eqo's profile drive this version (profiles/r01f_*, r02d_fp64_pipe_ubench.txt):
// a dependent fp64 instruction waits ~20 cycles, so a scheduler needs >= 10 independent ones in flight and two
// warps of 255 registers do not supply them; the gather at the top of every equation is an exposed memory round
// trip (5 per cell); and the register prefetch that hides it for hexahedra costs 62 registers.  Here
//   * the Taylor differences udf[f][e] of ALL equations (euler_gradient.cpp:60-72) are formed in the prologue, two
//     faces per batch of loads, and parked in shared memory, so the equation loop touches no global memory and
//     keeps neither neighbour indices nor displacements alive;
//   * adj(D) is stored once per PAIR of faces: its columns are the cross products idis[p] x idis[q] of the element's
//     face displacements (GradientElement.cpp:215-249 written out), a pair of faces is shared by two gradient
//     elements, and x*y - z*w == -(z*w - x*y) bit for bit, so hexahedra keep 12 x 3 rows instead of 8 x 9 (a zero
//     entry may change sign, which cannot reach the sum it enters: that sum starts from +0);
//   * the rows of an equation that is finished are reused for its new gradient and so0n, which the fused calc_solt
//     reads back from shared memory.
// Same operations on the same operands in the same order as cell_dsoln.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fge_pair_first(int type, int u) // local face pair of use u = ifge*3 + column: (p, q) with column = p x q
{
    return fge_face(type, u / 3, (u % 3 + 1) % 3);
}
__host__ __device__ constexpr int fge_pair_second(int type, int u) { return fge_face(type, u / 3, (u % 3 + 2) % 3); }
__host__ __device__ constexpr int fge_pair_key(int type, int u)
{
    int const p = fge_pair_first(type, u), q = fge_pair_second(type, u);
    return p < q ? p * 8 + q : q * 8 + p;
}
__host__ __device__ constexpr bool fge_pair_neg(int type, int u) { return fge_pair_first(type, u) > fge_pair_second(type, u); }
__host__ __device__ constexpr int fge_pair_firstuse(int type, int u)
{
    for (int v = 0; v < u; ++v)
    {
        if (fge_pair_key(type, v) == fge_pair_key(type, u)) { return v; }
    }
    return u;
}
__host__ __device__ constexpr int fge_pair_id(int type, int u) // dense index of the pair among the distinct ones
{
    int const v = fge_pair_firstuse(type, u);
    int id = 0;
    for (int w = 0; w < v; ++w) { id += (fge_pair_firstuse(type, w) == w) ? 1 : 0; }
    return id;
}
__host__ __device__ constexpr int fge_npair(int type)
{
    int n = 0;
    for (int w = 0; w < 3 * ct_nfge(type); ++w) { n += (fge_pair_firstuse(type, w) == w) ? 1 : 0; }
    return n;
}
static_assert(12 == fge_npair(CT_HEX) && 6 == fge_npair(CT_TET) && 9 == fge_npair(CT_PRISM), "face pairs of the gradient elements");

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E) -- the index is a constant expression inside f,
// which forces the constexpr table functions above to be evaluated by the compiler and not by the device
template <int B, int E, typename Fn>
__device__ __forceinline__ void static_for(Fn && f)
{
    if constexpr (B < E)
    {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

#ifndef SCG_EQO2_UNROLL
#define SCG_EQO2_UNROLL 1 // equations per trip of the equation loop (2: the serial tail of one overlaps the wide head of the next)
#endif
#ifndef SCG_EQO2_YROWS
#define SCG_EQO2_YROWS 0 // 1: the refined reciprocal of det(D) is parked next to det(D) instead of being refined again by every equation
#endif

template <int TYPE, int TB>
struct DsolnEqo2
{
    static constexpr int ND = 3;
    static constexpr int NEQ = 5;
    static constexpr int NFC = ct_nfc(TYPE);
    static constexpr int NFGE = ct_nfge(TYPE);
    static constexpr int NP = fge_npair(TYPE);
    static constexpr int R_U = 0;                  // Taylor differences [NFC][NEQ]; equation e, once finished: rows
                                                   // [0..2][e] = its new gradient, [3][e] = so0n of the cell
    static constexpr int R_X = R_U + NFC * NEQ;    // cross products [NP][3]
    static constexpr int R_DET = R_X + NP * 3;     // det(D) [NFGE]
    static constexpr int R_END = R_DET + NFGE;
    // Prologue staging: the neighbour rows of a face (so0c[5], so0t[5], so1c[15]) travel global -> shared by cp.async,
    // two faces in flight, into two buffers at the END of the column; they overlay the cross products and det(D),
    // which are written after the last face has been consumed.
    static constexpr int FROWS = 5 * NEQ;          // 25 rows per face
    static constexpr int ROWS = (TYPE == CT_HEX) ? 74 : (TYPE == CT_TET) ? 55 : 65;
    static constexpr int R_A = ROWS - 2 * FROWS;
    static constexpr int R_B = ROWS - FROWS;
    static_assert(ROWS >= R_END, "the staging column holds the final layout");
    static constexpr int BYTES = ROWS * TB * int(sizeof(double));
    __host__ __device__ static constexpr int buf(int f) { return (f & 1) ? R_B : R_A; }
    // May udf[f] be stored right after face f has been consumed?  Its rows must not lie in a buffer that still
    // holds (or is about to receive) an unconsumed face: face f+1's buffer, and face f's own if it is refilled with f+2.
    __host__ __device__ static constexpr bool store_now(int f)
    {
        int const lo = R_U + f * NEQ, hi = lo + NEQ;
        bool ok = true;
        if (f + 1 < NFC) { ok = ok && (hi <= buf(f + 1) || lo >= buf(f + 1) + FROWS); }
        if (f + 2 < NFC) { ok = ok && (hi <= buf(f) || lo >= buf(f) + FROWS); }
        return ok;
    }
};

template <int TYPE, int TB>
__device__ __forceinline__ void eqo2_stage_face(MdField const & F, int64_t j, double * st)
{
    int64_t const S = F.S;
#pragma unroll
    for (int e = 0; e < 5; ++e)
    {
        cp_async8(st + e * TB, &F.so0c[e * S + j]);
        cp_async8(st + (5 + e) * TB, &F.so0t[e * S + j]);
#pragma unroll
        for (int d = 0; d < 3; ++d) { cp_async8(st + (10 + e * 3 + d) * TB, &F.so1c[(e * 3 + d) * S + j]); }
    }
    cp_async_commit();
}

template <int TYPE, int TB, bool FAST>
__device__ __forceinline__ bool eqo2_solve(double const * sm, int e, double ofg1, double sgm0,
                                           double (&grad)[ct_nfge(TYPE)][3], double (&w)[ct_nfge(TYPE)], double & sigma_max)
{
    using L = DsolnEqo2<TYPE, TB>;
    constexpr int ND = 3;
    constexpr int NFC = L::NFC;
    constexpr int NFGE = L::NFGE;
    bool ok = true;
    double udf[NFC];
#pragma unroll
    for (int f = 0; f < NFC; ++f) { udf[f] = sm[(L::R_U + f * L::NEQ + e) * TB]; }
    static_for<0, NFGE>([&](auto IFGE) {
        constexpr int ifge = decltype(IFGE)::value;
#pragma unroll
        for (int a = 0; a < ND; ++a)
        {
            double r = 0.0;
            static_for<0, ND>([&](auto B) {
                constexpr int b = decltype(B)::value;
                constexpr int u = ifge * 3 + b;
                constexpr int row = L::R_X + fge_pair_id(TYPE, u) * 3;
                constexpr int fl = fge_face(TYPE, ifge, b);
                constexpr bool neg = fge_pair_neg(TYPE, u);
                double const p = sm[(row + a) * TB] * udf[fl];
                r = neg ? r - p : r + p;
            });
            grad[ifge][a] = r;
        }
    });
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        DivBy<FAST> by_det(sm[(L::R_DET + ifge) * TB]);
#pragma unroll
        for (int a = 0; a < ND; ++a) { grad[ifge][a] = by_det.template quot<SCG_EQO_ZERO>(grad[ifge][a]); }
        ok = ok && by_det.ok;
    }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        double sq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) { sq += grad[ifge][d] * grad[ifge][d]; }
        w[ifge] = sq + MD_ALMOST_ZERO;
    }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge) { w[ifge] = FAST ? sqrt_fast(w[ifge], ok) : sqrt(w[ifge]); }
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge) { w[ifge] = FAST ? rcp_fast(w[ifge], ok) : 1.0 / w[ifge]; }
    double wacc = 0.0;
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge) { wacc += w[ifge]; }
    {
        DivBy<FAST> by_wacc(wacc);
#pragma unroll
        for (int ifge = 0; ifge < NFGE; ++ifge) { w[ifge] = by_wacc.template quot<0>(w[ifge]) - ofg1; }
        ok = ok && by_wacc.ok;
    }
    double wmax = 0.0, wmin = 0.0;
#pragma unroll
    for (int ifge = 0; ifge < NFGE; ++ifge)
    {
        wmax = fmax(wmax, w[ifge]);
        wmin = fmin(wmin, w[ifge]);
    }
    DivBy<FAST> by_hi(wmax + MD_ALMOST_ZERO);
    DivBy<FAST> by_lo(wmin - MD_ALMOST_ZERO);
    double const sm_ = fmin(by_hi.template quot<0>(1.0 - ofg1), by_lo.template quot<0>(-ofg1));
    ok = ok && by_hi.ok && by_lo.ok;
    sigma_max = fmin(sm_, sgm0);
    return ok;
}

template <int TYPE, int TB>
__device__ __forceinline__ void cell_dsoln_eqo2(MdField const & F, MdGeo const & G, int32_t i, double * sm)
{
    using L = DsolnEqo2<TYPE, TB>;
    constexpr int ND = 3;
    constexpr int NEQ = 5;
    constexpr int NFC = L::NFC;
    constexpr int NFGE = L::NFGE;
    int64_t const n = G.n;
    int64_t const S = F.S;
    int64_t const k = G.beg + i;
    double const hdt = F.dt * 0.5;
    double const ofg1 = 1.0 / NFGE;
    double sgm0;
    double ga_own;
    {
        int32_t rcl[NFC];
        double jdis[NFC][ND];
#pragma unroll
        for (int f = 0; f < NFC; ++f) { rcl[f] = G.nbr[f * n + i]; }
        // the neighbour rows of the first two faces travel while the geometry is worked out
        eqo2_stage_face<TYPE, TB>(F, rcl[0], sm + L::R_A * TB);
        eqo2_stage_face<TYPE, TB>(F, rcl[1], sm + L::R_B * TB);
        double const cflc_in = F.cflc[k];
        ga_own = F.gamma[k]; // for the fused calc_solt at the end: requested now, used last
        double un[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { un[e] = F.so0n[e * S + k]; }
        double gp[NFC][ND];
        {
            // ---- gradient element geometry (GradientElement.cpp:137-213, :81-133)
            double icecnd[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { icecnd[d] = F.solpt[d * S + k]; }
            double mid[NFC][ND], jce[NFC][ND];
#pragma unroll
            for (int f = 0; f < NFC; ++f)
            {
                int64_t const j = rcl[f];
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    mid[f][d] = G.cecnd[(f * ND + d) * n + i];
                    jce[f][d] = F.solpt[d * S + j];
                }
            }
            double const acfl = fabs(cflc_in);
            sgm0 = F.sigma0 / acfl;
            double const tau = F.taumin + acfl * F.tauscale;
#pragma unroll
            for (int f = 0; f < NFC; ++f)
            {
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    gp[f][d] = mid[f][d] + tau * (jce[f][d] - mid[f][d]);
                    jdis[f][d] = gp[f][d] - jce[f][d]; // completed below: + icecnd - cnd
                }
            }
            double avg[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { avg[d] = 0.0; }
#pragma unroll
            for (int f = 0; f < NFC; ++f)
            {
#pragma unroll
                for (int d = 0; d < ND; ++d) { avg[d] += gp[f][d]; }
            }
            {
                double a2[ND];
                DivBy<true> by_nfc(static_cast<double>(NFC));
#pragma unroll
                for (int d = 0; d < ND; ++d) { a2[d] = by_nfc.template quot<0>(avg[d]); }
                if (!by_nfc.ok)
                {
#pragma unroll
                    for (int d = 0; d < ND; ++d) { a2[d] = avg[d] / static_cast<double>(NFC); }
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) { avg[d] = a2[d]; }
            }
            double cnd[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { cnd[d] = 0.0; }
            double voc = 0.0;
#pragma unroll
            for (int isub = 0; isub < NFGE; ++isub)
            {
                double subcnd[ND];
                double dst[ND][ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { subcnd[d] = avg[d]; }
#pragma unroll
                for (int ivx = 0; ivx < ND; ++ivx)
                {
                    int const fl = fge_face(TYPE, isub, ivx);
#pragma unroll
                    for (int d = 0; d < ND; ++d)
                    {
                        subcnd[d] += gp[fl][d];
                        dst[ivx][d] = gp[fl][d] - avg[d];
                    }
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) { subcnd[d] /= static_cast<double>(ND + 1); }
                double const vob = det_n<ND>(dst);
                voc += vob;
#pragma unroll
                for (int d = 0; d < ND; ++d) { cnd[d] += subcnd[d] * vob; }
            }
            {
                double c2[ND];
                DivBy<true> by_voc(voc);
#pragma unroll
                for (int d = 0; d < ND; ++d) { c2[d] = by_voc.template quot<0>(cnd[d]); }
                if (!by_voc.ok)
                {
#pragma unroll
                    for (int d = 0; d < ND; ++d) { c2[d] = cnd[d] / voc; }
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) { cnd[d] = c2[d]; }
            }
            // ---- displacements (GradientElement.cpp:196-211); idis replaces gp
#pragma unroll
            for (int f = 0; f < NFC; ++f)
            {
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    jdis[f][d] = jdis[f][d] + icecnd[d] - cnd[d];
                    gp[f][d] = gp[f][d] - cnd[d];
                }
            }
        }
        // ---- Taylor differences of every equation (euler_gradient.cpp:60-72) -> shared memory.  Face f is consumed from
        //      its staging buffer, the buffer is refilled with face f+2, and udf[f] is stored at once when its rows are
        //      clear of the buffers still in use (DsolnEqo2::store_now), else after the last face.
        double uv[NFC][NEQ];
        static_for<0, NFC>([&](auto FC) {
            constexpr int f = decltype(FC)::value;
            if constexpr (f + 1 < NFC) { cp_async_wait<1>(); }
            else { cp_async_wait<0>(); }
            double const * st = sm + L::buf(f) * TB;
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                double val = st[e * TB] + hdt * st[(5 + e) * TB] - un[e];
#pragma unroll
                for (int d = 0; d < ND; ++d) { val += jdis[f][d] * st[(10 + e * ND + d) * TB]; }
                uv[f][e] = val;
            }
            if constexpr (f + 2 < NFC)
            {
                asm volatile("" ::: "memory"); // the buffer is refilled only after its last read
                eqo2_stage_face<TYPE, TB>(F, rcl[f + 2], sm + L::buf(f) * TB);
            }
            if constexpr (L::store_now(f))
            {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) { sm[(L::R_U + f * NEQ + e) * TB] = uv[f][e]; }
            }
        });
        static_for<0, NFC>([&](auto FC) {
            constexpr int f = decltype(FC)::value;
            if constexpr (!L::store_now(f))
            {
#pragma unroll
                for (int e = 0; e < NEQ; ++e) { sm[(L::R_U + f * NEQ + e) * TB] = uv[f][e]; }
            }
        });
        // ---- adj(D) as cross products of face pairs, det(D) of every gradient element -> shared memory (over the
        //      staging buffers, which are idle now)
        static_for<0, 3 * NFGE>([&](auto U) {
            constexpr int u = decltype(U)::value;
            if constexpr (fge_pair_firstuse(TYPE, u) == u)
            {
                constexpr int key = fge_pair_key(TYPE, u);
                constexpr int p = key / 8, q = key % 8; // p < q
                constexpr int row = L::R_X + fge_pair_id(TYPE, u) * 3;
                sm[(row + 0) * TB] = gp[p][1] * gp[q][2] - gp[p][2] * gp[q][1];
                sm[(row + 1) * TB] = gp[p][2] * gp[q][0] - gp[p][0] * gp[q][2];
                sm[(row + 2) * TB] = gp[p][0] * gp[q][1] - gp[p][1] * gp[q][0];
            }
        });
        static_for<0, NFGE>([&](auto IFGE) {
            constexpr int ifge = decltype(IFGE)::value;
            constexpr int f0 = fge_face(TYPE, ifge, 0), f1 = fge_face(TYPE, ifge, 1), f2 = fge_face(TYPE, ifge, 2);
            double dst[ND][ND];
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                dst[0][d] = gp[f0][d];
                dst[1][d] = gp[f1][d];
                dst[2][d] = gp[f2][d];
            }
            sm[(L::R_DET + ifge) * TB] = det_n<ND>(dst);
        });
    }
    {
        // ---- one equation at a time; no global loads but the cell's own so0n, which is parked for the fused calc_solt
        //      where equation e's fourth Taylor difference was
        constexpr int UNR = SCG_EQO2_UNROLL;
#pragma unroll UNR
        for (int e = 0; e < NEQ; ++e)
        {
            asm volatile("" ::: "memory"); // keep the loop-invariant shared-memory reads inside the loop
            double const un_e = F.so0n[e * S + k]; // cached by the prologue's read; wanted at the end of the iteration only
            double grad[NFGE][ND], w[NFGE];
            double sigma_max;
            if (!eqo2_solve<TYPE, TB, true>(sm, e, ofg1, sgm0, grad, w, sigma_max))
            {
                eqo2_solve<TYPE, TB, false>(sm, e, ofg1, sgm0, grad, w, sigma_max);
            }
            double o[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { o[d] = 0.0; }
#pragma unroll
            for (int isub = 0; isub < NFGE; ++isub)
            {
                double const wgt = ofg1 + sigma_max * w[isub];
#pragma unroll
                for (int d = 0; d < ND; ++d) { o[d] += wgt * grad[isub][d]; }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                F.so1n[(e * ND + d) * S + k] = o[d];
                sm[(L::R_U + d * NEQ + e) * TB] = o[d];
            }
            sm[(L::R_U + 3 * NEQ + e) * TB] = un_e;
        }
    }
    // ---- fused calc_solt of the NEXT substep (euler_march.cpp:161-191)
    if (F.fuse_solt)
    {
        asm volatile("" ::: "memory");
        double u[NEQ], g[NEQ][ND], t[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            u[e] = sm[(L::R_U + 3 * NEQ + e) * TB];
#pragma unroll
            for (int d = 0; d < ND; ++d) { g[e][d] = sm[(L::R_U + d * NEQ + e) * TB]; }
        }
        Jaco<ND> J;
        jaco_update<ND>(J, ga_own, u);
        solt_row<ND>(J, g, t);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.so0t_next[e * S + k] = t[e]; }
#if SCG_SW_SHARE
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.sw_next[e * S + k] = u[e] + (F.dt * 0.5) * t[e]; }
#endif
#if SCG_CT_PAIR
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.ct_next[e * S + k] = make_double2(u[e], t[e]); }
#endif
#if SCG_JS_SHARE
        {
            JacoS<ND> Q;
            Q.set(ga_own, u);
            Q.pack(F.js_next, S, k);
        }
#endif
    }
}

// normal_matrix<2>/<3>, euler_bc.cpp:90-128 (host side: evaluated once per boundary face at add_bc time)
inline void normal_matrix_host(int ndim, double const * nv, double * out)
{
    if (2 == ndim)
    {
        out[0] = nv[0];
        out[1] = nv[1];
        out[2] = -nv[1];
        out[3] = nv[0];
        return;
    }
    double axis[3] = {0.0, 0.0, 0.0};
    double const ax = std::fabs(nv[0]), ay = std::fabs(nv[1]), az = std::fabs(nv[2]);
    if (ax <= ay && ax <= az) { axis[0] = 1.0; }
    else if (ay <= az) { axis[1] = 1.0; }
    else { axis[2] = 1.0; }
    double t1[3] = {nv[1] * axis[2] - nv[2] * axis[1], nv[2] * axis[0] - nv[0] * axis[2], nv[0] * axis[1] - nv[1] * axis[0]};
    double const len = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
    t1[0] = t1[0] / len;
    t1[1] = t1[1] / len;
    t1[2] = t1[2] / len;
    double const t2[3] = {nv[1] * t1[2] - nv[2] * t1[1], nv[2] * t1[0] - nv[0] * t1[2], nv[0] * t1[1] - nv[1] * t1[0]};
    for (int d = 0; d < 3; ++d)
    {
        out[d] = nv[d];
        out[3 + d] = t1[d];
        out[6 + d] = t2[d];
    }
}

} // namespace scg
