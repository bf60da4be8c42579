// scg_md.cu -- multidim Euler CESE on the device: the SoA field store, CE geometry setup, the per-phase
// and fused substep kernels, on-device boundary conditions, reductions and the C-ABI of include/scg.h.
//
// Device twin of solvcon::EulerCore (cpp/solvcon/multidim/euler.hpp:56-211).
#include "scg_md_kernels.cuh"
#include "scg_md_pair.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>

using namespace scg;

namespace
{

// ================================================================================================
// CE geometry: EulerCore::prepare_ce_2d / prepare_ce_3d (euler_ce.cpp:78-347), one thread per cell,
// straight from the reference-layout mesh tables (body pointers; ghost rows at negative indices).
// ================================================================================================
struct RawMesh
{
    int32_t const * clfcs; // body pointers
    int32_t const * fccls;
    int32_t const * fcnds;
    double const * ndcrd;
    double const * fccnd;
    double const * clcnd;
    int32_t const * perm;
};

template <int TYPE>
__global__ void __launch_bounds__(128) md_prepare_ce_kernel(RawMesh const M, MdGeo const G, double * solpt, int64_t S)
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NFC = ct_nfc(TYPE);
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.cnt) { return; }
    int64_t const n = G.n;
    int64_t const k = G.beg + i;
    int64_t const icl = M.perm[k];
    double volc = 0.0;
    double cc[ND];
    double ci[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
    {
        cc[d] = 0.0;
        ci[d] = M.clcnd[icl * ND + d];
    }
#pragma unroll 1
    for (int f = 0; f < NFC; ++f)
    {
        int64_t const ifc = M.clfcs[icl * 7 + 1 + f];
        int64_t const jcl = M.fccls[ifc * 4 + 0] + M.fccls[ifc * 4 + 1] - static_cast<int32_t>(icl);
        double ce[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { ce[d] = M.clcnd[jcl * ND + d]; }
        int const q0 = ct_sfoff(TYPE, f);
        if (2 == ND)
        {
            // euler_ce.cpp:92-170
            int64_t const ind0 = M.fcnds[ifc * 5 + 1];
            int64_t const ind1 = M.fcnds[ifc * 5 + 2];
            double const crd10 = M.ndcrd[ind0 * 2 + 0];
            double const crd11 = M.ndcrd[ind0 * 2 + 1];
            double const crd20 = M.ndcrd[ind1 * 2 + 0];
            double const crd21 = M.ndcrd[ind1 * 2 + 1];
            double const cndi0 = (crd10 + crd20 + ci[0]) / 3.0;
            double const cndi1 = (crd11 + crd21 + ci[1]) / 3.0;
            double const voli = fabs((crd10 - ci[0]) * (crd21 - ci[1]) - (crd11 - ci[1]) * (crd20 - ci[0])) / 2.0;
            double const cnde0 = (crd10 + crd20 + ce[0]) / 3.0;
            double const cnde1 = (crd11 + crd21 + ce[1]) / 3.0;
            double const cross = (crd10 - ce[0]) * (crd21 - ce[1]) - (crd11 - ce[1]) * (crd20 - ce[0]);
            double const vole = fabs(cross) / 2.0;
            double const volb = voli + vole;
            G.cevol[(f + 1) * n + i] = volb;
            G.cecnd[(f * 2 + 0) * n + i] = (cndi0 * voli + cnde0 * vole) / volb;
            G.cecnd[(f * 2 + 1) * n + i] = (cndi1 * voli + cnde1 * vole) / volb;
            volc += volb;
            cc[0] += (cndi0 * voli + cnde0 * vole);
            cc[1] += (cndi1 * voli + cnde1 * vole);
            double const sgn = (cross > 0.0) ? 1.0 : -1.0;
            G.sfcnd[((q0 + 0) * 2 + 0) * n + i] = (ce[0] + crd10) / 2.0;
            G.sfcnd[((q0 + 0) * 2 + 1) * n + i] = (ce[1] + crd11) / 2.0;
            double const a_dx = ce[0] - crd10;
            double const a_dy = ce[1] - crd11;
            G.sfnml[((q0 + 0) * 2 + 0) * n + i] = -a_dy * sgn;
            G.sfnml[((q0 + 0) * 2 + 1) * n + i] = a_dx * sgn;
            G.sfcnd[((q0 + 1) * 2 + 0) * n + i] = (crd20 + ce[0]) / 2.0;
            G.sfcnd[((q0 + 1) * 2 + 1) * n + i] = (crd21 + ce[1]) / 2.0;
            double const b_dx = crd20 - ce[0];
            double const b_dy = crd21 - ce[1];
            G.sfnml[((q0 + 1) * 2 + 0) * n + i] = -b_dy * sgn;
            G.sfnml[((q0 + 1) * 2 + 1) * n + i] = b_dx * sgn;
        }
        else
        {
            // euler_ce.cpp:192-340
            constexpr int Z = ND - 1;
            int const fcnnd = ct_fnnd(TYPE, f);
            double fc[3];
            double crd[FCMND + 2][3];
#pragma unroll
            for (int d = 0; d < 3; ++d) { fc[d] = M.fccnd[ifc * 3 + d]; }
            for (int inf = 1; inf <= fcnnd; ++inf)
            {
                int64_t const ind = M.fcnds[ifc * 5 + inf];
                for (int d = 0; d < 3; ++d) { crd[inf][d] = M.ndcrd[ind * 3 + d]; }
            }
            for (int d = 0; d < 3; ++d) { crd[fcnnd + 1][d] = crd[1][d]; }
            double volb = 0.0;
            double bc[3] = {0.0, 0.0, 0.0};
            double const voe = (M.fccls[ifc * 4] == icl) ? 0.5 : -0.5;
            for (int inf = 1; inf <= fcnnd; ++inf)
            {
                double du[3], dv[3], dw[3], wi[3], we[3];
                for (int d = 0; d < 3; ++d)
                {
                    du[d] = crd[inf][d] - fc[d];
                    dv[d] = crd[inf + 1][d] - fc[d];
                }
                dw[0] = du[1] * dv[2] - du[2] * dv[1];
                dw[1] = du[2] * dv[0] - du[0] * dv[2];
                dw[2] = du[0] * dv[1] - du[1] * dv[0];
                for (int d = 0; d < 3; ++d)
                {
                    wi[d] = ci[d < ND ? d : Z] - fc[d];
                    we[d] = ce[d < ND ? d : Z] - fc[d];
                }
                double const voli = fabs(dw[0] * wi[0] + dw[1] * wi[1] + dw[2] * wi[2]) / 6.0;
                double const vole = fabs(dw[0] * we[0] + dw[1] * we[1] + dw[2] * we[2]) / 6.0;
                volb += voli + vole;
                for (int d = 0; d < 3; ++d)
                {
                    double const cndi = (crd[inf][d] + crd[inf + 1][d] + fc[d] + ci[d < ND ? d : Z]) / 4.0;
                    double const cnde = (crd[inf][d] + crd[inf + 1][d] + fc[d] + ce[d < ND ? d : Z]) / 4.0;
                    bc[d] += cndi * voli + cnde * vole;
                }
                int64_t const q = q0 + (inf - 1);
                double su[3], sv[3];
                for (int d = 0; d < 3; ++d)
                {
                    double const cev = ce[d < ND ? d : Z];
                    G.sfcnd[(q * 3 + d) * n + i] = (cev + crd[inf][d] + crd[inf + 1][d]) / 3.0;
                    su[d] = crd[inf][d] - cev;
                    sv[d] = crd[inf + 1][d] - cev;
                }
                G.sfnml[(q * 3 + 0) * n + i] = (su[1] * sv[2] - su[2] * sv[1]) * voe;
                G.sfnml[(q * 3 + 1) * n + i] = (su[2] * sv[0] - su[0] * sv[2]) * voe;
                G.sfnml[(q * 3 + 2) * n + i] = (su[0] * sv[1] - su[1] * sv[0]) * voe;
            }
            G.cevol[(f + 1) * n + i] = volb;
            for (int d = 0; d < ND; ++d)
            {
                G.cecnd[(f * ND + d) * n + i] = bc[d] / volb;
                cc[d] += bc[d];
            }
            volc += volb;
        }
    }
    G.cevol[i] = volc;
#pragma unroll
    for (int d = 0; d < ND; ++d) { solpt[d * S + k] = cc[d] / volc; }
}

// ghost solution points are the (mirrored) mesh cell centroids (euler_march.cpp:231-234)
__global__ void md_ghost_solpt_kernel(double const * clcnd_stored, int ndim, int64_t ngst, int64_t ncell, double * solpt, int64_t S)
{
    int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= ngst) { return; }
    int64_t const row = ngst - 1 - g; // stored row of reference index -(g+1)
    for (int d = 0; d < ndim; ++d) { solpt[d * S + ncell + g] = clcnd_stored[row * ndim + d]; }
}

// ================================================================================================
// phases
// ================================================================================================
template <int ND>
__global__ void __launch_bounds__(256) md_solt_kernel(MdField const F, int32_t row0, int32_t nrow)
{
    int32_t const r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) { return; }
    int64_t const k = row0 + r;
    int64_t const S = F.S;
    double u[ND + 2], g[ND + 2][ND], t[ND + 2];
#pragma unroll
    for (int e = 0; e < ND + 2; ++e)
    {
        u[e] = F.so0c[e * S + k];
#pragma unroll
        for (int d = 0; d < ND; ++d) { g[e][d] = F.so1c[(e * ND + d) * S + k]; }
    }
    Jaco<ND> J;
    jaco_update<ND>(J, F.gamma[k], u);
    solt_row<ND>(J, g, t);
#pragma unroll
    for (int e = 0; e < ND + 2; ++e) { F.so0t[e * S + k] = t[e]; }
#if SCG_SW_SHARE
#pragma unroll
    for (int e = 0; e < ND + 2; ++e) { F.sw[e * S + k] = u[e] + (F.dt * 0.5) * t[e]; }
#endif
#if SCG_CT_PAIR
#pragma unroll
    for (int e = 0; e < ND + 2; ++e) { F.ct[e * S + k] = make_double2(u[e], t[e]); }
#endif
#if SCG_JS_SHARE
    {
        JacoS<ND> Q;
        Q.set(F.gamma[k], u);
        Q.pack(F.js, S, k);
    }
#endif
}

// calc_solt for a list of rows (device indices): the rows the fused path does not cover -- ghost rows without a
// boundary condition, halo copies of a partition
template <int ND>
__global__ void __launch_bounds__(256) md_solt_rows_kernel(MdField const F, int32_t const * rows, int32_t nrow)
{
    int32_t const r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) { return; }
    int64_t const k = rows[r];
    int64_t const S = F.S;
    double u[ND + 2], g[ND + 2][ND], t[ND + 2];
#pragma unroll
    for (int e = 0; e < ND + 2; ++e)
    {
        u[e] = F.so0c[e * S + k];
#pragma unroll
        for (int d = 0; d < ND; ++d) { g[e][d] = F.so1c[(e * ND + d) * S + k]; }
    }
    Jaco<ND> J;
    jaco_update<ND>(J, F.gamma[k], u);
    solt_row<ND>(J, g, t);
#pragma unroll
    for (int e = 0; e < ND + 2; ++e) { F.so0t[e * S + k] = t[e]; }
#if SCG_SW_SHARE
#pragma unroll
    for (int e = 0; e < ND + 2; ++e) { F.sw[e * S + k] = u[e] + (F.dt * 0.5) * t[e]; }
#endif
#if SCG_CT_PAIR
#pragma unroll
    for (int e = 0; e < ND + 2; ++e) { F.ct[e * S + k] = make_double2(u[e], t[e]); }
#endif
#if SCG_JS_SHARE
    {
        JacoS<ND> Q;
        Q.set(F.gamma[k], u);
        Q.pack(F.js, S, k);
    }
#endif
}

template <int TYPE>
__global__ void __launch_bounds__(128) md_soln_kernel(MdField const F, MdGeo const G)
{
    constexpr int NEQ = ct_ndim(TYPE) + 2;
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.nact) { return; }
    double un[NEQ];
    cell_soln<TYPE>(F, G, i, un);
    int64_t const k = G.beg + i;
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { F.so0n[e * F.S + k] = un[e]; }
}

template <int TYPE>
__global__ void __launch_bounds__(128) md_cfl_kernel(MdField const F, MdGeo const G)
{
    constexpr int NEQ = ct_ndim(TYPE) + 2;
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.nact) { return; }
    int64_t const k = G.beg + i;
    double un[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { un[e] = F.so0n[e * F.S + k]; }
    double cflo, cflc;
    cell_cfl<TYPE>(F, G, i, un, cflo, cflc);
    F.cflo[k] = cflo;
    F.cflc[k] = cflc;
    F.so0n[(NEQ - 1) * F.S + k] = un[NEQ - 1];
}

template <int TYPE>
__global__ void __launch_bounds__(128) md_dsoln_kernel(MdField const F, MdGeo const G)
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.nact) { return; }
    int64_t const k = G.beg + i;
    double un[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { un[e] = F.so0n[e * F.S + k]; }
    double gn[NEQ][ND];
    cell_dsoln<TYPE>(F, G, i, un, F.cflc[k], gn);
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
#pragma unroll
        for (int d = 0; d < ND; ++d) { F.so1n[(e * ND + d) * F.S + k] = gn[e][d]; }
    }
}

// Fused substep for body cells: calc_soln -> calc_cfl -> calc_dsoln of one cell in one pass.  The
// phases only couple through the cell's own so0n (SURVEY.md 8a hazard 2), so no grid-wide sync is
// needed.  Ghost cells must see the energy BEFORE calc_cfl's fix-up (hazard 1: bc_soln runs between
// calc_soln and calc_cfl, euler_march.cpp:343-345); cells beside a ghost therefore park that raw
// energy in eraw[cell] for the boundary kernel (one slot per body cell: the writer needs no neighbour index).
template <int TYPE>
__global__ void __launch_bounds__(128) md_substep_kernel(MdField const F, MdGeo const G, double * eraw)
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.nact) { return; }
    int64_t const k = G.beg + i;
    int64_t const S = F.S;
    double un[NEQ];
    cell_soln<TYPE>(F, G, i, un);
    {
        bool beside_ghost = false;
#pragma unroll
        for (int f = 0; f < NFC; ++f) { beside_ghost = beside_ghost || (G.nbr[f * int64_t(G.n) + i] >= F.ncell); }
        if (beside_ghost) { eraw[k] = un[NEQ - 1]; }
    }
    double cflo, cflc;
    cell_cfl<TYPE>(F, G, i, un, cflo, cflc);
    F.cflo[k] = cflo;
    F.cflc[k] = cflc;
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { F.so0n[e * S + k] = un[e]; }
    double gn[NEQ][ND];
    cell_dsoln<TYPE>(F, G, i, un, cflc, gn);
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
#pragma unroll
        for (int d = 0; d < ND; ++d) { F.so1n[(e * ND + d) * S + k] = gn[e][d]; }
    }
}

// Split substep, kernel 1: calc_soln -> (raw energy for the ghosts) -> calc_cfl, one thread per cell.
#ifndef SCG_SOLN_ROWS
#define SCG_SOLN_ROWS 1
#endif
#ifndef SCG_SOLN_MINBLOCKS
#define SCG_SOLN_MINBLOCKS 3
#endif
template <int TYPE>
__global__ void __launch_bounds__(128, SCG_SOLN_MINBLOCKS) md_solncfl_kernel(MdField const F, MdGeo const G, double * eraw)
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    int32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.nact) { return; }
    int64_t const k = G.beg + i;
    double un[NEQ];
#if SCG_SOLN_ROWS
    cell_soln_rows<TYPE>(F, G, i, un);
#else
    cell_soln<TYPE>(F, G, i, un);
#endif
    {
        bool beside_ghost = false;
#pragma unroll
        for (int f = 0; f < NFC; ++f) { beside_ghost = beside_ghost || (G.nbr[f * int64_t(G.n) + i] >= F.ncell); }
        if (beside_ghost) { eraw[k] = un[NEQ - 1]; }
    }
    double cflo, cflc;
    cell_cfl<TYPE>(F, G, i, un, cflo, cflc);
    F.cflo[k] = cflo;
    F.cflc[k] = cflc;
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { F.so0n[e * F.S + k] = un[e]; }
}

// Split substep, kernel 1, staged variant (march mode 3): same arithmetic as md_solncfl_kernel, with the
// cell's geometry double-buffered through shared memory by cp.async and the neighbour gather of the next
// face issued one face ahead (cell_soln_staged); calc_cfl's distance is taken from the staged cecnd rows
// instead of a second pass over them.
#ifndef SCG_STAGED_TB
#define SCG_STAGED_TB 64
#endif
// threads per CTA and resident CTAs per SM the register budget is sized for: the 3D flux evaluation wants ~250
// registers (8 warps per SM; 64-thread CTAs are 3 % faster than 128-thread ones at the same 8 warps), tetrahedra
// run at 204 (10 warps), the 2D types at 128 (16 warps)
#ifndef SCG_STAGED_MB_HEX
#define SCG_STAGED_MB_HEX 4
#endif
#ifndef SCG_STAGED_MB_PRISM
#define SCG_STAGED_MB_PRISM 4
#endif
#ifndef SCG_STAGED_MB_TET
#define SCG_STAGED_MB_TET 5
#endif
#ifndef SCG_STAGED_MB_QUAD
#define SCG_STAGED_MB_QUAD 8
#endif
#ifndef SCG_STAGED_MB_TRI
#define SCG_STAGED_MB_TRI 8
#endif
#ifndef SCG_STAGED_TB_HEX
#define SCG_STAGED_TB_HEX 64
#endif
#ifndef SCG_STAGED_TB_PRISM
#define SCG_STAGED_TB_PRISM 64
#endif
#ifndef SCG_STAGED_TB_TET
#define SCG_STAGED_TB_TET 64
#endif
template <int TYPE>
struct StagedCfg
{
    static constexpr int TB = (TYPE == CT_HEX) ? SCG_STAGED_TB_HEX
                            : (TYPE == CT_PRISM || TYPE == CT_PYRAMID) ? SCG_STAGED_TB_PRISM
                            : (TYPE == CT_TET) ? SCG_STAGED_TB_TET : SCG_STAGED_TB;
    static constexpr int MB = (TYPE == CT_HEX) ? SCG_STAGED_MB_HEX
                            : (TYPE == CT_PRISM || TYPE == CT_PYRAMID) ? SCG_STAGED_MB_PRISM
                            : (TYPE == CT_TET) ? SCG_STAGED_MB_TET
                            : (TYPE == CT_QUAD) ? SCG_STAGED_MB_QUAD : SCG_STAGED_MB_TRI;
};
template <int TYPE>
__global__ void __launch_bounds__(StagedCfg<TYPE>::TB, StagedCfg<TYPE>::MB) md_solncfl_staged_kernel(MdField const F, MdGeo const G, double * eraw)
{
    constexpr int ND = ct_ndim(TYPE);
    constexpr int NEQ = ND + 2;
    constexpr int NFC = ct_nfc(TYPE);
    constexpr int TB = StagedCfg<TYPE>::TB;
    extern __shared__ double stage_smem[];
    constexpr int MODE = soln_stage_mode(TYPE);
    int32_t i;
    double un[NEQ];
    double dist;
    bool dist_ok;
    int32_t jmax;
#if SCG_XCTA_PF
    {
        // rows the block SCG_XCTA_PF further on reads before its first staged group lands: the indices, and face 0's geometry
        int64_t const i2 = int64_t((0 != MODE) ? (G.ibeg & ~31) : G.ibeg) + (int64_t(blockIdx.x) + SCG_XCTA_PF) * TB + threadIdx.x;
        if (i2 < G.iend)
        {
            int64_t const n = G.n, k2 = G.beg + i2;
            int const nsf0 = ct_fnnd(TYPE, 0);
#pragma unroll
            for (int f = 0; f < NFC; ++f) { xcta_pf_row4(G.nbr + f * n, i2); }
#pragma unroll
            for (int d = 0; d < ND; ++d) { xcta_pf_row8(G.cecnd + d * n, i2); }
            xcta_pf_row8(G.cevol, i2);
            xcta_pf_row8(G.cevol + n, i2);
#pragma unroll
            for (int r = 0; r < nsf0 * ND; ++r)
            {
                xcta_pf_row8(G.sfcnd + r * n, i2);
                xcta_pf_row8(G.sfnml + r * n, i2);
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) { xcta_pf_row8(F.solpt + d * F.S, k2); }
            xcta_pf_row8(F.gamma, k2);
        }
    }
#endif
    if constexpr (0 != MODE)
    {
        // Warps start on multiples of 32 cells so that the copies of the geometry rows are aligned; lanes outside
        // [ibeg, iend) evaluate a clamped cell (they take part in the warp's barriers) and store nothing.
        int const lane = threadIdx.x & 31;
        int32_t const i_raw = (G.ibeg & ~31) + blockIdx.x * TB + threadIdx.x;
        int32_t const iw0 = i_raw - lane;
        if (iw0 >= G.iend) { return; } // the whole warp is past the range
        bool const active = i_raw >= G.ibeg && i_raw < G.iend;
        i = i_raw < G.ibeg ? G.ibeg : (i_raw >= G.iend ? G.iend - 1 : i_raw);
        uint64_t * const bars = reinterpret_cast<uint64_t *>(stage_smem + SolnStage<TYPE, TB>::ROWS * TB) + 2 * (threadIdx.x >> 5);
        if constexpr (1 == MODE)
        {
            if (0 == lane)
            {
                mbar_init(&bars[0], 1);
                mbar_init(&bars[1], 1);
                mbar_fence_init();
            }
            __syncwarp();
        }
        cell_soln_staged_bulk<TYPE, TB, MODE>(F, G, i, iw0, lane, stage_smem + threadIdx.x, bars, un, dist, dist_ok, jmax);
        if (!active) { return; }
    }
    else
    {
        i = G.ibeg + blockIdx.x * TB + threadIdx.x;
        if (i >= G.iend) { return; }
        cell_soln_staged<TYPE, TB>(F, G, i, stage_smem + threadIdx.x, un, dist, dist_ok, jmax);
    }
    int64_t const k = G.beg + i;
    if (!dist_ok)
    {
        // a square root of calc_cfl's distance left the fast range: redo the distance with the plain operator
        dist = DBL_MAX;
        for (int f = 0; f < NFC; ++f)
        {
            double d2 = 0.0;
            for (int d = 0; d < ND; ++d)
            {
                double const diff = G.cecnd[(f * ND + d) * int64_t(G.n) + i] - F.solpt[d * F.S + k];
                d2 += diff * diff;
            }
            double const r = sqrt(d2);
            dist = (r < dist) ? r : dist;
        }
    }
    if (jmax >= F.ncell) { eraw[k] = un[NEQ - 1]; } // beside a ghost: park the energy calc_cfl has not fixed yet for md_bc_soln
    double cflo, cflc;
    cfl_from_dist<ND>(F.dt / 2.0, (soln_early(TYPE) ? stage_smem[SolnStage<TYPE, TB>::R_GK * TB + threadIdx.x] : F.gamma[k]), dist, un, cflo, cflc);
    F.cflo[k] = cflo;
    F.cflc[k] = cflc;
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { F.so0n[e * F.S + k] = un[e]; }
}

// Split substep, kernel 2, equation-outer variant (march mode 3): one thread per cell, adj(D)/det(D) of the
// gradient elements parked in shared memory, then one equation at a time (cell_dsoln_eqo).
#ifndef SCG_EQO_TB
#define SCG_EQO_TB 64
#endif
#ifndef SCG_EQO_MB_HEX
#define SCG_EQO_MB_HEX 4
#endif
#ifndef SCG_EQO_MB_PRISM
#define SCG_EQO_MB_PRISM 5
#endif
#ifndef SCG_EQO_MB_TET
#define SCG_EQO_MB_TET 8
#endif
#ifndef SCG_EQO_MB_QUAD
#define SCG_EQO_MB_QUAD 8
#endif
#ifndef SCG_EQO_MB_TRI
#define SCG_EQO_MB_TRI 10
#endif
#ifndef SCG_EQO_PF_HEX
#define SCG_EQO_PF_HEX 1
#endif
#ifndef SCG_EQO_PF_PRISM
#define SCG_EQO_PF_PRISM 0
#endif
#ifndef SCG_EQO_PF_TET
#define SCG_EQO_PF_TET 0
#endif
#ifndef SCG_EQO_PF_2D
#define SCG_EQO_PF_2D 1 // fits the 2D register budgets since the neighbour indices moved to shared memory (quad 126, tri 96 registers):
                        // quad dsoln -7 %, tri -3 % (profiles/r02h_2d_register_prefetch_ab.txt)
#endif
template <int TYPE>
struct EqoCfg
{
    // prefetch of the next equation's neighbour gather: 0 none, 1 at the top of the previous equation into a second register set,
    // 2 in the middle of the previous equation into the same registers (cell_dsoln_eqo)
    static constexpr int PF = (TYPE == CT_HEX) ? SCG_EQO_PF_HEX : (TYPE == CT_PRISM || TYPE == CT_PYRAMID) ? SCG_EQO_PF_PRISM
                            : (TYPE == CT_TET) ? SCG_EQO_PF_TET : SCG_EQO_PF_2D;
    static constexpr int TB = SCG_EQO_TB;
    // resident CTAs per SM the register budget is sized for (shared memory allows at least as many)
    static constexpr int MB = (TYPE == CT_HEX) ? SCG_EQO_MB_HEX
                            : (TYPE == CT_PRISM || TYPE == CT_PYRAMID) ? SCG_EQO_MB_PRISM
                            : (TYPE == CT_TET) ? SCG_EQO_MB_TET
                            : (TYPE == CT_QUAD) ? SCG_EQO_MB_QUAD : SCG_EQO_MB_TRI;
};

template <int TYPE>
__global__ void __launch_bounds__(EqoCfg<TYPE>::TB, EqoCfg<TYPE>::MB) md_dsoln_eqo_kernel(MdField const F, MdGeo const G)
{
    extern __shared__ double eqo_smem[];
    int32_t const i = G.ibeg + blockIdx.x * EqoCfg<TYPE>::TB + threadIdx.x;
#if SCG_XCTA_PF
    {
        constexpr int ND = ct_ndim(TYPE), NFC = ct_nfc(TYPE);
        int64_t const i2 = int64_t(i) + int64_t(SCG_XCTA_PF) * EqoCfg<TYPE>::TB;
        if (i2 < G.iend)
        {
            int64_t const n = G.n, k2 = G.beg + i2;
#pragma unroll
            for (int f = 0; f < NFC; ++f) { xcta_pf_row4(G.nbr + f * n, i2); }
#pragma unroll
            for (int r = 0; r < NFC * ND; ++r) { xcta_pf_row8(G.cecnd + r * n, i2); }
            xcta_pf_row8(F.cflc, k2);
#pragma unroll
            for (int d = 0; d < ND; ++d) { xcta_pf_row8(F.solpt + d * F.S, k2); }
#pragma unroll
            for (int e = 0; e < ND + 2; ++e) { xcta_pf_row8(F.so0n + e * F.S, k2); }
        }
    }
#endif
    if (i >= G.iend) { return; }
    cell_dsoln_eqo<TYPE, EqoCfg<TYPE>::TB, EqoCfg<TYPE>::PF>(F, G, i, eqo_smem + threadIdx.x);
}

// the same phase with two lanes per cell (cell_dsoln_eqo_pair, scg_md_pair.cuh): hexahedra and prisms
#ifndef SCG_PAIR_TB
#define SCG_PAIR_TB 64
#endif
#ifndef SCG_PAIR_MB_HEX
#define SCG_PAIR_MB_HEX 6
#endif
#ifndef SCG_PAIR_MB_PRISM
#define SCG_PAIR_MB_PRISM 6
#endif
#ifndef SCG_PAIR_PF_HEX
#define SCG_PAIR_PF_HEX 1
#endif
#ifndef SCG_PAIR_PF_PRISM
#define SCG_PAIR_PF_PRISM 1
#endif
template <int TYPE>
struct PairCfg
{
    static constexpr bool ON = (TYPE == CT_HEX || TYPE == CT_PRISM) && (0 != (SCG_EQO_PAIR_MASK & (1 << (TYPE - CT_HEX))));
    static constexpr int TYPEP = (TYPE == CT_PRISM) ? CT_PRISM : CT_HEX; // the other types never instantiate the body
    static constexpr int TB = SCG_PAIR_TB;
    static constexpr int MB = (TYPE == CT_PRISM) ? SCG_PAIR_MB_PRISM : SCG_PAIR_MB_HEX;
    static constexpr bool PF = (TYPE == CT_PRISM) ? SCG_PAIR_PF_PRISM : SCG_PAIR_PF_HEX;
};

template <int TYPE>
__global__ void __launch_bounds__(PairCfg<TYPE>::TB, PairCfg<TYPE>::MB) md_dsoln_eqo_pair_kernel(MdField const F, MdGeo const G)
{
    extern __shared__ double eqo_smem[];
    constexpr int TB = PairCfg<TYPE>::TB;
    int const lane = threadIdx.x & 31;
    // a warp marches 16 cells: lanes l and l ^ 16 share cell (l & 15)
    int32_t const i_raw = G.ibeg + (blockIdx.x * (TB / 32) + (threadIdx.x >> 5)) * 16 + (lane & 15);
    if (i_raw - (lane & 15) >= G.iend) { return; } // the whole warp is past the range
    bool const active = i_raw < G.iend;
    cell_dsoln_eqo_pair<TYPE, TB, PairCfg<TYPE>::PF>(F, G, active ? i_raw : G.iend - 1, lane >= 16, active, eqo_smem + threadIdx.x);
}

// second generation of the same phase for the 3D cell types (cell_dsoln_eqo2): every global load before the equation loop
#ifndef SCG_EQO2_MASK
#define SCG_EQO2_MASK 0 // cell types marched by md_dsoln_eqo2_kernel: 1 hex, 2 tet, 4 prism, 8 pyramid
#endif
#ifndef SCG_EQO2_TB
#define SCG_EQO2_TB 64
#endif
#ifndef SCG_EQO2_MB_HEX
#define SCG_EQO2_MB_HEX 6
#endif
#ifndef SCG_EQO2_MB_PRISM
#define SCG_EQO2_MB_PRISM 6
#endif
#ifndef SCG_EQO2_MB_TET
#define SCG_EQO2_MB_TET 8
#endif
template <int TYPE>
struct Eqo2Cfg
{
    static constexpr bool ON = (3 == ct_ndim(TYPE)) && (0 != (SCG_EQO2_MASK & (1 << (TYPE - CT_HEX))));
    static constexpr int TYPE3 = (3 == ct_ndim(TYPE)) ? TYPE : CT_TET; // the 2D types never instantiate the 3D body
    static constexpr int TB = SCG_EQO2_TB;
    static constexpr int MB = (TYPE == CT_HEX) ? SCG_EQO2_MB_HEX : (TYPE == CT_TET) ? SCG_EQO2_MB_TET : SCG_EQO2_MB_PRISM;
};

template <int TYPE>
__global__ void __launch_bounds__(Eqo2Cfg<TYPE>::TB, Eqo2Cfg<TYPE>::MB) md_dsoln_eqo2_kernel(MdField const F, MdGeo const G)
{
    extern __shared__ double eqo_smem[];
    int32_t const i = G.ibeg + blockIdx.x * Eqo2Cfg<TYPE>::TB + threadIdx.x;
    if (i >= G.iend) { return; }
    cell_dsoln_eqo2<TYPE, Eqo2Cfg<TYPE>::TB>(F, G, i, eqo_smem + threadIdx.x);
}

// GradientElementTypeGroup tables for run-time lane -> face lookups (0-based local faces)
__constant__ int8_t c_fge[9][8][3] = {
    {{0}}, {{0}}, {{0}},
    /* 3 quadrilateral */ {{0, 1, 0}, {1, 2, 0}, {2, 3, 0}, {3, 0, 0}},
    /* 4 triangle      */ {{0, 1, 0}, {1, 2, 0}, {2, 0, 0}},
    /* 5 hexahedron    */ {{1, 2, 4}, {5, 2, 1}, {3, 2, 5}, {4, 2, 3}, {4, 0, 1}, {1, 0, 5}, {5, 0, 3}, {3, 0, 4}},
    /* 6 tetrahedron   */ {{2, 0, 1}, {1, 0, 3}, {3, 0, 2}, {1, 3, 2}},
    /* 7 prism         */ {{4, 1, 3}, {2, 1, 4}, {3, 1, 2}, {3, 0, 4}, {4, 0, 2}, {2, 0, 3}},
    /* 8 pyramid       */ {{0, 4, 1}, {1, 4, 2}, {2, 4, 3}, {3, 4, 0}, {0, 2, 3}, {2, 0, 1}},
};

// Split substep, kernel 2: calc_dsoln (euler_gradient.cpp:23-130) with a GROUP of NFGE lanes per cell.
//
// One thread per cell needs ~200 live doubles here (every FGE gradient must exist before the W-3/4
// limiter can run), which caps the SM at 8 warps and unrolls into >150 KB of straight-line code that
// misses the instruction cache.  Instead lane t of a cell's group owns local face t (t < nfc) and
// gradient element t: faces evaluate the displaced points and the Taylor differences, gradient elements
// solve their ndim x ndim system for all equations, and everything that the reference sums in a fixed
// order (centroid, weight totals, the weighted gradient) is parked in shared memory and summed in that
// same order by whichever lanes need it.  Same operations, same order, same rounding -- just spread
// over 3..8 lanes, so registers drop to ~1/3 and the code to ~1/8.
template <int TYPE>
struct Coop
{
    static constexpr int ND = ct_ndim(TYPE);
    static constexpr int NEQ = ND + 2;
    static constexpr int NFC = ct_nfc(TYPE);
    static constexpr int NFGE = ct_nfge(TYPE);
    static constexpr int GS = NFGE;       // lanes per cell
    static constexpr int CPW = 32 / GS;   // cells per warp
    static constexpr int WARPS = 4;
    // shared-memory map of one cell (doubles)
    // row strides are odd so that the lanes of a group (one row each) fall into different banks
    static constexpr int R_FGE = (ND + 1) | 1;              // vob, subcnd*vob
    static constexpr int R_FACE = (ND + NEQ) | 1;           // idis, udf
    static constexpr int R_TERM = NFGE | 1;
    static constexpr int R_W = NFGE | 1;
    static constexpr int O_GP = 0;                          // [NFC][ND]
    static constexpr int O_FGE = O_GP + NFC * ND;           // [NFGE][R_FGE]
    static constexpr int O_FACE = O_FGE + NFGE * R_FGE;     // [NFC][R_FACE]
    static constexpr int O_END1 = O_FACE + NFC * R_FACE;
    static constexpr int O_TERM = 0;                        // [NEQ*ND][R_TERM], aliases the above (dead by then)
    static constexpr int O_END2 = NEQ * ND * R_TERM;
    static constexpr int O_W = (O_END1 > O_END2 ? O_END1 : O_END2); // [NEQ][R_W]
    static constexpr int O_SIG = O_W + NEQ * R_W;           // [NEQ]
    static constexpr int SZ = (O_SIG + NEQ) | 1;            // odd stride: spreads the cells over the banks
};

template <int TYPE>
__global__ void __launch_bounds__(128) md_dsoln_coop_kernel(MdField const F, MdGeo const G)
{
    using C = Coop<TYPE>;
    constexpr int ND = C::ND, NEQ = C::NEQ, NFC = C::NFC, NFGE = C::NFGE, GS = C::GS, CPW = C::CPW;
    __shared__ double smem[C::WARPS * CPW * C::SZ];
    int const lane = threadIdx.x & 31;
    int const warp = threadIdx.x >> 5;
    int const c = lane / GS;
    int const t = lane - c * GS;
    bool const valid = c < CPW;
    int64_t const n = G.n;
    int64_t const S = F.S;
    int64_t icell = (static_cast<int64_t>(blockIdx.x) * C::WARPS + warp) * CPW + (valid ? c : 0);
    bool const live = valid && icell < G.nact;
    if (icell >= G.nact) { icell = G.nact - 1; } // idle lanes shadow a real cell: loads stay in range, nothing is stored
    int64_t const i = icell;
    int64_t const k = G.beg + i;
    double * const sm = smem + (warp * CPW + (valid ? c : 0)) * C::SZ;
    unsigned const FULL = 0xffffffffu;

    double const hdt = F.dt * 0.5;
    double const cflc = F.cflc[k];
    double const acfl = fabs(cflc);
    double const sgm0 = F.sigma0 / acfl;
    double const tau = F.taumin + acfl * F.tauscale;
    double const ofg1 = 1.0 / NFGE;
    bool const is_face = t < NFC;
    int const f = is_face ? t : 0;

    // ---- phase 1: displaced points of the faces (GradientElement.cpp:166-180)
    int64_t const j = G.nbr[f * n + i];
    // the neighbour state is needed in phase 3; start pulling it towards the SM now
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
        prefetch_l1(&F.so0c[e * S + j]);
        prefetch_l1(&F.so0t[e * S + j]);
#pragma unroll
        for (int d = 0; d < ND; ++d) { prefetch_l1(&F.so1c[(e * ND + d) * S + j]); }
    }
    double gp[ND], jd[ND], icecnd[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
    {
        double const mid = G.cecnd[(f * ND + d) * n + i];
        double const jce = F.solpt[d * S + j];
        icecnd[d] = F.solpt[d * S + k];
        gp[d] = mid + tau * (jce - mid);
        jd[d] = gp[d] - jce;
        if (valid && is_face) { sm[C::O_GP + f * ND + d] = gp[d]; }
    }
    __syncwarp(FULL);

    // ---- phase 2: centroid of the gradient element (GradientElement.cpp:81-133)
    double avg[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { avg[d] = 0.0; }
#pragma unroll
    for (int q = 0; q < NFC; ++q)
    {
#pragma unroll
        for (int d = 0; d < ND; ++d) { avg[d] += sm[C::O_GP + q * ND + d]; }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) { avg[d] /= NFC; }
    int fl[ND]; // the faces of this lane's gradient element
#pragma unroll
    for (int v = 0; v < ND; ++v) { fl[v] = c_fge[TYPE][t][v]; }
    {
        double subcnd[ND], dst[ND][ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { subcnd[d] = avg[d]; }
#pragma unroll
        for (int v = 0; v < ND; ++v)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                double const g = sm[C::O_GP + fl[v] * ND + d];
                subcnd[d] += g;
                dst[v][d] = g - avg[d];
            }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) { subcnd[d] /= static_cast<double>(ND + 1); }
        double const vob = det_n<ND>(dst);
        if (valid)
        {
            sm[C::O_FGE + t * C::R_FGE] = vob;
#pragma unroll
            for (int d = 0; d < ND; ++d) { sm[C::O_FGE + t * C::R_FGE + 1 + d] = subcnd[d] * vob; }
        }
    }
    __syncwarp(FULL);
    double cnd[ND];
    {
        double voc = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) { cnd[d] = 0.0; }
#pragma unroll
        for (int q = 0; q < NFGE; ++q)
        {
            voc += sm[C::O_FGE + q * C::R_FGE];
#pragma unroll
            for (int d = 0; d < ND; ++d) { cnd[d] += sm[C::O_FGE + q * C::R_FGE + 1 + d]; }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) { cnd[d] /= voc; }
    }

    // ---- phase 3: per-face displacement and Taylor difference (euler_gradient.cpp:60-72)
    {
        double jdis[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            double const idis = gp[d] - cnd[d];
            jdis[d] = jd[d] + icecnd[d] - cnd[d];
            if (valid && is_face) { sm[C::O_FACE + f * C::R_FACE + d] = idis; }
        }
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double val = F.so0c[e * S + j] + hdt * F.so0t[e * S + j] - F.so0n[e * S + k];
#pragma unroll
            for (int d = 0; d < ND; ++d) { val += jdis[d] * F.so1c[(e * ND + d) * S + j]; }
            if (valid && is_face) { sm[C::O_FACE + f * C::R_FACE + ND + e] = val; }
        }
    }
    __syncwarp(FULL);

    // ---- phase 4: this lane's gradient element: adj(D), det(D), gradients and W-1/2 weights
    double grad[NEQ][ND], w[NEQ];
    {
        double dst[ND][ND], adj[ND][ND];
#pragma unroll
        for (int v = 0; v < ND; ++v)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { dst[v][d] = sm[C::O_FACE + fl[v] * C::R_FACE + d]; }
        }
        adj_n<ND>(dst, adj);
        double const det = det_n<ND>(dst);
        double rr[NEQ][ND];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double u[ND];
#pragma unroll
            for (int b = 0; b < ND; ++b) { u[b] = sm[C::O_FACE + fl[b] * C::R_FACE + ND + e]; }
#pragma unroll
            for (int a = 0; a < ND; ++a)
            {
                double r = 0.0;
#pragma unroll
                for (int b = 0; b < ND; ++b) { r += adj[a][b] * u[b]; }
                rr[e][a] = r;
            }
        }
        // neq*ndim quotients share det: one reciprocal refinement instead of fifteen (see DivBy)
        bool fast_ok = false;
#if SCG_SHARED_RCP
        {
            DivBy<true> by_det(det);
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
#pragma unroll
                for (int a = 0; a < ND; ++a) { grad[e][a] = by_det(rr[e][a]); }
            }
            fast_ok = by_det.ok;
        }
#endif
        if (!fast_ok)
        {
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
#pragma unroll
                for (int a = 0; a < ND; ++a) { grad[e][a] = rr[e][a] / det; }
            }
        }
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double sq = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) { sq += grad[e][d] * grad[e][d]; }
            w[e] = 1.0 / sqrt(sq + MD_ALMOST_ZERO);
            if (valid) { sm[C::O_W + e * C::R_W + t] = w[e]; }
        }
    }
    __syncwarp(FULL);
    // weight totals in FGE order, then this lane's normalised weights (euler_gradient.cpp:90-100)
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
        double wacc = 0.0;
#pragma unroll
        for (int q = 0; q < NFGE; ++q) { wacc += sm[C::O_W + e * C::R_W + q]; }
        w[e] = w[e] / wacc - ofg1;
    }
    __syncwarp(FULL);
    if (valid)
    {
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { sm[C::O_W + e * C::R_W + t] = w[e]; }
    }
    __syncwarp(FULL);

    // ---- phase 5: W-3/4 limiter, one lane per equation (euler_gradient.cpp:101-110)
    for (int e = t; e < NEQ; e += GS)
    {
        double wmax = 0.0, wmin = 0.0;
#pragma unroll
        for (int q = 0; q < NFGE; ++q)
        {
            double const wq = sm[C::O_W + e * C::R_W + q];
            wmax = fmax(wmax, wq);
            wmin = fmin(wmin, wq);
        }
        double const smx = fmin((1.0 - ofg1) / (wmax + MD_ALMOST_ZERO), -ofg1 / (wmin - MD_ALMOST_ZERO));
        if (valid) { sm[C::O_SIG + e] = fmin(smx, sgm0); }
    }
    __syncwarp(FULL);

    // ---- phase 6: weighted gradients of this lane's element (euler_gradient.cpp:112-128)
    if (valid)
    {
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            double const wgt = ofg1 + sm[C::O_SIG + e] * w[e];
#pragma unroll
            for (int d = 0; d < ND; ++d) { sm[C::O_TERM + (e * ND + d) * C::R_TERM + t] = wgt * grad[e][d]; }
        }
    }
    __syncwarp(FULL);

    // ---- phase 7: sum over the elements in FGE order and store so1n
    for (int q = t; q < NEQ * ND; q += GS)
    {
        double o = 0.0;
#pragma unroll
        for (int r = 0; r < NFGE; ++r) { o += sm[C::O_TERM + q * C::R_TERM + r]; }
        if (live) { F.so1n[q * S + k] = o; }
    }
}

// ================================================================================================
// boundary conditions: one thread per (deduplicated) boundary face
// ================================================================================================
struct BcDev
{
    int32_t const * icl; // interior cell (device index)
    int32_t const * jcl; // ghost cell (device index)
    int32_t const * kind;
    double const * mat;  // [ndim*ndim][n] normal matrix
    double const * val;  // [neq][n] conserved inlet state
    int32_t n;
};

template <int ND>
__device__ __forceinline__ void matvec(double const (&m)[ND][ND], double const (&v)[ND], double (&r)[ND])
{
#pragma unroll
    for (int a = 0; a < ND; ++a)
    {
        r[a] = 0.0;
#pragma unroll
        for (int b = 0; b < ND; ++b) { r[a] += m[a][b] * v[b]; }
    }
}

template <int ND>
__device__ __forceinline__ void matmat(double const (&a)[ND][ND], double const (&b)[ND][ND], double (&r)[ND][ND])
{
#pragma unroll
    for (int p = 0; p < ND; ++p)
    {
#pragma unroll
        for (int q = 0; q < ND; ++q)
        {
            r[p][q] = 0.0;
#pragma unroll
            for (int c = 0; c < ND; ++c) { r[p][q] += a[p][c] * b[c][q]; }
        }
    }
}

// bc_soln_impl, euler_bc.cpp:143-215.  USE_ERAW: take the interior energy from eraw[interior cell] (fused march).
template <int ND, bool USE_ERAW>
__global__ void __launch_bounds__(128) md_bc_soln_kernel(MdField const F, BcDev const B, double const * eraw)
{
    constexpr int NEQ = ND + 2;
    int32_t const b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B.n) { return; }
    int64_t const S = F.S;
    int64_t const icl = B.icl[b];
    int64_t const jcl = B.jcl[b];
    int const kind = B.kind[b];
    double ui[NEQ], uj[NEQ];
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { ui[e] = F.so0n[e * S + icl]; }
    if (USE_ERAW) { ui[NEQ - 1] = eraw[icl]; }
    if (SCG_BC_NONREFL == kind)
    {
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { uj[e] = ui[e]; }
    }
    else if (SCG_BC_SLIPWALL == kind)
    {
        double mat[ND][ND], inv[ND][ND], mom[ND], rot[ND], out[ND];
#pragma unroll
        for (int p = 0; p < ND; ++p)
        {
#pragma unroll
            for (int q = 0; q < ND; ++q)
            {
                mat[p][q] = B.mat[(p * ND + q) * int64_t(B.n) + b];
                inv[q][p] = mat[p][q];
            }
        }
        uj[0] = ui[0];
        uj[NEQ - 1] = ui[NEQ - 1];
#pragma unroll
        for (int d = 0; d < ND; ++d) { mom[d] = ui[1 + d]; }
        matvec<ND>(mat, mom, rot);
        rot[0] = -rot[0];
        matvec<ND>(inv, rot, out);
#pragma unroll
        for (int d = 0; d < ND; ++d) { uj[1 + d] = out[d]; }
    }
    else
    {
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { uj[e] = B.val[e * int64_t(B.n) + b]; }
    }
#pragma unroll
    for (int e = 0; e < NEQ; ++e) { F.so0n[e * S + jcl] = uj[e]; }
}

// bc_dsoln_impl, euler_bc.cpp:219-320: reads so1c of the interior cell, writes so1n of the ghost
template <int ND>
__global__ void __launch_bounds__(128) md_bc_dsoln_kernel(MdField const F, BcDev const B)
{
    constexpr int NEQ = ND + 2;
    int32_t const b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B.n) { return; }
    int64_t const S = F.S;
    int64_t const icl = B.icl[b];
    int64_t const jcl = B.jcl[b];
    int const kind = B.kind[b];
    double gi[NEQ][ND], gj[NEQ][ND];
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
#pragma unroll
        for (int d = 0; d < ND; ++d) { gi[e][d] = F.so1c[(e * ND + d) * S + icl]; }
    }
    if (SCG_BC_INLET == kind)
    {
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { gj[e][d] = 0.0; }
        }
    }
    else
    {
        double mat[ND][ND], inv[ND][ND];
#pragma unroll
        for (int p = 0; p < ND; ++p)
        {
#pragma unroll
            for (int q = 0; q < ND; ++q)
            {
                mat[p][q] = B.mat[(p * ND + q) * int64_t(B.n) + b];
                inv[q][p] = mat[p][q];
            }
        }
        if (SCG_BC_NONREFL == kind)
        {
#pragma unroll
            for (int e = 0; e < NEQ; ++e)
            {
                double rot[ND];
                matvec<ND>(mat, gi[e], rot);
                rot[0] = 0.0;
                matvec<ND>(inv, rot, gj[e]);
            }
        }
        else
        {
#pragma unroll
            for (int q = 0; q < 2; ++q)
            {
                int const e = q ? NEQ - 1 : 0;
                double rot[ND];
                matvec<ND>(mat, gi[e], rot);
                rot[0] = -rot[0];
                matvec<ND>(inv, rot, gj[e]);
            }
            double mg[ND][ND], t0[ND][ND], uv[ND][ND], t1[ND][ND], out[ND][ND];
#pragma unroll
            for (int a = 0; a < ND; ++a)
            {
#pragma unroll
                for (int c = 0; c < ND; ++c) { mg[a][c] = gi[1 + a][c]; }
            }
            matmat<ND>(mat, mg, t0);
            matmat<ND>(t0, inv, uv);
#pragma unroll
            for (int it = 1; it < ND; ++it)
            {
                uv[0][it] = -uv[0][it];
                uv[it][0] = -uv[it][0];
            }
            matmat<ND>(inv, uv, t1);
            matmat<ND>(t1, mat, out);
#pragma unroll
            for (int a = 0; a < ND; ++a)
            {
#pragma unroll
                for (int c = 0; c < ND; ++c) { gj[1 + a][c] = out[a][c]; }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < NEQ; ++e)
    {
#pragma unroll
        for (int d = 0; d < ND; ++d) { F.so1n[(e * ND + d) * S + jcl] = gj[e][d]; }
    }
    if (F.fuse_solt)
    {
        // fused calc_solt of the next substep for this ghost row: its so0n was set by bc_soln, its so1n just now
        double u[NEQ], t[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { u[e] = F.so0n[e * S + jcl]; }
        Jaco<ND> J;
        jaco_update<ND>(J, F.gamma[jcl], u);
        solt_row<ND>(J, gj, t);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.so0t_next[e * S + jcl] = t[e]; }
#if SCG_SW_SHARE
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.sw_next[e * S + jcl] = u[e] + (F.dt * 0.5) * t[e]; }
#endif
#if SCG_CT_PAIR
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.ct_next[e * S + jcl] = make_double2(u[e], t[e]); }
#endif
#if SCG_JS_SHARE
        {
            JacoS<ND> Q;
            Q.set(F.gamma[jcl], u);
            Q.pack(F.js_next, S, jcl);
        }
#endif
    }
}

// ================================================================================================
// host layout (reference AoS, ghost rows first, reference cell order) <-> device SoA
// ================================================================================================
template <bool TO_DEVICE>
__global__ void md_rows_kernel(double * stage, int64_t row0, int64_t nrow, int ncomp, int32_t const * rowmap, double * dev, int64_t S)
{
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= nrow * ncomp) { return; }
    int64_t const r = t / ncomp;
    int const c = static_cast<int>(t % ncomp);
    int64_t const k = rowmap[row0 + r];
    if (TO_DEVICE) { dev[c * S + k] = stage[t]; }
    else { stage[t] = dev[c * S + k]; }
}

__device__ __forceinline__ int find_group(GroupTable const & GT, int32_t k)
{
    int g = 0;
    while (g + 1 < GT.ngroup && k >= GT.g[g + 1].beg) { ++g; }
    return g;
}

// which: SCG_A_CEVOL / CECND / SFCND / SFNML.  Host rows: cevol[T][7], cecnd[T][7*ndim] (ghost rows first),
// sfcnd/sfnml[ncell][24][ndim] (body only).
template <bool TO_DEVICE>
__global__ void md_geo_rows_kernel(int which, double * stage, int64_t row0, int64_t nrow, int ncomp, int ndim, int64_t ngst,
                                   int32_t const * rowmap, GroupTable const GT, double * geo, double * solpt, int64_t S)
{
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= nrow * ncomp) { return; }
    int64_t const r = row0 + t / ncomp;
    int const c = static_cast<int>(t % ncomp);
    bool const has_ghost_rows = (SCG_A_CEVOL == which || SCG_A_CECND == which);
    int64_t const stored = has_ghost_rows ? r : r + ngst;
    double * slot = nullptr;
    if (stored >= ngst)
    {
        int32_t const k = rowmap[stored];
        MeshGroup const & g = GT.g[find_group(GT, k)];
        int64_t const i = k - g.beg;
        int64_t const n = g.ld;
        int const nfc = ct_nfc(g.type);
        if (SCG_A_CEVOL == which)
        {
            if (c <= nfc) { slot = geo + g.geo_cevol() + c * n + i; }
        }
        else if (SCG_A_CECND == which)
        {
            int const f = c / ndim, d = c % ndim;
            if (0 == f) { slot = solpt + d * S + k; }
            else if (f <= nfc) { slot = geo + g.geo_cecnd() + ((f - 1) * ndim + d) * n + i; }
        }
        else
        {
            int const d = c % ndim;
            int const fl = (c / ndim) / FCMND, inf = (c / ndim) % FCMND;
            if (fl < nfc && inf < ct_fnnd(g.type, fl))
            {
                int64_t const q = ct_sfoff(g.type, fl) + inf;
                slot = geo + (SCG_A_SFCND == which ? g.geo_sfcnd() : g.geo_sfnml()) + (q * ndim + d) * n + i;
            }
        }
    }
    if (TO_DEVICE)
    {
        if (slot) { *slot = stage[t]; }
    }
    else { stage[t] = slot ? *slot : 0.0; }
}

// Derived body-cell fields of solvcon/multidim/euler/_field.py:56-126 evaluated on the device, one value per
// reference-order cell (the callers evaluate them with numpy after every chunk, which costs a full download of
// the state; this moves 8 B per cell).  Same operations in numpy's order: vel = mom / rho; speed =
// sqrt(sum vel^2); pressure = (gamma - 1) * (E - (0.5 * rho) * speed^2); mach = speed / sqrt(gamma * p / rho).
template <int ND>
__global__ void md_derived_kernel(int field, double const * so0, double const * gamma, double const * cflo, int64_t S,
                                  int32_t const * rowmap_body, int64_t r0, int64_t nr, double * out)
{
    int64_t const q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nr) { return; }
    int64_t const k = rowmap_body[r0 + q];
    double const rho = so0[k];
    double val;
    if (SCG_F_DENSITY == field) { val = rho; }
    else if (SCG_F_TOTAL_ENERGY == field) { val = so0[(ND + 1) * S + k]; }
    else if (SCG_F_CFL == field) { val = cflo[k]; }
    else if (field >= SCG_F_VELX && field <= SCG_F_VELZ) { val = so0[(1 + field - SCG_F_VELX) * S + k] / rho; }
    else
    {
        double sq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            double const v = so0[(1 + d) * S + k] / rho;
            sq = (0 == d) ? v * v : sq + v * v;
        }
        double const speed = sqrt(sq);
        if (SCG_F_SPEED == field) { val = speed; }
        else
        {
            double const ga = gamma[k];
            double const pr = (ga - 1.0) * (so0[(ND + 1) * S + k] - 0.5 * rho * (speed * speed));
            val = (SCG_F_PRESSURE == field) ? pr : speed / sqrt(ga * pr / rho);
        }
    }
    out[q] = val;
}

__global__ void md_fill_kernel(double * p, int64_t n, double v)
{
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < n) { p[t] = v; }
}

// init_solution, euler_solution.cpp:54-99: uniform conserved state into so0n body rows
__global__ void md_init_kernel(double * so0n, int64_t S, int64_t ncell, int neq, double u0, double u1, double u2, double u3, double u4)
{
    int64_t const k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= ncell) { return; }
    double const u[5] = {u0, u1, u2, u3, u4};
    for (int e = 0; e < neq; ++e) { so0n[e * S + k] = u[e]; }
}

// callers' numpy reductions done on the device (solvcon/multidim/euler/_field.py:138-148)
__global__ void __launch_bounds__(256) md_reduce_kernel(double const * cflo, double const * rho, double const * vol, int64_t ncell, double * out)
{
    double mn = DBL_MAX, mx = -DBL_MAX, sm = 0.0;
    int64_t const stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < ncell; k += stride)
    {
        double const c = cflo[k];
        mn = fmin(mn, c);
        mx = fmax(mx, c);
        if (vol) { sm += rho[k] * vol[k]; }
    }
    for (int off = 16; off > 0; off >>= 1)
    {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
        sm += __shfl_down_sync(0xffffffffu, sm, off);
    }
    __shared__ double smn[8], smx[8], ssm[8];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (0 == lane)
    {
        smn[warp] = mn;
        smx[warp] = mx;
        ssm[warp] = sm;
    }
    __syncthreads();
    if (0 == threadIdx.x)
    {
        for (int w = 1; w < 8; ++w)
        {
            mn = fmin(mn, smn[w]);
            mx = fmax(mx, smx[w]);
            sm += ssm[w];
        }
        out[3 * blockIdx.x + 0] = mn;
        out[3 * blockIdx.x + 1] = mx;
        out[3 * blockIdx.x + 2] = sm;
    }
}

// halo exchange support: gather the listed cells' new state / scatter into ghost rows
__global__ void md_halo_pack_kernel(double const * so0, double const * so1, double const * solpt, double const * gamma, int what,
                                    int32_t const * cells, int64_t ncells, int neq, int ndim, int64_t S, double * buf)
{
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= ncells) { return; }
    int64_t const k = cells[t];
    if (0 == what)
    {
        int64_t const w = neq * (1 + ndim);
        for (int e = 0; e < neq; ++e) { buf[t * w + e] = so0[e * S + k]; }
        for (int c = 0; c < neq * ndim; ++c) { buf[t * w + neq + c] = so1[c * S + k]; }
    }
    else
    {
        int64_t const w = ndim + 1;
        for (int d = 0; d < ndim; ++d) { buf[t * w + d] = solpt[d * S + k]; }
        buf[t * w + ndim] = gamma[k];
    }
}

__global__ void md_halo_unpack_kernel(double * so0, double * so1, double * solpt, double * gamma, int what, int32_t const * ghosts,
                                      int64_t nghost, int neq, int ndim, int64_t S, double const * buf)
{
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= nghost) { return; }
    int64_t const k = ghosts[t];
    if (0 == what)
    {
        int64_t const w = neq * (1 + ndim);
        for (int e = 0; e < neq; ++e) { so0[e * S + k] = buf[t * w + e]; }
        for (int c = 0; c < neq * ndim; ++c) { so1[c * S + k] = buf[t * w + neq + c]; }
    }
    else
    {
        int64_t const w = ndim + 1;
        for (int d = 0; d < ndim; ++d) { solpt[d * S + k] = buf[t * w + d]; }
        gamma[k] = buf[t * w + ndim];
    }
}

// bit-exactness self test of DivBy: out_fast[i] = DivBy<true>(b[i])(a[i]) (0/1 flag in ok[i]), out_ref[i] = a[i] / b[i]
__global__ void md_selftest_rsq_kernel(double const * x, double * out_fast, double * out_ref, int * okf, int64_t n)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    bool oks = true, okr = true;
    out_fast[i] = sqrt_fast(x[i], oks);
    out_fast[n + i] = rcp_fast(x[i], okr);
    okf[i] = oks ? 1 : 0;
    okf[n + i] = okr ? 1 : 0;
    out_ref[i] = sqrt(x[i]);
    out_ref[n + i] = 1.0 / x[i];
}

__global__ void md_selftest_div_kernel(double const * a, double const * b, double * out_fast, double * out_ref, int * okf, int64_t n)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    DivBy<true> d(b[i]);
    out_fast[i] = d(a[i]);
    okf[i] = d.ok ? 1 : 0;
    out_ref[i] = a[i] / b[i];
}

struct BcReg
{
    int kind;
    std::vector<int32_t> faces;
    std::vector<double> value;
};

constexpr int64_t STAGE_DOUBLES = int64_t(1) << 23; // 64 MiB staging buffer

} // namespace

struct scg_md
{
    scg_mesh * mesh = nullptr;
    int ndim = 0, neq = 0;
    double dt = 0, sigma0 = 3.0, taumin = 0.0, tauscale = 1.0; // euler.hpp:184-188
    int cur = 0;                                               // so0[cur] is so0c, so0[1-cur] is so0n
    DevBuf<double> so0[2], so1[2], so0t[2], sw[2], ct[2], js[2], gamma, cflo, cflc, solpt, geo, eraw, stage, red, clvol;
    // fused calc_solt (march mode 3): so0t[tcur] belongs to the current so0c/so1c; while a substep runs, its gradient
    // kernels fill so0t[1 - tcur] for the next one.  so0t_valid: so0t[1 - tcur] is complete for every covered row
    // (owned body cells and ghost rows with a boundary condition); solt_rows lists the others.
    int tcur = 0;
    bool so0t_valid = false;
    bool fuse_now = false;
    DevBuf<int32_t> solt_rows;
    int32_t nsolt_rows = 0;
    bool have_clvol = false;
    int fused = 3; // 0: one kernel per reference phase, 1: one fused thread-per-cell kernel, 2: soln+cfl kernel + cooperative dsoln kernel,
                   // 3: cp.async-staged soln+cfl kernel + equation-outer dsoln kernel
    std::vector<BcReg> regs;
    bool bc_dirty = true;
    int32_t nbc = 0;
    DevBuf<int32_t> bc_icl, bc_jcl, bc_kind;
    DevBuf<double> bc_mat, bc_val;
    // halo
    DevBuf<int32_t> halo_send, halo_recv;
    DevBuf<double> halo_buf, halo_rbuf;
    int64_t nsend = 0, nrecv = 0;
    // neighbour ranks of the partition: segment [begin, end) of each peer in the send list and in the recv list
    struct Peer
    {
        int rank;
        int64_t sb, se, rb, re;
    };
    std::vector<Peer> peers;
    cudaEvent_t ev_packed = nullptr, ev_recv = nullptr;
    ~scg_md()
    {
        if (ev_packed) { cudaEventDestroy(ev_packed); }
        if (ev_recv) { cudaEventDestroy(ev_recv); }
    }
};

namespace
{

MdField field_of(scg_md * s)
{
    MdField F;
    F.so0c = s->so0[s->cur].p;
    F.so0n = s->so0[1 - s->cur].p;
    F.so1c = s->so1[s->cur].p;
    F.so1n = s->so1[1 - s->cur].p;
    F.so0t = s->so0t[s->tcur].p;
    F.so0t_next = s->so0t[1 - s->tcur].p;
    F.sw = s->sw[s->tcur].p; // same life cycle as so0t: same producers, same buffer flip
    F.sw_next = s->sw[1 - s->tcur].p;
    F.ct = reinterpret_cast<double2 *>(s->ct[s->tcur].p);
    F.ct_next = reinterpret_cast<double2 *>(s->ct[1 - s->tcur].p);
    F.js = s->js[s->tcur].p;
    F.js_next = s->js[1 - s->tcur].p;
    F.fuse_solt = s->fuse_now ? 1 : 0;
    F.gamma = s->gamma.p;
    F.cflo = s->cflo.p;
    F.cflc = s->cflc.p;
    F.solpt = s->solpt.p;
    F.S = s->mesh->stride;
    F.ncell = static_cast<int32_t>(s->mesh->ncell);
    F.T = static_cast<int32_t>(s->mesh->T);
    F.dt = s->dt;
    F.sigma0 = s->sigma0;
    F.taumin = s->taumin;
    F.tauscale = s->tauscale;
    return F;
}

MdGeo geo_of(scg_md * s, MeshGroup const & g)
{
    MdGeo G;
    G.nbr = s->mesh->d_nbr.p + g.off_nbr;
    G.cevol = s->geo.p + g.geo_cevol();
    G.cecnd = s->geo.p + g.geo_cecnd();
    G.sfcnd = s->geo.p + g.geo_sfcnd();
    G.sfnml = s->geo.p + g.geo_sfnml();
    G.beg = g.beg;
    G.n = g.ld;
    G.cnt = g.n;
    G.nact = g.nact;
    G.ibeg = 0;
    G.iend = g.nact;
    return G;
}

GroupTable table_of(scg_md * s)
{
    GroupTable GT;
    GT.ngroup = static_cast<int>(s->mesh->groups.size());
    for (int q = 0; q < GT.ngroup; ++q) { GT.g[q] = s->mesh->groups[q]; }
    return GT;
}

enum CellPhase
{
    CP_PREPARE,
    CP_SOLN,
    CP_CFL,
    CP_DSOLN,
    CP_SUBSTEP,
    CP_SOLNCFL,
    CP_SOLNCFL_STAGED,
    CP_DSOLN_COOP,
    CP_DSOLN_EQO,
};

// part: 0 all owned cells, 1 the interface cells [0, nif), 2 the interior cells [nif, nact) (mode-3 kernels only)
template <int TYPE>
void launch_cell_phase_t(scg_md * s, MeshGroup const & g, CellPhase ph, int part)
{
    MdField const F = field_of(s);
    MdGeo G = geo_of(s, g);
    bool const all_cells = (CP_PREPARE == ph);
    if (1 == part) { G.iend = g.nif; }
    if (2 == part) { G.ibeg = g.nif; }
    int64_t const ncl = all_cells ? g.n : (G.iend - G.ibeg);
    if (ncl <= 0) { return; }
    unsigned const grid = static_cast<unsigned>(ceil_div(ncl, 128));
    static char const * const phase_name[] = {"prepare", "soln", "cfl", "dsoln", "substep", "solncfl", "solncfl_staged", "dsoln_coop", "dsoln_eqo"};
    char key[48];
    if (timing_on()) { std::snprintf(key, sizeof key, "%s<%d>", phase_name[ph], TYPE); }
    TimeScope const ts(key);
    switch (ph)
    {
    case CP_PREPARE:
    {
        scg_mesh * m = s->mesh;
        RawMesh M;
        M.clfcs = m->d_clfcs.p + m->ngstcell * 7;
        M.fccls = m->d_fccls.p + m->ngstface * 4;
        M.fcnds = m->d_fcnds.p + m->ngstface * 5;
        M.ndcrd = m->d_ndcrd.p + m->ngstnode * m->ndim;
        M.fccnd = m->d_fccnd.p + m->ngstface * m->ndim;
        M.clcnd = m->d_clcnd.p + m->ngstcell * m->ndim;
        M.perm = m->d_perm.p;
        md_prepare_ce_kernel<TYPE><<<grid, 128, 0, stream()>>>(M, G, s->solpt.p, F.S);
        break;
    }
    case CP_SOLN: md_soln_kernel<TYPE><<<grid, 128, 0, stream()>>>(F, G); break;
    case CP_CFL: md_cfl_kernel<TYPE><<<grid, 128, 0, stream()>>>(F, G); break;
    case CP_DSOLN: md_dsoln_kernel<TYPE><<<grid, 128, 0, stream()>>>(F, G); break;
    case CP_SUBSTEP: md_substep_kernel<TYPE><<<grid, 128, 0, stream()>>>(F, G, s->eraw.p); break;
    case CP_SOLNCFL: md_solncfl_kernel<TYPE><<<grid, 128, 0, stream()>>>(F, G, s->eraw.p); break;
    case CP_SOLNCFL_STAGED:
    {
        constexpr int TBS = StagedCfg<TYPE>::TB;
        constexpr int MODE = soln_stage_mode(TYPE);
        // mode 1: + one mbarrier per warp and buffer; modes 1 and 2: warps start on multiples of 32 cells
        constexpr int BYTES = SolnStage<TYPE, TBS>::BYTES + (1 == MODE ? 2 * (TBS / 32) * int(sizeof(uint64_t)) : 0);
        int64_t const span = (0 != MODE) ? G.iend - (G.ibeg & ~31) : ncl;
        static bool attr_set = false;
        if (!attr_set)
        {
            cudaFuncSetAttribute(md_solncfl_staged_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);
            attr_set = true;
        }
        md_solncfl_staged_kernel<TYPE><<<static_cast<unsigned>(ceil_div(span, TBS)), TBS, BYTES, stream()>>>(F, G, s->eraw.p);
        break;
    }
    case CP_DSOLN_EQO:
    {
        if constexpr (PairCfg<TYPE>::ON)
        {
            constexpr int TYPEP = PairCfg<TYPE>::TYPEP;
            constexpr int TBP = PairCfg<TYPE>::TB;
            constexpr int BYTESP = EqoPair<TYPEP>::template bytes<TBP>();
            static bool attrp_set = false;
            if (!attrp_set)
            {
                cudaFuncSetAttribute(md_dsoln_eqo_pair_kernel<TYPEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTESP);
                attrp_set = true;
            }
            md_dsoln_eqo_pair_kernel<TYPEP><<<static_cast<unsigned>(ceil_div(ncl, TBP / 2)), TBP, BYTESP, stream()>>>(F, G);
            break;
        }
        if constexpr (Eqo2Cfg<TYPE>::ON)
        {
            constexpr int TB2 = Eqo2Cfg<TYPE>::TB;
            constexpr int BYTES2 = DsolnEqo2<TYPE, TB2>::BYTES;
            static bool attr2_set = false;
            if (!attr2_set)
            {
                cudaFuncSetAttribute(md_dsoln_eqo2_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES2);
                attr2_set = true;
            }
            md_dsoln_eqo2_kernel<TYPE><<<static_cast<unsigned>(ceil_div(ncl, TB2)), TB2, BYTES2, stream()>>>(F, G);
            break;
        }
        constexpr int TB = EqoCfg<TYPE>::TB;
        constexpr int BYTES = DsolnEqo<TYPE, TB>::BYTES;
        static bool attr_set = false;
        if (!attr_set)
        {
            cudaFuncSetAttribute(md_dsoln_eqo_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);
            attr_set = true;
        }
        md_dsoln_eqo_kernel<TYPE><<<static_cast<unsigned>(ceil_div(ncl, TB)), TB, BYTES, stream()>>>(F, G);
        break;
    }
    case CP_DSOLN_COOP:
    {
        int64_t const cells_per_block = int64_t(Coop<TYPE>::WARPS) * Coop<TYPE>::CPW;
        md_dsoln_coop_kernel<TYPE><<<static_cast<unsigned>(ceil_div(g.nact, cells_per_block)), 128, 0, stream()>>>(F, G);
        break;
    }
    }
    count_launch();
}

void launch_cell_phase(scg_md * s, CellPhase ph, int part = 0)
{
    for (auto const & g : s->mesh->groups)
    {
        switch (g.type)
        {
        case CT_QUAD: launch_cell_phase_t<CT_QUAD>(s, g, ph, part); break;
        case CT_TRI: launch_cell_phase_t<CT_TRI>(s, g, ph, part); break;
        case CT_HEX: launch_cell_phase_t<CT_HEX>(s, g, ph, part); break;
        case CT_TET: launch_cell_phase_t<CT_TET>(s, g, ph, part); break;
        case CT_PRISM: launch_cell_phase_t<CT_PRISM>(s, g, ph, part); break;
        case CT_PYRAMID: launch_cell_phase_t<CT_PYRAMID>(s, g, ph, part); break;
        default: break;
        }
    }
}

void launch_solt(scg_md * s)
{
    MdField const F = field_of(s);
    unsigned const grid = static_cast<unsigned>(ceil_div(F.T, 256));
    TimeScope const ts("solt");
    if (2 == s->ndim) { md_solt_kernel<2><<<grid, 256, 0, stream()>>>(F, 0, F.T); }
    else { md_solt_kernel<3><<<grid, 256, 0, stream()>>>(F, 0, F.T); }
    count_launch();
}

void launch_solt_rows(scg_md * s)
{
    if (0 == s->nsolt_rows) { return; }
    MdField const F = field_of(s);
    unsigned const grid = static_cast<unsigned>(ceil_div(s->nsolt_rows, 256));
    TimeScope const ts("solt_rows");
    if (2 == s->ndim) { md_solt_rows_kernel<2><<<grid, 256, 0, stream()>>>(F, s->solt_rows.p, s->nsolt_rows); }
    else { md_solt_rows_kernel<3><<<grid, 256, 0, stream()>>>(F, s->solt_rows.p, s->nsolt_rows); }
    count_launch();
}

int prepare_ce(scg_md * s)
{
    scg_mesh * m = s->mesh;
    SCG_REQUIRE(m->d_clfcs.p != nullptr, "prepare_ce: the raw mesh tables were released");
    launch_cell_phase(s, CP_PREPARE);
    if (m->ngstcell > 0)
    {
        md_ghost_solpt_kernel<<<static_cast<unsigned>(ceil_div(m->ngstcell, 256)), 256, 0, stream()>>>(
            m->d_clcnd.p, m->ndim, m->ngstcell, m->ncell, s->solpt.p, m->stride);
        count_launch();
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

// Flatten the registered boundaries (registration order, later entries overwrite earlier ones on the
// same face: SURVEY.md 8a hazard 6) into one entry per ghost cell and upload.
int rebuild_bc(scg_md * s)
{
    if (!s->bc_dirty) { return 0; }
    scg_mesh * m = s->mesh;
    int const nd = s->ndim, neq = s->neq;
    std::vector<int32_t> last(static_cast<size_t>(m->nface), -1);
    for (size_t r = 0; r < s->regs.size(); ++r)
    {
        for (int32_t const f : s->regs[r].faces) { last[f] = static_cast<int32_t>(r); }
    }
    std::vector<char> emitted(static_cast<size_t>(m->nface), 0);
    std::vector<int32_t> icl, jcl, kind, faces, reg_of;
    for (size_t r = 0; r < s->regs.size(); ++r)
    {
        for (int32_t const f : s->regs[r].faces)
        {
            if (last[f] != static_cast<int32_t>(r) || emitted[f]) { continue; }
            emitted[f] = 1;
            faces.push_back(f);
            reg_of.push_back(static_cast<int32_t>(r));
            kind.push_back(s->regs[r].kind);
        }
    }
    int64_t const n = static_cast<int64_t>(faces.size());
    // owner / ghost cell and unit normal of the registered faces, from the device copies of fccls / fcnml
    std::vector<int32_t> fc2(static_cast<size_t>(n) * 2);
    std::vector<double> fnml(static_cast<size_t>(n) * nd);
    if (int rc = mesh_fetch_faces(m, faces.data(), n, fc2.data(), fnml.data())) { return rc; }
    icl.resize(static_cast<size_t>(n));
    jcl.resize(static_cast<size_t>(n));
    for (int64_t b = 0; b < n; ++b)
    {
        icl[b] = m->iperm[fc2[2 * b]];
        jcl[b] = static_cast<int32_t>(m->ncell + (-fc2[2 * b + 1] - 1));
    }
    std::vector<double> mat(static_cast<size_t>(n) * nd * nd, 0.0), val(static_cast<size_t>(n) * neq, 0.0);
    for (int64_t b = 0; b < n; ++b)
    {
        double mm[9];
        normal_matrix_host(nd, &fnml[static_cast<size_t>(b) * nd], mm);
        for (int c = 0; c < nd * nd; ++c) { mat[c * n + b] = mm[c]; }
        BcReg const & reg = s->regs[reg_of[b]];
        if (SCG_BC_INLET == reg.kind)
        {
            // euler_bc.cpp:192-210
            double const * v = reg.value.data();
            double const rho = v[0];
            double const p = v[1 + nd];
            double const ga = v[2 + nd];
            double vsq = 0.0;
            for (int d = 0; d < nd; ++d) { vsq += v[1 + d] * v[1 + d]; }
            val[0 * n + b] = rho;
            for (int d = 0; d < nd; ++d) { val[(1 + d) * n + b] = rho * v[1 + d]; }
            val[(neq - 1) * n + b] = p / (ga - 1.0) + 0.5 * rho * vsq;
        }
    }
    s->nbc = static_cast<int32_t>(n);
    if (n > 0)
    {
        SCG_CUDA(s->bc_icl.alloc(n));
        SCG_CUDA(s->bc_jcl.alloc(n));
        SCG_CUDA(s->bc_kind.alloc(n));
        SCG_CUDA(s->bc_mat.alloc(n * nd * nd));
        SCG_CUDA(s->bc_val.alloc(n * neq));
        SCG_CUDA(cudaMemcpyAsync(s->bc_icl.p, icl.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream()));
        SCG_CUDA(cudaMemcpyAsync(s->bc_jcl.p, jcl.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream()));
        SCG_CUDA(cudaMemcpyAsync(s->bc_kind.p, kind.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream()));
        SCG_CUDA(cudaMemcpyAsync(s->bc_mat.p, mat.data(), sizeof(double) * mat.size(), cudaMemcpyHostToDevice, stream()));
        SCG_CUDA(cudaMemcpyAsync(s->bc_val.p, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice, stream()));
        SCG_CUDA(cudaStreamSynchronize(stream()));
    }
    // rows the fused calc_solt does not cover: everything but the owned body cells and the ghost rows written above
    {
        std::vector<char> covered(static_cast<size_t>(m->T), 0);
        for (auto const & g : m->groups)
        {
            for (int32_t i = 0; i < g.nact; ++i) { covered[static_cast<size_t>(g.beg) + i] = 1; }
        }
        for (int32_t const j : jcl) { covered[static_cast<size_t>(j)] = 1; }
        std::vector<int32_t> rows;
        for (int64_t k = 0; k < m->T; ++k)
        {
            if (!covered[static_cast<size_t>(k)]) { rows.push_back(static_cast<int32_t>(k)); }
        }
        s->nsolt_rows = static_cast<int32_t>(rows.size());
        if (!rows.empty())
        {
            SCG_CUDA(s->solt_rows.alloc(static_cast<int64_t>(rows.size())));
            SCG_CUDA(cudaMemcpyAsync(s->solt_rows.p, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, stream()));
            SCG_CUDA(cudaStreamSynchronize(stream()));
        }
    }
    s->so0t_valid = false;
    s->bc_dirty = false;
    return 0;
}

BcDev bc_of(scg_md * s)
{
    BcDev B;
    B.icl = s->bc_icl.p;
    B.jcl = s->bc_jcl.p;
    B.kind = s->bc_kind.p;
    B.mat = s->bc_mat.p;
    B.val = s->bc_val.p;
    B.n = s->nbc;
    return B;
}

int launch_bc_soln(scg_md * s, bool use_eraw)
{
    if (int rc = rebuild_bc(s)) { return rc; }
    if (0 == s->nbc) { return 0; }
    MdField const F = field_of(s);
    BcDev const B = bc_of(s);
    unsigned const grid = static_cast<unsigned>(ceil_div(s->nbc, 128));
    if (2 == s->ndim)
    {
        if (use_eraw) { md_bc_soln_kernel<2, true><<<grid, 128, 0, stream()>>>(F, B, s->eraw.p); }
        else { md_bc_soln_kernel<2, false><<<grid, 128, 0, stream()>>>(F, B, s->eraw.p); }
    }
    else
    {
        if (use_eraw) { md_bc_soln_kernel<3, true><<<grid, 128, 0, stream()>>>(F, B, s->eraw.p); }
        else { md_bc_soln_kernel<3, false><<<grid, 128, 0, stream()>>>(F, B, s->eraw.p); }
    }
    count_launch();
    return 0;
}

int launch_bc_dsoln(scg_md * s)
{
    if (int rc = rebuild_bc(s)) { return rc; }
    if (0 == s->nbc) { return 0; }
    MdField const F = field_of(s);
    BcDev const B = bc_of(s);
    unsigned const grid = static_cast<unsigned>(ceil_div(s->nbc, 128));
    if (2 == s->ndim) { md_bc_dsoln_kernel<2><<<grid, 128, 0, stream()>>>(F, B); }
    else { md_bc_dsoln_kernel<3><<<grid, 128, 0, stream()>>>(F, B); }
    count_launch();
    return 0;
}

// update() and calc_solt of a mode-3 substep.  When the previous substep was a fused one, so0t of every covered row
// was already evaluated by its gradient kernels (into the other so0t buffer): flip to it and evaluate only the rows
// nobody covered.  Otherwise calc_solt runs over all rows.
int begin_fused_substep(scg_md * s)
{
    if (int rc = rebuild_bc(s)) { return rc; } // may invalidate so0t_next (the set of covered rows depends on the BCs)
    s->cur ^= 1; // update(): swap so0c<->so0n, so1c<->so1n (euler.hpp:123-127)
    if (s->so0t_valid)
    {
        s->tcur ^= 1;
        launch_solt_rows(s);
    }
    else { launch_solt(s); }
    s->so0t_valid = false;
    s->fuse_now = true;
    return 0;
}

// march_substep in two stages for a partition of a decomposed mesh (always the mode-3 kernels; every mode is
// bit-identical).  Stage 0: update(), calc_solt on every row, then soln+cfl+dsoln of the INTERFACE cells -- after
// it their so0n/so1n are final and can be packed and sent.  Stage 1: the interior cells and the boundary kernels.
// calc_dsoln of a cell reads so0n of that cell only (SURVEY.md 8a hazard 2), and the boundary kernels read
// interior so0n / old so1c and write ghost rows that the NEXT substep reads, so the reordering changes nothing.
int substep_split(scg_md * s, int stage)
{
    if (0 == stage)
    {
        if (int rc = begin_fused_substep(s)) { return rc; }
        launch_cell_phase(s, CP_SOLNCFL_STAGED, 1);
        launch_cell_phase(s, CP_DSOLN_EQO, 1);
        return 0;
    }
    launch_cell_phase(s, CP_SOLNCFL_STAGED, 2);
    if (int rc = launch_bc_soln(s, true)) { return rc; }
    launch_cell_phase(s, CP_DSOLN_EQO, 2);
    int const rc = launch_bc_dsoln(s);
    s->fuse_now = false;
    s->so0t_valid = (0 == rc);
    return rc;
}

// march_substep, euler_march.cpp:339-348
int substep(scg_md * s)
{
    if (3 == s->fused)
    {
        if (int rc = begin_fused_substep(s)) { return rc; }
        launch_cell_phase(s, CP_SOLNCFL_STAGED);
        if (int rc = launch_bc_soln(s, true)) { return rc; }
        launch_cell_phase(s, CP_DSOLN_EQO);
        int const rc = launch_bc_dsoln(s);
        s->fuse_now = false;
        s->so0t_valid = (0 == rc);
        return rc;
    }
    s->so0t_valid = false;
    s->cur ^= 1; // update(): swap so0c<->so0n, so1c<->so1n (euler.hpp:123-127)
    launch_solt(s);
    if (2 == s->fused)
    {
        launch_cell_phase(s, CP_SOLNCFL);
        if (int rc = launch_bc_soln(s, true)) { return rc; }
        launch_cell_phase(s, CP_DSOLN_COOP);
    }
    else if (1 == s->fused)
    {
        launch_cell_phase(s, CP_SUBSTEP);
        if (int rc = launch_bc_soln(s, true)) { return rc; }
    }
    else
    {
        launch_cell_phase(s, CP_SOLN);
        if (int rc = launch_bc_soln(s, false)) { return rc; }
        launch_cell_phase(s, CP_CFL);
        launch_cell_phase(s, CP_DSOLN);
    }
    return launch_bc_dsoln(s);
}

struct ArrayInfo
{
    int ncomp;
    bool geo;
    bool body_only;
    double * dev;
};

int array_info(scg_md * s, int id, ArrayInfo & a)
{
    int const nd = s->ndim, neq = s->neq;
    a.geo = false;
    a.body_only = false;
    switch (id)
    {
    case SCG_A_CEVOL: a = {CLMFC + 1, true, false, nullptr}; break;
    case SCG_A_CECND: a = {(CLMFC + 1) * nd, true, false, nullptr}; break;
    case SCG_A_SFCND: a = {CLMFC * FCMND * nd, true, true, nullptr}; break;
    case SCG_A_SFNML: a = {CLMFC * FCMND * nd, true, true, nullptr}; break;
    case SCG_A_SO0C: a = {neq, false, false, s->so0[s->cur].p}; break;
    case SCG_A_SO0N: a = {neq, false, false, s->so0[1 - s->cur].p}; break;
    case SCG_A_SO0T: a = {neq, false, false, s->so0t[s->tcur].p}; break;
    case SCG_A_SO1C: a = {neq * nd, false, false, s->so1[s->cur].p}; break;
    case SCG_A_SO1N: a = {neq * nd, false, false, s->so1[1 - s->cur].p}; break;
    case SCG_A_CFLO: a = {1, false, false, s->cflo.p}; break;
    case SCG_A_CFLC: a = {1, false, false, s->cflc.p}; break;
    case SCG_A_GAMMA: a = {1, false, false, s->gamma.p}; break;
    default: return fail("scg_md: unknown array id %d", id);
    }
    return 0;
}

// rows [row_begin, row_end) of the stored array (ghost rows first); row_end < 0 = all rows
template <bool TO_DEVICE>
int transfer(scg_md * s, int id, double * host, int64_t row_begin = 0, int64_t row_end = -1)
{
    ArrayInfo a;
    if (int rc = array_info(s, id, a)) { return rc; }
    scg_mesh * m = s->mesh;
    int64_t const all = a.body_only ? m->ncell : m->T;
    int64_t const nrows = (row_end < 0 || row_end > all) ? all : row_end;
    int64_t const chunk = STAGE_DOUBLES / a.ncomp;
    GroupTable const GT = table_of(s);
    for (int64_t r0 = row_begin < 0 ? 0 : row_begin; r0 < nrows; r0 += chunk)
    {
        int64_t const nr = (nrows - r0 < chunk) ? nrows - r0 : chunk;
        int64_t const cnt = nr * a.ncomp;
        unsigned const grid = static_cast<unsigned>(ceil_div(cnt, 256));
        if (TO_DEVICE)
        {
            SCG_CUDA(cudaMemcpyAsync(s->stage.p, host + r0 * a.ncomp, sizeof(double) * cnt, cudaMemcpyHostToDevice, stream()));
        }
        if (a.geo)
        {
            md_geo_rows_kernel<TO_DEVICE><<<grid, 256, 0, stream()>>>(id, s->stage.p, r0, nr, a.ncomp, s->ndim, m->ngstcell,
                                                                      m->d_rowmap.p, GT, s->geo.p, s->solpt.p, m->stride);
        }
        else
        {
            md_rows_kernel<TO_DEVICE><<<grid, 256, 0, stream()>>>(s->stage.p, r0, nr, a.ncomp, m->d_rowmap.p, a.dev, m->stride);
        }
        count_launch();
        if (!TO_DEVICE)
        {
            SCG_CUDA(cudaMemcpyAsync(host + r0 * a.ncomp, s->stage.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost, stream()));
        }
        // the staging buffer is reused by the next chunk
        SCG_CUDA(cudaStreamSynchronize(stream()));
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

} // namespace

extern "C" {

int scg_md_create(scg_md ** out, scg_mesh * mesh, double time_increment)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_REQUIRE(out && mesh, "scg_md_create: null argument");
    auto * s = new scg_md;
    s->mesh = mesh;
    s->ndim = mesh->ndim;
    s->neq = mesh->ndim + 2;
    s->dt = time_increment;
    int64_t const S = mesh->stride;
    int const nd = s->ndim, neq = s->neq;
    bool ok = true;
    auto z = [&](DevBuf<double> & b, int64_t n) { ok = ok && cudaSuccess == b.alloc_zero(n); };
    for (int q = 0; q < 2; ++q)
    {
        z(s->so0[q], S * neq);
        z(s->so1[q], S * neq * nd);
    }
    z(s->so0t[0], S * neq);
    z(s->so0t[1], S * neq);
    z(s->sw[0], SCG_SW_SHARE ? S * neq : 1);
    z(s->sw[1], SCG_SW_SHARE ? S * neq : 1);
    z(s->ct[0], SCG_CT_PAIR ? S * neq * 2 : 2);
    z(s->ct[1], SCG_CT_PAIR ? S * neq * 2 : 2);
    z(s->js[0], SCG_JS_SHARE ? S * (2 * nd + 3) : 1);
    z(s->js[1], SCG_JS_SHARE ? S * (2 * nd + 3) : 1);
    z(s->gamma, S);
    z(s->cflo, S);
    z(s->cflc, S);
    z(s->solpt, S * nd);
    z(s->geo, mesh->geo_total);
    z(s->eraw, mesh->ncell > 0 ? mesh->ncell : 1); // one slot per body cell, written by the cells beside a ghost
    ok = ok && cudaSuccess == s->stage.alloc(STAGE_DOUBLES) && cudaSuccess == s->red.alloc(3 * 1024);
    if (!ok)
    {
        cudaError_t const e = cudaGetLastError();
        delete s;
        return fail("scg_md_create: device allocation failed: %s", cudaGetErrorString(e));
    }
    mesh->refcount += 1;
    if (int rc = prepare_ce(s))
    {
        mesh->refcount -= 1;
        delete s;
        return rc;
    }
    if (cudaSuccess != cudaStreamSynchronize(stream()))
    {
        mesh->refcount -= 1;
        delete s;
        return fail("scg_md_create: prepare_ce failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = s;
    return 0;
}

int scg_md_destroy(scg_md * s)
{
    if (s)
    {
        cudaStreamSynchronize(stream());
        s->mesh->refcount -= 1;
        delete s;
    }
    return 0;
}

int scg_md_set_params(scg_md * s, double sigma0, double taumin, double tauscale)
{
    SCG_REQUIRE(s, "scg_md_set_params: null solver");
    s->sigma0 = sigma0;
    s->taumin = taumin;
    s->tauscale = tauscale;
    return 0;
}

int scg_md_set_time_increment(scg_md * s, double time_increment)
{
    SCG_REQUIRE(s, "scg_md_set_time_increment: null solver");
    if (time_increment != s->dt) { s->so0t_valid = false; } // the carried so0c + hdt * so0t (MdField::sw) was formed with the old step
    s->dt = time_increment;
    return 0;
}

int scg_md_set_fused(scg_md * s, int fused)
{
    SCG_REQUIRE(s, "scg_md_set_fused: null solver");
    SCG_REQUIRE(fused >= 0 && fused <= 3, "scg_md_set_fused: mode must be 0, 1, 2 or 3");
    s->fused = fused;
    s->so0t_valid = false;
    return 0;
}

int scg_md_init_solution(scg_md * s, double gamma, double rho, const double * v, double p)
{
    SCG_REQUIRE(s && v, "scg_md_init_solution: null argument");
    // same guards and messages as euler_solution.cpp:63-78
    SCG_REQUIRE(gamma > 1.0, "EulerCore::init_solution: gamma %g must be > 1", gamma);
    SCG_REQUIRE(rho > 0.0, "EulerCore::init_solution: rho %g must be > 0", rho);
    SCG_REQUIRE(p >= 0.0, "EulerCore::init_solution: pressure %g must be >= 0", p);
    double vsq = 0.0;
    for (int d = 0; d < s->ndim; ++d) { vsq += v[d] * v[d]; }
    double const energy = p / (gamma - 1.0) + 0.5 * rho * vsq;
    s->so0t_valid = false;
    double u[5] = {0, 0, 0, 0, 0};
    u[0] = rho;
    for (int d = 0; d < s->ndim; ++d) { u[1 + d] = rho * v[d]; }
    u[s->neq - 1] = energy;
    scg_mesh * m = s->mesh;
    md_fill_kernel<<<static_cast<unsigned>(ceil_div(m->stride, 256)), 256, 0, stream()>>>(s->gamma.p, m->stride, gamma);
    md_init_kernel<<<static_cast<unsigned>(ceil_div(m->ncell, 256)), 256, 0, stream()>>>(s->so0[1 - s->cur].p, m->stride, m->ncell,
                                                                                      s->neq, u[0], u[1], u[2], u[3], u[4]);
    count_launch(2);
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_md_add_bc(scg_md * s, int kind, const int32_t * faces, int64_t nface, const double * value, int nvalue)
{
    SCG_REQUIRE(s, "scg_md_add_bc: null solver");
    SCG_REQUIRE(kind >= SCG_BC_NONREFL && kind <= SCG_BC_INLET, "scg_md_add_bc: unknown boundary kind %d", kind);
    scg_mesh * m = s->mesh;
    // same validation and messages as EulerCore::add_bc, euler_bc.cpp:324-389
    for (int64_t q = 0; q < nface; ++q)
    {
        int32_t const ifc = faces[q];
        SCG_REQUIRE(ifc >= 0 && ifc < m->nface, "EulerCore::add_bc: face %d out of range [0, %lld)", ifc,
                    static_cast<long long>(m->nface));
    }
    {
        std::vector<int32_t> fc2(static_cast<size_t>(nface) * 2);
        if (int rc = mesh_fetch_faces(m, faces, nface, fc2.data(), nullptr)) { return rc; }
        for (int64_t q = 0; q < nface; ++q)
        {
            SCG_REQUIRE(fc2[2 * static_cast<size_t>(q) + 1] < 0, "EulerCore::add_bc: face %d is not a boundary face", faces[q]);
        }
    }
    if (SCG_BC_INLET == kind)
    {
        SCG_REQUIRE(nvalue == s->ndim + 3, "EulerCore::add_bc: inlet value size %d must equal ndim+3 = %d", nvalue, s->ndim + 3);
        double const rho = value[0];
        double const p = value[1 + s->ndim];
        double const gamma = value[2 + s->ndim];
        SCG_REQUIRE(gamma > 1.0, "EulerCore::add_bc: inlet gamma %g must be > 1", gamma);
        SCG_REQUIRE(rho > 0.0, "EulerCore::add_bc: inlet rho %g must be > 0", rho);
        SCG_REQUIRE(p >= 0.0, "EulerCore::add_bc: inlet pressure %g must be >= 0", p);
    }
    BcReg reg;
    reg.kind = kind;
    reg.faces.assign(faces, faces + nface);
    if (value && nvalue > 0) { reg.value.assign(value, value + nvalue); }
    s->regs.push_back(std::move(reg));
    s->bc_dirty = true;
    return 0;
}

int scg_md_clear_bc(scg_md * s)
{
    SCG_REQUIRE(s, "scg_md_clear_bc: null solver");
    s->regs.clear();
    s->bc_dirty = true;
    return 0;
}

int scg_md_upload(scg_md * s, int array_id, const double * host)
{
    SCG_REQUIRE(s && host, "scg_md_upload: null argument");
    s->so0t_valid = false;
    return transfer<true>(s, array_id, const_cast<double *>(host));
}

int scg_md_upload_rows(scg_md * s, int array_id, const double * host, int64_t row_begin, int64_t row_end)
{
    SCG_REQUIRE(s && host, "scg_md_upload_rows: null argument");
    s->so0t_valid = false;
    return transfer<true>(s, array_id, const_cast<double *>(host), row_begin, row_end);
}

int scg_md_download(scg_md * s, int array_id, double * host)
{
    SCG_REQUIRE(s && host, "scg_md_download: null argument");
    return transfer<false>(s, array_id, host);
}

int scg_md_phase(scg_md * s, int phase)
{
    SCG_REQUIRE(s, "scg_md_phase: null solver");
    if (phase != SCG_MD_SUBSTEP && phase != SCG_MD_SUBSTEP_IFACE && phase != SCG_MD_SUBSTEP_REST)
    {
        s->so0t_valid = false; // a fine-grained phase changes the state behind the fused calc_solt's back
        s->fuse_now = false;
    }
    switch (phase)
    {
    case SCG_MD_UPDATE: s->cur ^= 1; break;
    case SCG_MD_SOLT: launch_solt(s); break;
    case SCG_MD_SOLN: launch_cell_phase(s, CP_SOLN); break;
    case SCG_MD_BC_SOLN:
        if (int rc = launch_bc_soln(s, false)) { return rc; }
        break;
    case SCG_MD_CFL: launch_cell_phase(s, CP_CFL); break;
    case SCG_MD_DSOLN: launch_cell_phase(s, CP_DSOLN); break;
    case SCG_MD_BC_DSOLN:
        if (int rc = launch_bc_dsoln(s)) { return rc; }
        break;
    case SCG_MD_PREPARE_CE:
        if (int rc = prepare_ce(s)) { return rc; }
        break;
    case SCG_MD_SUBSTEP:
        if (int rc = substep(s)) { return rc; }
        break;
    case SCG_MD_SUBSTEP_IFACE:
        if (int rc = substep_split(s, 0)) { return rc; }
        break;
    case SCG_MD_SUBSTEP_REST:
        if (int rc = substep_split(s, 1)) { return rc; }
        break;
    default: return fail("scg_md_phase: unknown phase %d", phase);
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_md_march(scg_md * s, int64_t steps)
{
    SCG_REQUIRE(s, "scg_md_march: null solver");
    if (steps <= 0) { return 0; } // a non-positive count is a no-op, like the reference's `for (it = 0; it < steps; ++it)`
    for (int64_t it = 0; it < steps; ++it)
    {
        for (int sub = 0; sub < 2; ++sub) // SUBSTEP_RUN, euler.hpp:165
        {
            if (int rc = substep(s)) { return rc; }
        }
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_md_reduce(scg_md * s, const double * clvol_host_or_null, double * cfl_min, double * cfl_max, double * mass)
{
    SCG_REQUIRE(s, "scg_md_reduce: null solver");
    scg_mesh * m = s->mesh;
    if (clvol_host_or_null)
    {
        // clvol arrives in the host layout (ghost rows first); keep a device copy in device order
        if (!s->clvol.p) { SCG_CUDA(s->clvol.alloc_zero(m->stride)); }
        int64_t const chunk = STAGE_DOUBLES;
        for (int64_t r0 = 0; r0 < m->T; r0 += chunk)
        {
            int64_t const nr = (m->T - r0 < chunk) ? m->T - r0 : chunk;
            SCG_CUDA(cudaMemcpyAsync(s->stage.p, clvol_host_or_null + r0, sizeof(double) * nr, cudaMemcpyHostToDevice, stream()));
            md_rows_kernel<true><<<static_cast<unsigned>(ceil_div(nr, 256)), 256, 0, stream()>>>(s->stage.p, r0, nr, 1, m->d_rowmap.p,
                                                                                              s->clvol.p, m->stride);
            count_launch();
            SCG_CUDA(cudaStreamSynchronize(stream()));
        }
        s->have_clvol = true;
    }
    // owned cells only, one launch per cell-type group (halo copies hold another partition's cells)
    int const nblk = 296;
    double mn = DBL_MAX, mx = -DBL_MAX, sm = 0.0;
    std::vector<double> h(3 * nblk);
    for (auto const & g : m->groups)
    {
        if (0 == g.nact) { continue; }
        md_reduce_kernel<<<nblk, 256, 0, stream()>>>(s->cflo.p + g.beg, s->so0[1 - s->cur].p + g.beg,
                                                     s->have_clvol ? s->clvol.p + g.beg : nullptr, g.nact, s->red.p);
        count_launch();
        SCG_CUDA(cudaMemcpyAsync(h.data(), s->red.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, stream()));
        SCG_CUDA(cudaStreamSynchronize(stream()));
        for (int b = 0; b < nblk; ++b)
        {
            mn = h[3 * b] < mn ? h[3 * b] : mn;
            mx = h[3 * b + 1] > mx ? h[3 * b + 1] : mx;
            sm += h[3 * b + 2];
        }
    }
    if (cfl_min) { *cfl_min = mn; }
    if (cfl_max) { *cfl_max = mx; }
    if (mass) { *mass = sm; }
    return 0;
}

int scg_md_derived(scg_md * s, int field, double * out)
{
    SCG_REQUIRE(s && out, "scg_md_derived: null argument");
    SCG_REQUIRE(field >= SCG_F_DENSITY && field <= SCG_F_CFL, "scg_md_derived: unknown field %d", field);
    SCG_REQUIRE(!(SCG_F_VELZ == field && 2 == s->ndim), "scg_md_derived: no z velocity in 2D");
    scg_mesh * m = s->mesh;
    double const * so0n = s->so0[1 - s->cur].p; // march() leaves the newest step in so0n (_field.py:47-54)
    int32_t const * rowmap_body = m->d_rowmap.p + m->ngstcell;
    int64_t const chunk = STAGE_DOUBLES;
    for (int64_t r0 = 0; r0 < m->ncell; r0 += chunk)
    {
        int64_t const nr = (m->ncell - r0 < chunk) ? m->ncell - r0 : chunk;
        unsigned const grid = static_cast<unsigned>(ceil_div(nr, 256));
        if (2 == s->ndim) { md_derived_kernel<2><<<grid, 256, 0, stream()>>>(field, so0n, s->gamma.p, s->cflo.p, m->stride, rowmap_body, r0, nr, s->stage.p); }
        else { md_derived_kernel<3><<<grid, 256, 0, stream()>>>(field, so0n, s->gamma.p, s->cflo.p, m->stride, rowmap_body, r0, nr, s->stage.p); }
        count_launch();
        SCG_CUDA(cudaMemcpyAsync(out + r0, s->stage.p, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream()));
        SCG_CUDA(cudaStreamSynchronize(stream()));
    }
    return 0;
}

int scg_md_algorithmic_bytes(scg_md * s, int64_t * bytes_per_substep)
{
    SCG_REQUIRE(s && bytes_per_substep, "scg_md_algorithmic_bytes: null argument");
    int64_t b = 0;
    for (auto const & g : s->mesh->groups) { b += int64_t(ct_balg(g.type)) * g.nact; }
    *bytes_per_substep = b;
    return 0;
}

int scg_md_normal_matrix(scg_md * s, int32_t ifc, double * out)
{
    SCG_REQUIRE(s && out, "scg_md_normal_matrix: null argument");
    SCG_REQUIRE(ifc >= 0 && ifc < s->mesh->nface, "get_normal_matrix: face %d out of range", ifc);
    double nml[3] = {0.0, 0.0, 0.0};
    if (int rc = mesh_fetch_faces(s->mesh, &ifc, 1, nullptr, nml)) { return rc; }
    normal_matrix_host(s->ndim, nml, out);
    return 0;
}

int scg_selftest_div(const double * a, const double * b, double * out_fast, double * out_ref, int32_t * ok, int64_t n)
{
    if (int rc = ensure_init()) { return rc; }
    DevBuf<double> da, db, df, dr;
    DevBuf<int> dk;
    SCG_CUDA(da.alloc(n));
    SCG_CUDA(db.alloc(n));
    SCG_CUDA(df.alloc(n));
    SCG_CUDA(dr.alloc(n));
    SCG_CUDA(dk.alloc(n));
    SCG_CUDA(cudaMemcpyAsync(da.p, a, sizeof(double) * n, cudaMemcpyHostToDevice, stream()));
    SCG_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * n, cudaMemcpyHostToDevice, stream()));
    md_selftest_div_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, stream()>>>(da.p, db.p, df.p, dr.p, dk.p, n);
    count_launch();
    SCG_CUDA(cudaMemcpyAsync(out_fast, df.p, sizeof(double) * n, cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaMemcpyAsync(out_ref, dr.p, sizeof(double) * n, cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaMemcpyAsync(ok, dk.p, sizeof(int) * n, cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_selftest_rsq(const double * x, double * out_fast, double * out_ref, int32_t * ok, int64_t n)
{
    if (int rc = ensure_init()) { return rc; }
    DevBuf<double> dx, df, dr;
    DevBuf<int> dk;
    SCG_CUDA(dx.alloc(n));
    SCG_CUDA(df.alloc(2 * n));
    SCG_CUDA(dr.alloc(2 * n));
    SCG_CUDA(dk.alloc(2 * n));
    SCG_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, stream()));
    md_selftest_rsq_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, stream()>>>(dx.p, df.p, dr.p, dk.p, n);
    count_launch();
    SCG_CUDA(cudaMemcpyAsync(out_fast, df.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaMemcpyAsync(out_ref, dr.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaMemcpyAsync(ok, dk.p, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_md_halo_set(scg_md * s, const int32_t * send_cells, int64_t nsend, const int32_t * recv_cells, int64_t nrecv)
{
    SCG_REQUIRE(s, "scg_md_halo_set: null solver");
    scg_mesh * m = s->mesh;
    std::vector<int32_t> snd(static_cast<size_t>(nsend)), rcv(static_cast<size_t>(nrecv));
    for (int64_t q = 0; q < nsend; ++q)
    {
        SCG_REQUIRE(send_cells[q] >= 0 && send_cells[q] < m->ncell, "scg_md_halo_set: send cell %d out of range", send_cells[q]);
        snd[q] = m->iperm[send_cells[q]];
    }
    for (int64_t q = 0; q < nrecv; ++q)
    {
        SCG_REQUIRE(recv_cells[q] >= 0 && recv_cells[q] < m->ncell, "scg_md_halo_set: halo cell %d out of range", recv_cells[q]);
        rcv[q] = m->iperm[recv_cells[q]];
    }
    s->nsend = nsend;
    s->nrecv = nrecv;
    int64_t const width = s->neq * (1 + s->ndim);
    SCG_CUDA(s->halo_send.alloc(nsend > 0 ? nsend : 1));
    SCG_CUDA(s->halo_recv.alloc(nrecv > 0 ? nrecv : 1));
    SCG_CUDA(s->halo_buf.alloc((nsend > 0 ? nsend : 1) * width));
    if (nsend) { SCG_CUDA(cudaMemcpyAsync(s->halo_send.p, snd.data(), sizeof(int32_t) * nsend, cudaMemcpyHostToDevice, stream())); }
    if (nrecv) { SCG_CUDA(cudaMemcpyAsync(s->halo_recv.p, rcv.data(), sizeof(int32_t) * nrecv, cudaMemcpyHostToDevice, stream())); }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_md_halo_pack(scg_md * s, int what, double ** dev_buffer, int64_t * ndouble)
{
    SCG_REQUIRE(s && dev_buffer && ndouble, "scg_md_halo_pack: null argument");
    int64_t const width = (0 == what) ? s->neq * (1 + s->ndim) : s->ndim + 1;
    if (s->nsend > 0)
    {
        md_halo_pack_kernel<<<static_cast<unsigned>(ceil_div(s->nsend, 128)), 128, 0, stream()>>>(
            s->so0[1 - s->cur].p, s->so1[1 - s->cur].p, s->solpt.p, s->gamma.p, what, s->halo_send.p, s->nsend, s->neq, s->ndim,
            s->mesh->stride, s->halo_buf.p);
        count_launch();
    }
    *dev_buffer = s->halo_buf.p;
    *ndouble = s->nsend * width;
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_md_halo_unpack(scg_md * s, int what, const double * dev_buffer)
{
    SCG_REQUIRE(s && dev_buffer, "scg_md_halo_unpack: null argument");
    if (s->nrecv > 0)
    {
        md_halo_unpack_kernel<<<static_cast<unsigned>(ceil_div(s->nrecv, 128)), 128, 0, stream()>>>(
            s->so0[1 - s->cur].p, s->so1[1 - s->cur].p, s->solpt.p, s->gamma.p, what, s->halo_recv.p, s->nrecv, s->neq, s->ndim,
            s->mesh->stride, dev_buffer);
        count_launch();
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}


// ---- halo exchange issued by the library itself (scg_comm.cu): stream-ordered, no host synchronisation ------------
// pack on the compute stream -> event -> grouped ncclSend / ncclRecv with every neighbour rank on the communication
// stream -> event -> unpack on the compute stream.  Between begin and end the caller may enqueue the interior march on
// the compute stream: it runs while the messages travel.
static int halo_exchange_begin(scg_md * s, int what)
{
    SCG_REQUIRE(comm_ready(), "scg_md_halo_exchange: no communicator (call scg_comm_init first)");
    int64_t const width = (0 == what) ? s->neq * (1 + s->ndim) : s->ndim + 1;
    if (!s->ev_packed)
    {
        SCG_CUDA(cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
        SCG_CUDA(cudaEventCreateWithFlags(&s->ev_recv, cudaEventDisableTiming));
    }
    double * sendbuf = nullptr;
    int64_t nd = 0;
    if (int rc = scg_md_halo_pack(s, what, &sendbuf, &nd)) { return rc; }
    SCG_CUDA(cudaEventRecord(s->ev_packed, stream()));
    SCG_CUDA(cudaStreamWaitEvent(comm_stream(), s->ev_packed, 0));
    if (int rc = comm_group_start()) { return rc; }
    for (auto const & p : s->peers)
    {
        if (p.se > p.sb)
        {
            if (int rc = comm_send(sendbuf + p.sb * width, (p.se - p.sb) * width, p.rank)) { return rc; }
        }
        if (p.re > p.rb)
        {
            if (int rc = comm_recv(s->halo_rbuf.p + p.rb * width, (p.re - p.rb) * width, p.rank)) { return rc; }
        }
    }
    if (int rc = comm_group_end()) { return rc; }
    SCG_CUDA(cudaEventRecord(s->ev_recv, comm_stream()));
    return 0;
}

static int halo_exchange_end(scg_md * s, int what)
{
    SCG_CUDA(cudaStreamWaitEvent(stream(), s->ev_recv, 0));
    return scg_md_halo_unpack(s, what, s->halo_rbuf.p);
}

int scg_md_halo_peers(scg_md * s, int npeer, const int32_t * peer_rank, const int64_t * send_begin, const int64_t * send_end,
                      const int64_t * recv_begin, const int64_t * recv_end)
{
    SCG_REQUIRE(s && npeer >= 0, "scg_md_halo_peers: bad argument");
    SCG_REQUIRE(0 == npeer || (peer_rank && send_begin && send_end && recv_begin && recv_end), "scg_md_halo_peers: null table");
    std::vector<scg_md::Peer> peers;
    for (int q = 0; q < npeer; ++q)
    {
        SCG_REQUIRE(0 <= send_begin[q] && send_begin[q] <= send_end[q] && send_end[q] <= s->nsend,
                    "scg_md_halo_peers: send segment of peer %d outside the send list", q);
        SCG_REQUIRE(0 <= recv_begin[q] && recv_begin[q] <= recv_end[q] && recv_end[q] <= s->nrecv,
                    "scg_md_halo_peers: recv segment of peer %d outside the recv list", q);
        peers.push_back({peer_rank[q], send_begin[q], send_end[q], recv_begin[q], recv_end[q]});
    }
    s->peers = std::move(peers);
    int64_t const width = s->neq * (1 + s->ndim);
    SCG_CUDA(s->halo_rbuf.alloc_zero((s->nrecv > 0 ? s->nrecv : 1) * width));
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_md_halo_exchange(scg_md * s, int what)
{
    SCG_REQUIRE(s, "scg_md_halo_exchange: null solver");
    if (int rc = halo_exchange_begin(s, what)) { return rc; }
    return halo_exchange_end(s, what);
}

// march of one partition: per substep the interface cells first, their message packed and handed to
// the communication stream, the interior cells and the boundary kernels meanwhile, then the unpack.  Same kernels and
// the same order as the two-stage substep the caller can drive through scg_md_phase; bit-identical to scg_md_march of
// the undecomposed mesh for any number of parts.
int scg_md_march_dist(scg_md * s, int64_t steps)
{
    SCG_REQUIRE(s, "scg_md_march_dist: null solver");
    SCG_REQUIRE(3 == s->fused, "scg_md_march_dist: needs march mode 3");
    if (steps <= 0) { return 0; }
    for (int64_t it = 0; it < 2 * steps; ++it) // SUBSTEP_RUN = 2, euler.hpp:165
    {
        if (int rc = substep_split(s, 0)) { return rc; }
        if (int rc = halo_exchange_begin(s, 0)) { return rc; }
        if (int rc = substep_split(s, 1)) { return rc; }
        if (int rc = halo_exchange_end(s, 0)) { return rc; }
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

} // extern "C"
