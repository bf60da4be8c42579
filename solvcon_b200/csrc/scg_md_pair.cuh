// scg_md_pair.cuh -- calc_dsoln, equation-outer, with TWO LANES PER CELL (hexahedra and prisms).
//
// cell_dsoln_eqo (scg_md_kernels.cuh) keeps one cell per thread: 250 registers for a hexahedron (the gradients and
// weights of its 8 gradient elements, the Taylor differences of 6 faces and the prefetched gather of the next
// equation), hence 2 warps per scheduler and an fp64 pipe that is ~50 % busy (profiles/r02g_*).  Here lanes l and
// l ^ 16 of a warp share a cell: role 0 (lanes 0-15) owns the first half of the gradient elements, role 1 the second
// half, and each lane gathers half of the faces.  Per lane that is half the live state, so 3-4 warps fit a scheduler.
//
// Same operations on the same operands in the same order as cell_dsoln_eqo (euler_gradient.cpp:23-130,
// GradientElement.cpp:81-213); what makes the split exact:
//   * GradientElementTypeGroup's tables (GradientElement.cpp:16-79) are mirror images: element NH + j of a hexahedron
//     or prism uses the faces of element j with the first and the last one exchanged and the other cap face in the
//     middle (checked at compile time, EqoPair::mirrored).  One static instruction stream therefore serves both roles:
//     a lane keeps the values of the faces it gathered (`own`) and its partner's (`oth`, by shuffle), and every
//     operand is a select between two of those registers on the role bit.
//   * every sum the reference takes over ALL gradient elements in a fixed order (centroid and volume of the cell's
//     gradient element group, the weight total, the weighted gradient) is summed by role 0 over its elements from +0,
//     handed over, and continued by role 1 over its elements: the additions of the reference in the reference's order.
//   * minima / maxima are taken per lane and combined (exact and order-independent).
//   * the deferred range tests of the fast divisions / roots are combined over the WARP: when any lane's test fails
//     the whole warp redoes the equation with the plain operators (the fast sequences equal them bit for bit whenever
//     their own test passes), so the shuffles never sit in divergent code.
#pragma once

#include "scg_md_kernels.cuh"

namespace scg
{

#ifndef SCG_EQO_PAIR_MASK
#define SCG_EQO_PAIR_MASK 0 // cell types marched by md_dsoln_eqo_pair_kernel: 1 hexahedra, 4 prisms
#endif

template <int TYPE>
struct EqoPair
{
    static constexpr int ND = 3;
    static constexpr int NFC = ct_nfc(TYPE);
    static constexpr int NFGE = ct_nfge(TYPE);
    static constexpr int NH = NFGE / 2; // gradient elements per lane
    static constexpr int HS = 3;        // faces gathered per lane
    static constexpr int MID0 = fge_face(TYPE, 0, 1);  // the cap face every element of role 0 uses
    static constexpr int MID1 = fge_face(TYPE, NH, 1); // ... of role 1
    // faces a lane gathers; slot 0 is the role's cap face.  A prism has five faces: role 1 gathers face 4 twice and its
    // slot 2 is never referenced (rho / sig return the first match).
    __host__ __device__ static constexpr int own_face(int r, int s)
    {
        constexpr int hx[2][3] = {{2, 1, 4}, {0, 5, 3}};
        constexpr int pr[2][3] = {{1, 2, 3}, {0, 4, 4}};
        return TYPE == CT_HEX ? hx[r][s] : pr[r][s];
    }
    __host__ __device__ static constexpr int rho(int f) // role that gathers face f
    {
        for (int r = 0; r < 2; ++r)
        {
            for (int s = 0; s < HS; ++s)
            {
                if (own_face(r, s) == f) { return r; }
            }
        }
        return -1;
    }
    __host__ __device__ static constexpr int sig(int f) // its slot there
    {
        for (int r = 0; r < 2; ++r)
        {
            for (int s = 0; s < HS; ++s)
            {
                if (own_face(r, s) == f) { return s; }
            }
        }
        return -1;
    }
    __host__ __device__ static constexpr bool mirrored()
    {
        if (2 * NH != NFGE || own_face(0, 0) != MID0 || own_face(1, 0) != MID1) { return false; }
        for (int f = 0; f < NFC; ++f)
        {
            if (rho(f) < 0) { return false; }
        }
        for (int j = 0; j < NH; ++j)
        {
            if (fge_face(TYPE, j, 1) != MID0 || fge_face(TYPE, NH + j, 1) != MID1) { return false; }
            if (fge_face(TYPE, NH + j, 0) != fge_face(TYPE, j, 2) || fge_face(TYPE, NH + j, 2) != fge_face(TYPE, j, 0)) { return false; }
            // the side faces are never a cap face, so slot 0 of the partner is needed by the face average only
            if (fge_face(TYPE, j, 0) == MID0 || fge_face(TYPE, j, 0) == MID1 || fge_face(TYPE, j, 2) == MID0 || fge_face(TYPE, j, 2) == MID1) { return false; }
        }
        return true;
    }
    static constexpr int RPF = ND * ND + 1 + eqo_yrow(TYPE);
    template <int TB>
    static constexpr int bytes() { return NH * RPF * TB * int(sizeof(double)); }
};

// value of face f0 (role 0) / f1 (role 1) out of what this lane gathered itself and what its partner handed over
template <int TYPE>
__device__ __forceinline__ double pair_val(bool r, int f0, int f1, double const (&own)[3], double const (&oth)[3])
{
    using P = EqoPair<TYPE>;
    double const a0 = (0 == P::rho(f0)) ? own[P::sig(f0)] : oth[P::sig(f0)];
    double const a1 = (1 == P::rho(f1)) ? own[P::sig(f1)] : oth[P::sig(f1)];
    return r ? a1 : a0;
}

__device__ __forceinline__ double pair_xchg(double v) { return __shfl_xor_sync(0xffffffffu, v, 16); }

// sum_{all elements, in element order} t, starting from +0 like the reference's accumulators.  Valid on role 1.
template <int NH>
__device__ __forceinline__ double pair_ordered_sum(double const (&t)[NH])
{
    double lo = 0.0;
#pragma unroll
    for (int j = 0; j < NH; ++j) { lo += t[j]; } // role 0: the reference's partial sum over the first NH elements
    double acc = pair_xchg(lo);
#pragma unroll
    for (int j = 0; j < NH; ++j) { acc += t[j]; } // role 1: continued over its own elements
    return acc;
}

// the same, handed back so that both lanes hold the total
template <int NH>
__device__ __forceinline__ double pair_ordered_sum_both(bool r, double const (&t)[NH])
{
    double const acc = pair_ordered_sum<NH>(t);
    double const back = pair_xchg(acc);
    return r ? acc : back;
}

// eqo_solve for the NH elements of one lane.  udo: Taylor differences of the faces this lane gathered, udx: the partner's.
template <int TYPE, int TB, bool FAST>
__device__ __forceinline__ bool eqo_pair_solve(bool r, double const * sadj, double const (&udo)[3], double const (&udx)[3], double ofg1, double sgm0,
                                               double (&grad)[EqoPair<TYPE>::NH][3], double (&w)[EqoPair<TYPE>::NH], double & sigma_max)
{
    using P = EqoPair<TYPE>;
    constexpr int ND = 3;
    constexpr int NH = P::NH;
    constexpr int RPF = P::RPF;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NH; ++j)
    {
        int const fa = fge_face(TYPE, j, 0), fc = fge_face(TYPE, j, 2);
        double ud[ND];
        ud[0] = pair_val<TYPE>(r, fa, fc, udo, udx);
        ud[1] = udo[0];
        ud[2] = pair_val<TYPE>(r, fc, fa, udo, udx);
#pragma unroll
        for (int a = 0; a < ND; ++a)
        {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < ND; ++b)
            {
                double const p = sadj[(j * RPF + a * ND + b) * TB] * ud[b];
                acc = (SCG_SKIP_ZERO_ADD && 0 == b) ? p : acc + p;
            }
            grad[j][a] = acc;
        }
    }
#pragma unroll
    for (int j = 0; j < NH; ++j)
    {
        DivBy<FAST> by_det = (FAST && eqo_yrow(TYPE)) ? DivBy<FAST>(sadj[(j * RPF + ND * ND) * TB], sadj[(j * RPF + RPF - 1) * TB])
                                                      : DivBy<FAST>(sadj[(j * RPF + ND * ND) * TB]);
#pragma unroll
        for (int a = 0; a < ND; ++a) { grad[j][a] = by_det.template quot<SCG_EQO_ZERO>(grad[j][a]); }
        ok = ok && by_det.ok;
    }
#pragma unroll
    for (int j = 0; j < NH; ++j)
    {
        double sq = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) { sq += grad[j][d] * grad[j][d]; }
        w[j] = sq + MD_ALMOST_ZERO;
    }
#pragma unroll
    for (int j = 0; j < NH; ++j) { w[j] = FAST ? sqrt_fast(w[j], ok) : sqrt(w[j]); }
#pragma unroll
    for (int j = 0; j < NH; ++j) { w[j] = FAST ? rcp_fast(w[j], ok) : 1.0 / w[j]; }
    double const wacc = pair_ordered_sum_both<NH>(r, w);
    {
        DivBy<FAST> by_wacc(wacc);
#pragma unroll
        for (int j = 0; j < NH; ++j) { w[j] = by_wacc.template quot<0>(w[j]) - ofg1; }
        ok = ok && by_wacc.ok;
    }
    double wmax = 0.0, wmin = 0.0;
#pragma unroll
    for (int j = 0; j < NH; ++j)
    {
        wmax = fmax(wmax, w[j]);
        wmin = fmin(wmin, w[j]);
    }
    wmax = fmax(wmax, pair_xchg(wmax));
    wmin = fmin(wmin, pair_xchg(wmin));
    DivBy<FAST> by_hi(wmax + MD_ALMOST_ZERO);
    DivBy<FAST> by_lo(wmin - MD_ALMOST_ZERO);
    double const sm = fmin(by_hi.template quot<0>(1.0 - ofg1), by_lo.template quot<0>(-ofg1));
    ok = ok && by_hi.ok && by_lo.ok;
    sigma_max = fmin(sm, sgm0);
    return ok;
}

// what one lane gathers for one equation (euler_gradient.cpp:66-69)
struct EqoPairGather
{
    double un;
    double u[3], g[3][3];
    __device__ __forceinline__ void load(MdField const & F, int32_t const (&rcl)[3], int64_t k, int e)
    {
        int64_t const S = F.S;
        un = F.so0n[e * S + k];
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            int64_t const j = rcl[s];
            u[s] = F.sw[e * S + j];
#pragma unroll
            for (int d = 0; d < 3; ++d) { g[s][d] = F.so1c[(e * 3 + d) * S + j]; }
        }
    }
};

// `i`: the cell (clamped into the launch range for the lanes past its end: they run along and store nothing), `r`: the
// lane's role, `sadj`: this lane's column of the CTA's shared memory [NH * RPF][TB].
template <int TYPE, int TB, bool PF>
__device__ __forceinline__ void cell_dsoln_eqo_pair(MdField const & F, MdGeo const & G, int32_t i, bool r, bool active, double * sadj)
{
    using P = EqoPair<TYPE>;
    static_assert(P::mirrored(), "the gradient element table of this cell type does not split into mirrored halves");
    static_assert(SCG_SW_SHARE, "md_dsoln_eqo_pair gathers MdField::sw");
    constexpr int ND = 3;
    constexpr int NEQ = ND + 2;
    constexpr int NFC = P::NFC;
    constexpr int NFGE = P::NFGE;
    constexpr int NH = P::NH;
    constexpr int RPF = P::RPF;
    int64_t const n = G.n;
    int64_t const S = F.S;
    int64_t const k = G.beg + i;
    double const acfl = fabs(F.cflc[k]);
    double const sgm0 = F.sigma0 / acfl;
    double const tau = F.taumin + acfl * F.tauscale;
    double const ofg1 = 1.0 / NFGE;

    int32_t rcl[3];
    double jdis[3][ND]; // [slot][d]
    EqoPairGather cur;
    {
        // ---- gradient element geometry (GradientElement.cpp:137-213, :81-133)
        double icecnd[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { icecnd[d] = F.solpt[d * S + k]; }
        int fo[3];
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            fo[s] = r ? P::own_face(1, s) : P::own_face(0, s);
            rcl[s] = G.nbr[fo[s] * n + i];
        }
        if (PF) { cur.load(F, rcl, k, 0); }
        double gpo[ND][3], gpx[ND][3]; // displaced points [d][slot]: gathered by this lane / by its partner
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            int64_t const j = rcl[s];
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                double const mid = G.cecnd[(fo[s] * ND + d) * n + i];
                double const jce = F.solpt[d * S + j];
                gpo[d][s] = mid + tau * (jce - mid);
                jdis[s][d] = gpo[d][s] - jce; // completed below: + icecnd - cnd
            }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
#pragma unroll
            for (int s = 0; s < 3; ++s) { gpx[d][s] = pair_xchg(gpo[d][s]); }
        }
        double avg[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { avg[d] = 0.0; }
#pragma unroll
        for (int f = 0; f < NFC; ++f)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d) { avg[d] += pair_val<TYPE>(r, f, f, gpo[d], gpx[d]); }
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
        double tvob[NH], tcnd[ND][NH];
#pragma unroll
        for (int j = 0; j < NH; ++j)
        {
            int const fa = fge_face(TYPE, j, 0), fc = fge_face(TYPE, j, 2);
            double row[ND][ND]; // [vertex][d]
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                row[0][d] = pair_val<TYPE>(r, fa, fc, gpo[d], gpx[d]);
                row[1][d] = gpo[d][0];
                row[2][d] = pair_val<TYPE>(r, fc, fa, gpo[d], gpx[d]);
            }
            double subcnd[ND];
            double dst[ND][ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) { subcnd[d] = avg[d]; }
#pragma unroll
            for (int ivx = 0; ivx < ND; ++ivx)
            {
#pragma unroll
                for (int d = 0; d < ND; ++d)
                {
                    subcnd[d] += row[ivx][d];
                    dst[ivx][d] = row[ivx][d] - avg[d];
                }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) { subcnd[d] /= static_cast<double>(ND + 1); }
            double const vob = det_n<ND>(dst);
            tvob[j] = vob;
#pragma unroll
            for (int d = 0; d < ND; ++d) { tcnd[d][j] = subcnd[d] * vob; }
        }
        double const voc = pair_ordered_sum_both<NH>(r, tvob);
        double cnd[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { cnd[d] = pair_ordered_sum_both<NH>(r, tcnd[d]); }
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
        // ---- displacements (GradientElement.cpp:196-211)
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                jdis[s][d] = jdis[s][d] + icecnd[d] - cnd[d];
                gpo[d][s] = gpo[d][s] - cnd[d];
                gpx[d][s] = gpx[d][s] - cnd[d];
            }
        }
        // ---- adj(D), det(D) of this lane's gradient elements -> shared memory
#pragma unroll
        for (int j = 0; j < NH; ++j)
        {
            int const fa = fge_face(TYPE, j, 0), fc = fge_face(TYPE, j, 2);
            double dst[ND][ND], adj[ND][ND];
#pragma unroll
            for (int d = 0; d < ND; ++d)
            {
                dst[0][d] = pair_val<TYPE>(r, fa, fc, gpo[d], gpx[d]);
                dst[1][d] = gpo[d][0];
                dst[2][d] = pair_val<TYPE>(r, fc, fa, gpo[d], gpx[d]);
            }
            adj_n<ND>(dst, adj);
            double const det = det_n<ND>(dst);
#pragma unroll
            for (int a = 0; a < ND; ++a)
            {
#pragma unroll
                for (int b = 0; b < ND; ++b) { sadj[(j * RPF + a * ND + b) * TB] = adj[a][b]; }
            }
            sadj[(j * RPF + ND * ND) * TB] = det;
            if (eqo_yrow(TYPE)) { sadj[(j * RPF + RPF - 1) * TB] = DivBy<true>(det).y; }
        }
    }

    // ---- one equation at a time
#pragma unroll 1
    for (int e = 0; e < NEQ; ++e)
    {
        asm volatile("" ::: "memory"); // see cell_dsoln_eqo: keeps the adj / det loads inside the loop
        EqoPairGather nxt;
        if (PF) { nxt.load(F, rcl, k, e + 1 < NEQ ? e + 1 : e); }
        else { cur.load(F, rcl, k, e); }
        double udo[3], udx[3];
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            double val = cur.u[s] - cur.un;
#pragma unroll
            for (int d = 0; d < ND; ++d) { val += jdis[s][d] * cur.g[s][d]; }
            udo[s] = val;
        }
        if (PF) { cur = nxt; }
        udx[0] = 0.0; // the partner's cap face: no element of this lane uses it
        udx[1] = pair_xchg(udo[1]);
        udx[2] = pair_xchg(udo[2]);
        double grad[NH][ND], w[NH];
        double sigma_max;
        bool const ok = eqo_pair_solve<TYPE, TB, true>(r, sadj, udo, udx, ofg1, sgm0, grad, w, sigma_max);
        if (!__all_sync(0xffffffffu, ok)) { eqo_pair_solve<TYPE, TB, false>(r, sadj, udo, udx, ofg1, sgm0, grad, w, sigma_max); }
        double term[ND][NH];
#pragma unroll
        for (int j = 0; j < NH; ++j)
        {
            double const wgt = ofg1 + sigma_max * w[j];
#pragma unroll
            for (int d = 0; d < ND; ++d) { term[d][j] = wgt * grad[j][d]; }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d)
        {
            double const o = pair_ordered_sum<NH>(term[d]);
            if (r && active) { F.so1n[(e * ND + d) * S + k] = o; }
        }
    }
    // ---- fused calc_solt of the NEXT substep (euler_march.cpp:161-191), by the lane that stored the gradient
    if (F.fuse_solt && r && active)
    {
        double u[NEQ], g[NEQ][ND], t[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; ++e)
        {
            u[e] = F.so0n[e * S + k];
#pragma unroll
            for (int d = 0; d < ND; ++d) { g[e][d] = __ldcg(&F.so1n[(e * ND + d) * S + k]); }
        }
        Jaco<ND> J;
        jaco_update<ND>(J, F.gamma[k], u);
        solt_row<ND>(J, g, t);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.so0t_next[e * S + k] = t[e]; }
#pragma unroll
        for (int e = 0; e < NEQ; ++e) { F.sw_next[e * S + k] = u[e] + (F.dt * 0.5) * t[e]; }
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

} // namespace scg
