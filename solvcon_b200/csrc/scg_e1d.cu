// scg_e1d.cu -- 1D Euler CESE on the device.
//
// Device twin of onedim::Euler1DCore (cpp/solvcon/onedim/Euler1DCore.hpp:41-349,
// Euler1DCore.cpp:70-172).
//
// Layout.  The reference stores coord[N], cfl[N], gamma[N], so0[N][3], so1[N][3] with the two time
// planes interleaved (even indices / odd indices).  On the device every plane is SoA per component:
//     even plane  e = k/2      (N+1)/2 slots        odd plane  o = (k-1)/2    (N-1)/2 slots
// so a half step streams 9 unit-stride arrays in and 7 out.
//
// Work split.  The reference derives each source solution element once and reuses it for the two
// targets beside it (kernxn = kernxp, Euler1DCore.cpp:115).  Here each lane derives ONE source
// (Jacobian, f, ut, ft, u', both space-time fluxes, the source-point CFL) and takes the right-hand
// neighbour's flux/u' from lane+1 with warp shuffles; a warp therefore owns 32 sources and writes
// 31 targets, adjacent warps overlap by one source.  so0, cfl and so1 of a half step are produced
// by one launch, the NRBC copies (treat_boundary_*) ride on the first half step of march_alpha.
#include "scg_common.cuh"
#include "scg_fastmath.cuh"

#include <cfloat>
#include <vector>

using namespace scg;

namespace
{

struct EPlane
{
    double * x;
    double * ga;
    double * cfl;
    double * u[3];
    double * g[3];
};

struct EPlanes
{
    EPlane p[2];
};

enum : int
{
    F_SO0 = 1,
    F_CFL = 2,
    F_SO1 = 4,
    F_BC = 8,
};

constexpr double E1D_TINY = 1.e-100; // Euler1DCore.hpp:47,182

template <int ALPHA>
__device__ __forceinline__ double pow_alpha(double v)
{
    if (1 == ALPHA) { return v; }
    return v * v;
}

#ifndef SCG_E1D_MINBLOCKS
#define SCG_E1D_MINBLOCKS 3
#endif
#ifndef SCG_E1D_PIN_X
#define SCG_E1D_PIN_X 1
#endif

// Everything a source solution element contributes (Euler1DKernel::derive, Euler1DCore.hpp:210-244, and
// update_cfl at the source point, Euler1DCore.cpp:70-94).  FAST: the ten Jacobian / flux quotients share three
// refined reciprocals (u0 + tiny, u0^2 + tiny, u0^3 + tiny), the CFL part three more, the two square roots run the
// inlined IEEE sequence, and every range test lands in one flag -- one basic block instead of 16 slow-path branches;
// the caller redoes the element with the plain operators when the flag is false (scg_fastmath.cuh).
struct E1dElem
{
    double f1, f2, ut0, ut1, ut2, ft0, ft1, ft2, cfl;
};

template <bool FAST, bool WITH_CFL>
__device__ __forceinline__ bool e1d_derive(double ga, double u0, double u1, double u2, double ux0, double ux1, double ux2,
                                           double hdt, double dmin, E1dElem & o)
{
    double const tiny = E1D_TINY;
    DivBy<FAST> by_a(u0 + tiny);
    DivBy<FAST> by_b(u0 * u0 + tiny);
    DivBy<FAST> by_c(u0 * u0 * u0 + tiny);
    double const j00 = 0.0, j01 = 1.0, j02 = 0.0;
    double const j10 = by_b((ga - 3.0) / 2.0 * u1 * u1);
    double const j11 = by_a(-(ga - 3.0) * u1);
    double const j12 = ga - 1.0;
    double const j20 = by_c((ga - 1.0) * u1 * u1 * u1) - by_b(ga * u1 * u2);
    double const j21 = by_a(ga * u2) - by_b(3.0 / 2.0 * (ga - 1.0) * u1 * u1);
    double const j22 = by_a(ga * u1);

    o.f1 = (ga - 1.0) * u2 + by_a((3.0 - ga) / 2.0 * u1 * u1);
    o.f2 = by_a(ga * u1 * u2) - by_b((ga - 1.0) / 2.0 * u1 * u1 * u1);

    o.ut0 = -j00 * ux0 - j01 * ux1 - j02 * ux2;
    o.ut1 = -j10 * ux0 - j11 * ux1 - j12 * ux2;
    o.ut2 = -j20 * ux0 - j21 * ux1 - j22 * ux2;

    o.ft0 = j00 * o.ut0 + j01 * o.ut1 + j02 * o.ut2;
    o.ft1 = j10 * o.ut0 + j11 * o.ut1 + j12 * o.ut2;
    o.ft2 = j20 * o.ut0 + j21 * o.ut1 + j22 * o.ut2;
    bool ok = by_a.ok && by_b.ok && by_c.ok;
    if (WITH_CFL)
    {
        DivBy<FAST> by_2u0(2.0 * u0);
        DivBy<FAST> by_u0(u0);
        DivBy<FAST> by_dmin(dmin);
        double wspd = u1;
        wspd *= wspd;
        double const ke = by_2u0(wspd);
        double pr = (ga - 1.0) * (u2 - ke);
        pr = (pr + fabs(pr)) / 2.0;
        double const c2 = by_u0(ga * pr);
        double const s1 = FAST ? sqrt_fast_z(c2, ok) : sqrt(c2);
        double const s2 = FAST ? sqrt_fast_z(wspd, ok) : sqrt(wspd);
        wspd = s1 + by_u0(s2);
        o.cfl = by_dmin(hdt * wspd);
        ok = ok && by_2u0.ok && by_u0.ok && by_dmin.ok;
    }
    return ok;
}

// march_half_so0's flux balance (Euler1DCore.cpp:117-128) and march_half_so1_alpha (Euler1DCore.hpp:306-318) of one
// target; n[] holds so0 of the target on exit (and on entry when DO_SO0 is false)
template <int ALPHA, bool FAST, bool DO_SO0, bool DO_SO1>
__device__ __forceinline__ bool e1d_target(double const (&bal)[3], double dx, double dn, double dp, double const (&upn)[3],
                                           double const (&upp)[3], double (&n)[3], double (&g)[3])
{
    bool ok = true;
    if (DO_SO0)
    {
        DivBy<FAST> by_dx(dx);
#pragma unroll
        for (int iv = 0; iv < 3; ++iv) { n[iv] = by_dx(bal[iv]); }
        ok = ok && by_dx.ok;
    }
    if (DO_SO1)
    {
        DivBy<FAST> by_dn(dn);
        DivBy<FAST> by_dp(dp);
#pragma unroll
        for (int iv = 0; iv < 3; ++iv)
        {
            double const duxn = by_dn(n[iv] - upn[iv]);
            double const duxp = by_dp(upp[iv] - n[iv]);
            double const fan = pow_alpha<ALPHA>(fabs(duxn));
            double const fap = pow_alpha<ALPHA>(fabs(duxp));
            DivBy<FAST> by_w(fap + fan + E1D_TINY);
            g[iv] = by_w(fap * duxn + fan * duxp);
            ok = ok && by_w.ok;
        }
        ok = ok && by_dn.ok && by_dp.ok;
    }
    return ok;
}

// S: source plane, T = 1-S.  Source s of plane S has xneg = xT[s+b], xpos = xT[s+b+1] and the target
// between sources s and s+1 is t = s+b+1, with b = -1 for S = even and b = 0 for S = odd.
// Sources run over [sbeg, slast]; targets over t(sbeg) .. t(slast-1); cfl over sources [sbeg, slast-1]
// (update_cfl's loop stops one source short, Euler1DCore.cpp:73-76).
template <int ALPHA, int FLAGS>
__global__ void __launch_bounds__(256, SCG_E1D_MINBLOCKS) e1d_half_kernel(EPlane const ps, EPlane const pt, int S, int64_t sbeg, int64_t slast, double dt)
{
    int const lane = threadIdx.x & 31;
    int64_t const warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int64_t const s_raw = sbeg + warp * 31 + lane;
    if (sbeg + warp * 31 > slast) { return; } // whole warp out of range
    bool const live = s_raw <= slast;
    int64_t const s = live ? s_raw : slast;
    int64_t const b = S ? 0 : -1;

    double const hdt = dt / 2.0; // Euler1DKernel::set_time_increment, Euler1DCore.hpp:186-191
    double const qdt = hdt / 2.0;

    // ---- Euler1DKernel::set_value, Euler1DCore.hpp:194-209
#if SCG_E1D_PIN_X
    // The three coordinates are first needed ~350 instructions into the element; under the 80-register cap ptxas sinks their loads
    // down to that point and the element pays a second exposed DRAM round trip there (ncu: 15 % of the samples on the first
    // consumer).  Their lines are requested into L1 up front (an instruction ptxas keeps in place; no register held), so the sunk
    // loads hit.
    asm volatile("prefetch.global.L1 [%0];" ::"l"(&ps.x[s]));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(&pt.x[s + b]));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(&pt.x[s + b + 1]));
#endif
    double const ga = ps.ga[s];
    double const x = ps.x[s];
    double const xneg = pt.x[s + b];
    double const xpos = pt.x[s + b + 1];
    double const u0 = ps.u[0][s], u1 = ps.u[1][s], u2 = ps.u[2][s];
    double const ux0 = ps.g[0][s], ux1 = ps.g[1][s], ux2 = ps.g[2][s];
    double const xctr = (xpos + xneg) * 0.5;

    // ---- derive + update_cfl of the source (expression order kept)
    constexpr bool WITH_CFL = 0 != (FLAGS & F_CFL);
    double const dxpos = xpos - x;
    double const dxneg = x - xneg;
    double const dmin = dxpos < dxneg ? dxpos : dxneg;
    E1dElem el;
    if (!e1d_derive<true, WITH_CFL>(ga, u0, u1, u2, ux0, ux1, ux2, hdt, dmin, el))
    {
        e1d_derive<false, WITH_CFL>(ga, u0, u1, u2, ux0, ux1, ux2, hdt, dmin, el);
    }
    double const f0 = u1, f1 = el.f1, f2 = el.f2;
    double const ut0 = el.ut0, ut1 = el.ut1, ut2 = el.ut2;
    double const ft0 = el.ft0, ft1 = el.ft1, ft2 = el.ft2;

    double const up0 = u0 + (x - xctr) * ux0 + hdt * ut0;
    double const up1 = u1 + (x - xctr) * ux1 + hdt * ut1;
    double const up2 = u2 + (x - xctr) * ux2 + hdt * ut2;

    if (WITH_CFL && live && lane < 31 && s < slast) { ps.cfl[s] = el.cfl; }

    // ---- NRBC copies riding on the even->odd half step (treat_boundary_so0/so1, Euler1DCore.cpp:132-172):
    // so[1] <- so[2] is odd slot 0 <- even slot 1, so[N-2] <- so[N-3] is odd slot slast <- even slot slast.
    if ((FLAGS & F_BC) && live && 0 == S)
    {
        for (int side = 0; side < 2; ++side)
        {
            if (s != (side ? slast : sbeg)) { continue; }
            int64_t const o = side ? slast : 0;
            pt.u[0][o] = u0;
            pt.u[1][o] = u1;
            pt.u[2][o] = u2;
            pt.g[0][o] = ux0;
            pt.g[1][o] = ux1;
            pt.g[2][o] = ux2;
        }
    }

    if (!(FLAGS & (F_SO0 | F_SO1))) { return; }

    // ---- calc_flux_lr of this source (it is the RIGHT element of the target on its left),
    //      Euler1DCore.hpp:257-266
    double const dxctr = x - xctr;
    double lr0, lr1, lr2;
    {
        double const deltax = x - xneg;
        double const dxmid = 0.5 * (x + xneg) - xctr;
        lr0 = deltax * (u0 + dxmid * ux0) - hdt * (f0 - (dxctr * ut0) + (qdt * ft0));
        lr1 = deltax * (u1 + dxmid * ux1) - hdt * (f1 - (dxctr * ut1) + (qdt * ft1));
        lr2 = deltax * (u2 + dxmid * ux2) - hdt * (f2 - (dxctr * ut2) + (qdt * ft2));
    }
    // right-hand neighbour's quantities
    unsigned const full = 0xffffffffu;
    double const r_lr0 = __shfl_down_sync(full, lr0, 1);
    double const r_lr1 = __shfl_down_sync(full, lr1, 1);
    double const r_lr2 = __shfl_down_sync(full, lr2, 1);
    double const r_up0 = __shfl_down_sync(full, up0, 1);
    double const r_up1 = __shfl_down_sync(full, up1, 1);
    double const r_up2 = __shfl_down_sync(full, up2, 1);
    double const r_x = __shfl_down_sync(full, x, 1);

    if (lane == 31 || s_raw >= slast) { return; } // no target to the right of this source here
    int64_t const t = s + b + 1;

    constexpr bool DO_SO0 = 0 != (FLAGS & F_SO0);
    constexpr bool DO_SO1 = 0 != (FLAGS & F_SO1);
    double bal[3] = {0.0, 0.0, 0.0}, n[3], g[3];
    if (DO_SO0)
    {
        // calc_flux_ll of this source (LEFT element), Euler1DCore.hpp:246-255, and the flux balance
        double const deltax = xpos - x;
        double const dxmid = 0.5 * (x + xpos) - xctr;
        double const ll0 = deltax * (u0 + dxmid * ux0) + hdt * (f0 - (dxctr * ut0) + (qdt * ft0));
        double const ll1 = deltax * (u1 + dxmid * ux1) + hdt * (f1 - (dxctr * ut1) + (qdt * ft1));
        double const ll2 = deltax * (u2 + dxmid * ux2) + hdt * (f2 - (dxctr * ut2) + (qdt * ft2));
        bal[0] = ll0 + r_lr0;
        bal[1] = ll1 + r_lr1;
        bal[2] = ll2 + r_lr2;
    }
    else
    {
        n[0] = pt.u[0][t];
        n[1] = pt.u[1][t];
        n[2] = pt.u[2][t];
    }
    // kernxn = this source, kernxp = lane+1; kernxp.xneg is this source's xpos (the target point itself)
    double const upn[3] = {up0, up1, up2};
    double const upp[3] = {r_up0, r_up1, r_up2};
    double const dx = r_x - x;
    double const dn = xpos - x;
    double const dp = r_x - xpos;
    if (!e1d_target<ALPHA, true, DO_SO0, DO_SO1>(bal, dx, dn, dp, upn, upp, n, g))
    {
        if (!DO_SO0)
        {
            n[0] = pt.u[0][t];
            n[1] = pt.u[1][t];
            n[2] = pt.u[2][t];
        }
        e1d_target<ALPHA, false, DO_SO0, DO_SO1>(bal, dx, dn, dp, upn, upp, n, g);
    }
    if (DO_SO0)
    {
        pt.u[0][t] = n[0];
        pt.u[1][t] = n[1];
        pt.u[2][t] = n[2];
    }
    if (DO_SO1)
    {
        pt.g[0][t] = g[0];
        pt.g[1][t] = g[1];
        pt.g[2][t] = g[2];
    }
}

__global__ void e1d_bc_kernel(EPlanes P, int64_t slast, int which /*0 so0, 1 so1*/)
{
    int const c = threadIdx.x;
    if (blockIdx.x == 0 && c < 3)
    {
        double * const * src = which ? P.p[0].g : P.p[0].u;
        double * const * dst = which ? P.p[1].g : P.p[1].u;
        dst[c][0] = src[c][1];
        dst[c][slast] = src[c][slast];
    }
}

// AoS interleaved host layout <-> SoA planes.  ncomp = 1 or 3.
__global__ void e1d_scatter_kernel(double const * __restrict__ chunk, int64_t k0, int64_t count, int ncomp,
                                   double * e0, double * e1, double * e2, double * o0, double * o1, double * o2)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count * ncomp) { return; }
    int64_t const k = k0 + i / ncomp;
    int const c = static_cast<int>(i % ncomp);
    double const v = chunk[i];
    double * const ev[3] = {e0, e1, e2};
    double * const od[3] = {o0, o1, o2};
    if (k & 1) { od[c][(k - 1) >> 1] = v; }
    else { ev[c][k >> 1] = v; }
}

__global__ void e1d_gather_kernel(double * __restrict__ chunk, int64_t k0, int64_t count, int ncomp,
                                  double const * e0, double const * e1, double const * e2,
                                  double const * o0, double const * o1, double const * o2)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count * ncomp) { return; }
    int64_t const k = k0 + i / ncomp;
    int const c = static_cast<int>(i % ncomp);
    double const * const ev[3] = {e0, e1, e2};
    double const * const od[3] = {o0, o1, o2};
    chunk[i] = (k & 1) ? od[c][(k - 1) >> 1] : ev[c][k >> 1];
}

__global__ void e1d_reduce_kernel(EPlanes P, int64_t ne, int64_t no, double * out)
{
    double mn = DBL_MAX, mx = -DBL_MAX, sm = 0.0;
    int64_t const stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < ne; i += stride)
    {
        double const c = P.p[0].cfl[i];
        mn = fmin(mn, c);
        mx = fmax(mx, c);
        sm += P.p[0].u[0][i];
        if (i < no)
        {
            double const d = P.p[1].cfl[i];
            mn = fmin(mn, d);
            mx = fmax(mx, d);
        }
    }
    for (int off = 16; off > 0; off >>= 1)
    {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
        sm += __shfl_down_sync(0xffffffffu, sm, off);
    }
    __shared__ double smn[32], smx[32], ssm[32];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (0 == lane)
    {
        smn[warp] = mn;
        smx[warp] = mx;
        ssm[warp] = sm;
    }
    __syncthreads();
    if (0 == warp)
    {
        int const nw = blockDim.x >> 5;
        mn = lane < nw ? smn[lane] : DBL_MAX;
        mx = lane < nw ? smx[lane] : -DBL_MAX;
        sm = lane < nw ? ssm[lane] : 0.0;
        for (int off = 16; off > 0; off >>= 1)
        {
            mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
            mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
            sm += __shfl_down_sync(0xffffffffu, sm, off);
        }
        if (0 == lane)
        {
            out[3 * blockIdx.x + 0] = mn;
            out[3 * blockIdx.x + 1] = mx;
            out[3 * blockIdx.x + 2] = sm;
        }
    }
}

constexpr int64_t STAGE_POINTS = int64_t(1) << 23; // points per staging chunk (x3 doubles for so0/so1)

} // namespace

struct scg_e1d
{
    int64_t N = 0;
    int64_t ne = 0, no = 0;
    double dt = 0;
    DevBuf<double> buf[18];
    DevBuf<double> stage;
    DevBuf<double> red;
    EPlanes P{};
    cudaGraphExec_t graph[3] = {nullptr, nullptr, nullptr}; // per alpha: one full step
};

namespace
{

template <int ALPHA>
void launch_half_a(scg_e1d * s, int S, int flags)
{
    // source range of the half step (Euler1DCore.cpp:99-100): even sources e = 1..(N-3)/2,
    // odd sources o = 0..(N-3)/2
    int64_t const sbeg = S ? 0 : 1;
    int64_t const slast = (s->N - 3) / 2;
    if (slast < sbeg) { return; }
    int64_t const nsrc = slast - sbeg + 1;
    int64_t const nwarp = ceil_div(nsrc > 1 ? nsrc - 1 : 1, 31);
    unsigned const grid = static_cast<unsigned>(ceil_div(nwarp * 32, 256));
#define SCG_E1D_CASE(FL)                                                                         \
    case FL: e1d_half_kernel<ALPHA, FL><<<grid, 256, 0, stream()>>>(s->P.p[S], s->P.p[1 - S], S, sbeg, slast, s->dt); break;
    switch (flags)
    {
        SCG_E1D_CASE(F_SO0)
        SCG_E1D_CASE(F_CFL)
        SCG_E1D_CASE(F_SO1)
        SCG_E1D_CASE(F_SO0 | F_CFL | F_SO1)
        SCG_E1D_CASE(F_SO0 | F_CFL | F_SO1 | F_BC)
    default: break;
    }
#undef SCG_E1D_CASE
    count_launch();
}

void launch_half(scg_e1d * s, int S, int flags, int alpha)
{
    if (1 == alpha) { launch_half_a<1>(s, S, flags); }
    else { launch_half_a<2>(s, S, flags); }
}

int copy_in(scg_e1d * s, double const * host, int ncomp, double * const * ev, double * const * od)
{
    for (int64_t k0 = 0; k0 < s->N; k0 += STAGE_POINTS)
    {
        int64_t const cnt = (s->N - k0 < STAGE_POINTS) ? s->N - k0 : STAGE_POINTS;
        SCG_CUDA(cudaMemcpyAsync(s->stage.p, host + k0 * ncomp, sizeof(double) * cnt * ncomp, cudaMemcpyHostToDevice, stream()));
        e1d_scatter_kernel<<<static_cast<unsigned>(ceil_div(cnt * ncomp, 256)), 256, 0, stream()>>>(
            s->stage.p, k0, cnt, ncomp, ev[0], ncomp > 1 ? ev[1] : nullptr, ncomp > 1 ? ev[2] : nullptr,
            od[0], ncomp > 1 ? od[1] : nullptr, ncomp > 1 ? od[2] : nullptr);
        count_launch();
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int copy_out(scg_e1d * s, double * host, int ncomp, double * const * ev, double * const * od)
{
    for (int64_t k0 = 0; k0 < s->N; k0 += STAGE_POINTS)
    {
        int64_t const cnt = (s->N - k0 < STAGE_POINTS) ? s->N - k0 : STAGE_POINTS;
        e1d_gather_kernel<<<static_cast<unsigned>(ceil_div(cnt * ncomp, 256)), 256, 0, stream()>>>(
            s->stage.p, k0, cnt, ncomp, ev[0], ncomp > 1 ? ev[1] : nullptr, ncomp > 1 ? ev[2] : nullptr,
            od[0], ncomp > 1 ? od[1] : nullptr, ncomp > 1 ? od[2] : nullptr);
        count_launch();
        SCG_CUDA(cudaMemcpyAsync(host + k0 * ncomp, s->stage.p, sizeof(double) * cnt * ncomp, cudaMemcpyDeviceToHost, stream()));
    }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    SCG_CUDA(cudaGetLastError());
    return 0;
}

void enqueue_step(scg_e1d * s, int alpha)
{
    // march_alpha body, Euler1DCore.hpp:339-349: half1, NRBC copies, half2
    launch_half(s, 0, F_SO0 | F_CFL | F_SO1 | F_BC, alpha);
    launch_half(s, 1, F_SO0 | F_CFL | F_SO1, alpha);
}

} // namespace

extern "C" {

int scg_e1d_create(scg_e1d ** out, int64_t ncoord, double time_increment)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_REQUIRE(out != nullptr, "scg_e1d_create: null output");
    SCG_REQUIRE(ncoord % 2 != 0, "ncoord cannot be even"); // Euler1DCore.cpp:25-28
    SCG_REQUIRE(ncoord >= 5, "scg_e1d_create: ncoord %lld smaller than 5", static_cast<long long>(ncoord));
    auto * s = new scg_e1d;
    s->N = ncoord;
    s->ne = (ncoord + 1) / 2;
    s->no = (ncoord - 1) / 2;
    s->dt = time_increment;
    for (int plane = 0; plane < 2; ++plane)
    {
        int64_t const len = (plane ? s->no : s->ne) + 8;
        for (int a = 0; a < 9; ++a)
        {
            if (cudaSuccess != s->buf[plane * 9 + a].alloc_zero(len))
            {
                delete s;
                return fail("scg_e1d_create: device allocation failed");
            }
        }
        EPlane & p = s->P.p[plane];
        p.x = s->buf[plane * 9 + 0].p;
        p.ga = s->buf[plane * 9 + 1].p;
        p.cfl = s->buf[plane * 9 + 2].p;
        for (int c = 0; c < 3; ++c)
        {
            p.u[c] = s->buf[plane * 9 + 3 + c].p;
            p.g[c] = s->buf[plane * 9 + 6 + c].p;
        }
    }
    int64_t const stage = 3 * (ncoord < STAGE_POINTS ? ncoord : STAGE_POINTS);
    if (cudaSuccess != s->stage.alloc(stage) || cudaSuccess != s->red.alloc(3 * 1024))
    {
        delete s;
        return fail("scg_e1d_create: device allocation failed");
    }
    // gamma defaults to 1.4 on every point (Euler1DCore.cpp:34)
    std::vector<double> ga(static_cast<size_t>(ncoord), 1.4);
    double * ev[3] = {s->P.p[0].ga, nullptr, nullptr};
    double * od[3] = {s->P.p[1].ga, nullptr, nullptr};
    if (int rc = copy_in(s, ga.data(), 1, ev, od))
    {
        delete s;
        return rc;
    }
    cudaStreamSynchronize(stream());
    *out = s;
    return 0;
}

int scg_e1d_destroy(scg_e1d * s)
{
    if (s)
    {
        cudaStreamSynchronize(stream());
        for (auto & g : s->graph)
        {
            if (g) { cudaGraphExecDestroy(g); }
        }
        delete s;
    }
    return 0;
}

int scg_e1d_upload(scg_e1d * s, const double * coord, const double * gamma, const double * so0, const double * so1,
                   const double * cfl)
{
    EPlane & e = s->P.p[0];
    EPlane & o = s->P.p[1];
    if (coord)
    {
        double * ev[3] = {e.x, nullptr, nullptr};
        double * od[3] = {o.x, nullptr, nullptr};
        if (int rc = copy_in(s, coord, 1, ev, od)) { return rc; }
    }
    if (gamma)
    {
        double * ev[3] = {e.ga, nullptr, nullptr};
        double * od[3] = {o.ga, nullptr, nullptr};
        if (int rc = copy_in(s, gamma, 1, ev, od)) { return rc; }
    }
    if (cfl)
    {
        double * ev[3] = {e.cfl, nullptr, nullptr};
        double * od[3] = {o.cfl, nullptr, nullptr};
        if (int rc = copy_in(s, cfl, 1, ev, od)) { return rc; }
    }
    if (so0) { if (int rc = copy_in(s, so0, 3, e.u, o.u)) { return rc; } }
    if (so1) { if (int rc = copy_in(s, so1, 3, e.g, o.g)) { return rc; } }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_e1d_download(scg_e1d * s, double * so0, double * so1, double * cfl)
{
    EPlane & e = s->P.p[0];
    EPlane & o = s->P.p[1];
    if (so0) { if (int rc = copy_out(s, so0, 3, e.u, o.u)) { return rc; } }
    if (so1) { if (int rc = copy_out(s, so1, 3, e.g, o.g)) { return rc; } }
    if (cfl)
    {
        double * ev[3] = {e.cfl, nullptr, nullptr};
        double * od[3] = {o.cfl, nullptr, nullptr};
        if (int rc = copy_out(s, cfl, 1, ev, od)) { return rc; }
    }
    return 0;
}

int scg_e1d_phase(scg_e1d * s, int phase, int odd_plane, int alpha)
{
    SCG_REQUIRE(alpha == 1 || alpha == 2, "scg_e1d_phase: alpha must be 1 or 2");
    int const S = odd_plane ? 1 : 0; // odd_plane names the SOURCE plane (Euler1DCore.cpp:99)
    switch (phase)
    {
    case SCG_PH_SO0: launch_half(s, S, F_SO0, alpha); break;
    case SCG_PH_CFL: launch_half(s, S, F_CFL, alpha); break;
    case SCG_PH_SO1: launch_half(s, S, F_SO1, alpha); break;
    case SCG_PH_BC0:
    case SCG_PH_BC1:
        e1d_bc_kernel<<<1, 32, 0, stream()>>>(s->P, (s->N - 3) / 2, phase == SCG_PH_BC1);
        count_launch();
        break;
    default: return fail("scg_e1d_phase: unknown phase %d", phase);
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_e1d_march(scg_e1d * s, int alpha, int64_t steps)
{
    SCG_REQUIRE(alpha == 1 || alpha == 2, "scg_e1d_march: alpha must be 1 or 2");
    if (steps <= 0) { return 0; } // a non-positive count is a no-op, like the reference's `for (it = 0; it < steps; ++it)`
    // Small grids are launch-latency bound (config 1: 10^4 points, 2000 half steps): replay one
    // captured step as a CUDA graph.  Large grids launch directly.
    bool const use_graph = s->N <= (int64_t(1) << 22) && steps >= 8;
    if (use_graph)
    {
        if (!s->graph[alpha])
        {
            cudaGraph_t g = nullptr;
            SCG_CUDA(cudaStreamBeginCapture(stream(), cudaStreamCaptureModeThreadLocal));
            for (int k = 0; k < 4; ++k) { enqueue_step(s, alpha); }
            SCG_CUDA(cudaStreamEndCapture(stream(), &g));
            SCG_CUDA(cudaGraphInstantiate(&s->graph[alpha], g, 0));
            cudaGraphDestroy(g);
            count_launch(-8); // capture does not execute
        }
        int64_t it = 0;
        for (; it + 4 <= steps; it += 4)
        {
            SCG_CUDA(cudaGraphLaunch(s->graph[alpha], stream()));
            count_launch(8);
        }
        for (; it < steps; ++it) { enqueue_step(s, alpha); }
    }
    else
    {
        for (int64_t it = 0; it < steps; ++it) { enqueue_step(s, alpha); }
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_e1d_reduce(scg_e1d * s, double * cfl_min, double * cfl_max, double * mass)
{
    int const nblk = 512;
    e1d_reduce_kernel<<<nblk, 256, 0, stream()>>>(s->P, s->ne, s->no, s->red.p);
    count_launch();
    std::vector<double> h(3 * nblk);
    SCG_CUDA(cudaMemcpyAsync(h.data(), s->red.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, stream()));
    SCG_CUDA(cudaStreamSynchronize(stream()));
    double mn = DBL_MAX, mx = -DBL_MAX, sm = 0.0;
    for (int b = 0; b < nblk; ++b)
    {
        mn = h[3 * b] < mn ? h[3 * b] : mn;
        mx = h[3 * b + 1] > mx ? h[3 * b + 1] : mx;
        sm += h[3 * b + 2];
    }
    if (cfl_min) { *cfl_min = mn; }
    if (cfl_max) { *cfl_max = mx; }
    if (mass) { *mass = sm; }
    return 0;
}

} // extern "C"
