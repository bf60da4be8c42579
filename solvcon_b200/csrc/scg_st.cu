// scg_st.cu -- 1D scalar CESE (linear scalar, inviscid Burgers) on the device.
//
// Device twin of spacetime::SolverBase<...> (cpp/solvcon/spacetime/core.hpp:946-1036) with the
// LinearScalarSelm / InviscidBurgersSelm kernels (kernel/linear_scalar.hpp:45-87,
// kernel/inviscid_burgers.hpp:51-109).
//
// Layout.  The reference interleaves the two time planes in one array of nx = 2*ncelm+5 slots
// (core.hpp:183-191).  On the device each plane is its own unit-stride array, so one half step
// streams the source plane in and the target plane out with no wasted sectors:
//     even plane  e = k/2        k = 0,2,...,2n+4   (n+3 slots)
//     odd  plane  o = (k-1)/2    k = 1,3,...,2n+3   (n+2 slots, o=0 and o=n+1 are the periodic ghosts)
// A half step into plane T reads plane S = 1-T:  target t uses sources s = t+a and t+a+1 with
// a = 0 for T = odd and a = -1 for T = even.
//
// One fused kernel per half step does so0 -> (ghost so0) -> cfl -> so1 -> (ghost so1) for its own
// target point; the reference's order of those phases (core.hpp:1006-1025) only matters per point
// because every phase is a pure gather from the *other* plane plus the point's own new so0.
// Arithmetic keeps the reference's expression order; the strict build (-fmad=false) is bit-equal.
#include "scg_common.cuh"

#include <cfloat>
#include <vector>

using namespace scg;

namespace
{

struct Plane
{
    double * x;
    double * so0;
    double * so1;
    double * cfl;
};

struct Planes
{
    double * x[2];
    double * so0[2];
    double * so1[2];
    double * cfl[2];
    __host__ __device__ Plane plane(int p) const { return Plane{x[p], so0[p], so1[p], cfl[p]}; }
};

enum : int
{
    F_SO0 = 1,
    F_CFL = 2,
    F_SO1 = 4,
    F_BC0 = 8,  // also mirror so0 (and cfl when F_CFL) into the periodic ghosts
    F_BC1 = 16, // also mirror so1 into the periodic ghosts
};

template <int ALPHA>
__device__ __forceinline__ double pow_alpha(double v)
{
    if (0 == ALPHA) { return 1.0; }
    if (1 == ALPHA) { return v; }
    return v * v;
}

// Selm::tp for the two equations (linear_scalar.hpp:66-73, inviscid_burgers.hpp:83-92)
template <int EQ>
__device__ __forceinline__ double se_tp(double x, double xctr, double u, double ux, double hdt, double qdt)
{
    double const displacement = x - xctr;
    double ret;
    if (0 == EQ)
    {
        ret = u;
        ret += displacement * ux;
        ret -= qdt * ux;
    }
    else
    {
        double const u_2 = u * u;
        ret = 0.5 * u_2;
        ret += displacement * u * ux;
        ret -= qdt * u_2 * ux;
    }
    return hdt * ret;
}

template <int EQ>
__device__ __forceinline__ double se_cfl(double xneg, double x, double xpos, double u, double hdt)
{
    double const dxneg = x - xneg;
    double const dxpos = xpos - x;
    double const hdx = (dxpos < dxneg) ? dxpos : dxneg; // std::min(dxneg, dxpos)
    if (0 == EQ) { return hdt / hdx; }
    return fabs(u) * hdt / hdx;
}

// One thread per target point t in [tbeg, tend] of plane T.
template <int EQ, int ALPHA, int FLAGS>
__global__ void __launch_bounds__(256) st_half_kernel(Plane const PS, Plane const PT, int T, int64_t tbeg, int64_t tend,
                                                      int64_t n, double hdt, double qdt)
{
    int64_t const t = tbeg + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t > tend) { return; }
    int64_t const s = t + (T ? 0 : -1);
    double const * __restrict__ xS = PS.x;
    double const * __restrict__ xT = PT.x;
    double const * __restrict__ uS = PS.so0;
    double const * __restrict__ gS = PS.so1;

    // coordinates k-2 .. k+2 around the target
    double const xm2 = xT[t - 1];
    double const xm1 = xS[s];
    double const x0 = xT[t];
    double const xp1 = xS[s + 1];
    double const xp2 = xT[t + 1];
    // left (xn) and right (xp) source solution elements
    double const un = uS[s];
    double const gn = gS[s];
    double const up = uS[s + 1];
    double const gp = gS[s + 1];
    double const xctr_n = (xm2 + x0) / 2;  // Selm::xctr, core.hpp:483
    double const xctr_p = (x0 + xp2) / 2;

    double unew;
    if (FLAGS & F_SO0)
    {
        // Celm::calc_so0, core.hpp:645-654
        double const disp_n = 0.5 * (xm1 + x0) - xctr_n;             // se_xn.xp(): displacement to (x+xpos)/2
        double const xp_n = (x0 - xm1) * (un + disp_n * gn);         // dxpos * (...)
        double const flux_ll = xp_n + se_tp<EQ>(xm1, xctr_n, un, gn, hdt, qdt);
        double const disp_p = 0.5 * (xp1 + x0) - xctr_p;             // se_xp.xn(): displacement to (x+xneg)/2
        double const xn_p = (xp1 - x0) * (up + disp_p * gp);         // dxneg * (...)
        double const flux_ur = xn_p - se_tp<EQ>(xp1, xctr_p, up, gp, hdt, qdt);
        unew = (flux_ll + flux_ur) / (xp1 - xm1);
        PT.so0[t] = unew;
    }
    else
    {
        unew = PT.so0[t];
    }
    if (FLAGS & F_CFL) { PT.cfl[t] = se_cfl<EQ>(xm1, x0, xp1, unew, hdt); }
    double gnew = 0.0;
    if (FLAGS & F_SO1)
    {
        // Celm::calc_so1_alpha, core.hpp:656-672; so0p: linear_scalar.hpp:75-81
        double upn = un;
        upn += (xm1 - xctr_n) * gn;
        upn -= hdt * gn;
        double upp = up;
        upp += (xp1 - xctr_p) * gp;
        upp -= hdt * gp;
        double const duxn = (unew - upn) / (x0 - xm1);
        double const duxp = (upp - unew) / (xp1 - x0);
        double const fan = pow_alpha<ALPHA>(fabs(duxn));
        double const fap = pow_alpha<ALPHA>(fabs(duxp));
        gnew = (fap * duxn + fan * duxp) / (fap + fan + DBL_MIN);
        PT.so1[t] = gnew;
    }
    if (FLAGS & (F_BC0 | F_BC1))
    {
        // periodic ghosts of the odd plane (core.hpp:982-1004): o=0 <- o=n, o=n+1 <- o=1.
        // The ghost cfl (update_cfl(true) covers the ghosts, core.hpp:958-967) uses the ghost's own
        // coordinates and the mirrored so0.
        if (t == n)
        {
            if (FLAGS & F_BC0) { PT.so0[0] = unew; }
            if ((FLAGS & F_BC0) && (FLAGS & F_CFL)) { PT.cfl[0] = se_cfl<EQ>(xS[0], xT[0], xS[1], unew, hdt); }
            if (FLAGS & F_BC1) { PT.so1[0] = gnew; }
        }
        if (t == 1)
        {
            if (FLAGS & F_BC0) { PT.so0[n + 1] = unew; }
            if ((FLAGS & F_BC0) && (FLAGS & F_CFL))
            {
                PT.cfl[n + 1] = se_cfl<EQ>(xS[n + 1], xT[n + 1], xS[n + 2], unew, hdt);
            }
            if (FLAGS & F_BC1) { PT.so1[n + 1] = gnew; }
        }
    }
}

// update_cfl(odd) alone: every point of plane T the reference visits (core.hpp:958-967)
template <int EQ>
__global__ void st_cfl_kernel(Plane const PS, Plane const PT, int T, int64_t tbeg, int64_t tend, double hdt)
{
    int64_t const t = tbeg + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t > tend) { return; }
    int64_t const s = t + (T ? 0 : -1);
    PT.cfl[t] = se_cfl<EQ>(PS.x[s], PT.x[t], PS.x[s + 1], PT.so0[t], hdt);
}

__global__ void st_bc_kernel(double * a, int64_t n)
{
    if (0 == threadIdx.x && 0 == blockIdx.x)
    {
        a[0] = a[n];
        a[n + 1] = a[1];
    }
}

// interleaved (reference) chunk -> planes.  k0 is the global index of chunk[0].
__global__ void st_deinterleave_kernel(double const * __restrict__ chunk, int64_t k0, int64_t count,
                                       double * __restrict__ even, double * __restrict__ odd)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) { return; }
    int64_t const k = k0 + i;
    double const v = chunk[i];
    if (k & 1) { odd[(k - 1) >> 1] = v; }
    else { even[k >> 1] = v; }
}

__global__ void st_interleave_kernel(double * __restrict__ chunk, int64_t k0, int64_t count,
                                     double const * __restrict__ even, double const * __restrict__ odd)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) { return; }
    int64_t const k = k0 + i;
    chunk[i] = (k & 1) ? odd[(k - 1) >> 1] : even[k >> 1];
}

// min/max of cfl over both planes' visited points and sum of so0 over the even plane interior
__global__ void st_reduce_kernel(Planes P, int64_t n, double * out /* [3*gridDim.x] */)
{
    double mn = DBL_MAX, mx = -DBL_MAX, sm = 0.0;
    int64_t const stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i <= n + 1; i += stride)
    {
        if (i >= 1)
        {
            double const c = P.cfl[0][i]; // even e = 1..n+1
            mn = fmin(mn, c);
            mx = fmax(mx, c);
            sm += P.so0[0][i];
        }
        if (i >= 1 && i <= n)
        {
            double const c = P.cfl[1][i]; // odd interior o = 1..n
            mn = fmin(mn, c);
            mx = fmax(mx, c);
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

// halo support for slab decomposition: the odd-plane ghosts carry the neighbour rank's interior
__global__ void st_halo_pack_kernel(Planes P, int64_t n, double * left, double * right)
{
    if (0 == threadIdx.x && 0 == blockIdx.x)
    {
        left[0] = P.so0[1][1];
        left[1] = P.so1[1][1];
        right[0] = P.so0[1][n];
        right[1] = P.so1[1][n];
    }
}

template <int EQ>
__global__ void st_halo_unpack_kernel(Planes P, int64_t n, double const * left, double const * right, double hdt)
{
    if (0 == threadIdx.x && 0 == blockIdx.x)
    {
        // value arriving from the left neighbour is its o = n; from the right neighbour its o = 1
        P.so0[1][0] = left[0];
        P.so1[1][0] = left[1];
        P.cfl[1][0] = se_cfl<EQ>(P.x[0][0], P.x[1][0], P.x[0][1], left[0], hdt);
        P.so0[1][n + 1] = right[0];
        P.so1[1][n + 1] = right[1];
        P.cfl[1][n + 1] = se_cfl<EQ>(P.x[0][n + 1], P.x[1][n + 1], P.x[0][n + 2], right[0], hdt);
    }
}

constexpr int64_t STAGE_DOUBLES = int64_t(1) << 25; // 256 MiB staging chunk

} // namespace

struct scg_st
{
    int eq = 0;
    int64_t n = 0; // ncelm
    int64_t nx = 0;
    double dt = 0, hdt = 0, qdt = 0;
    DevBuf<double> buf[8]; // x,so0,so1,cfl for even then odd
    DevBuf<double> stage;
    DevBuf<double> red;
    Planes P{};
    bool halo_mode = false; // ghosts come from scg_st_halo_unpack instead of the periodic wrap
    // ring of slabs over the library's communicator: [send left(2) | send right(2) | recv from left(2) | recv from right(2)]
    DevBuf<double> ring;
    cudaEvent_t ev_packed = nullptr, ev_recv = nullptr;
    ~scg_st()
    {
        if (ev_packed) { cudaEventDestroy(ev_packed); }
        if (ev_recv) { cudaEventDestroy(ev_recv); }
    }
};

namespace
{

int64_t plane_len(scg_st const * s, int plane) { return plane ? s->n + 2 : s->n + 3; }

template <int EQ, int ALPHA>
void launch_half_a(scg_st * s, int T, int flags)
{
    int64_t const tbeg = 1;
    int64_t const tend = T ? s->n : s->n + 1;
    int64_t const cnt = tend - tbeg + 1;
    dim3 const block(256);
    dim3 const grid(static_cast<unsigned>(ceil_div(cnt, 256)));
#define SCG_ST_CASE(FL)                                                                                    \
    case FL:                                                                                               \
        st_half_kernel<EQ, ALPHA, FL><<<grid, block, 0, stream()>>>(s->P.plane(1 - T), s->P.plane(T), T, tbeg, tend, s->n, s->hdt, s->qdt); \
        break;
    switch (flags)
    {
        SCG_ST_CASE(F_SO0)
        SCG_ST_CASE(F_SO1)
        SCG_ST_CASE(F_SO0 | F_CFL | F_SO1)
        SCG_ST_CASE(F_SO0 | F_CFL | F_SO1 | F_BC0 | F_BC1)
    default: break;
    }
#undef SCG_ST_CASE
    count_launch();
}

template <int EQ>
void launch_half_e(scg_st * s, int T, int flags, int alpha)
{
    if (0 == alpha) { launch_half_a<EQ, 0>(s, T, flags); }
    else if (1 == alpha) { launch_half_a<EQ, 1>(s, T, flags); }
    else { launch_half_a<EQ, 2>(s, T, flags); }
}

void launch_half(scg_st * s, int T, int flags, int alpha)
{
    if (0 == s->eq) { launch_half_e<0>(s, T, flags, alpha); }
    else { launch_half_e<1>(s, T, flags, alpha); }
}

int copy_in(scg_st * s, double const * host, double * even, double * odd)
{
    for (int64_t k0 = 0; k0 < s->nx; k0 += STAGE_DOUBLES)
    {
        int64_t const cnt = (s->nx - k0 < STAGE_DOUBLES) ? s->nx - k0 : STAGE_DOUBLES;
        SCG_CUDA(cudaMemcpyAsync(s->stage.p, host + k0, sizeof(double) * cnt, cudaMemcpyHostToDevice, stream()));
        st_deinterleave_kernel<<<static_cast<unsigned>(ceil_div(cnt, 256)), 256, 0, stream()>>>(s->stage.p, k0, cnt, even, odd);
        count_launch();
        // the staging buffer is reused by the next chunk; same-stream ordering makes that safe for
        // pinned memory, and pageable copies are synchronous w.r.t. the host anyway
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int copy_out(scg_st * s, double * host, double const * even, double const * odd)
{
    for (int64_t k0 = 0; k0 < s->nx; k0 += STAGE_DOUBLES)
    {
        int64_t const cnt = (s->nx - k0 < STAGE_DOUBLES) ? s->nx - k0 : STAGE_DOUBLES;
        st_interleave_kernel<<<static_cast<unsigned>(ceil_div(cnt, 256)), 256, 0, stream()>>>(s->stage.p, k0, cnt, even, odd);
        count_launch();
        SCG_CUDA(cudaMemcpyAsync(host + k0, s->stage.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost, stream()));
    }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    SCG_CUDA(cudaGetLastError());
    return 0;
}

} // namespace

extern "C" {

int scg_st_create(scg_st ** out, int eq, int64_t ncelm, double time_increment)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_REQUIRE(out != nullptr, "scg_st_create: null output");
    SCG_REQUIRE(eq == SCG_EQ_LINEAR || eq == SCG_EQ_BURGERS, "scg_st_create: eq must be 0 (linear) or 1 (Burgers)");
    SCG_REQUIRE(ncelm >= 1, "scg_st_create: ncelm %lld smaller than 1", static_cast<long long>(ncelm));
    SCG_REQUIRE(2 * ncelm + 5 < (int64_t(1) << 31), "scg_st_create: 2*ncelm+5 must fit int32 like the reference");
    auto * s = new scg_st;
    s->eq = eq;
    s->n = ncelm;
    s->nx = 2 * ncelm + 5;
    for (int plane = 0; plane < 2; ++plane)
    {
        for (int a = 0; a < 4; ++a)
        {
            cudaError_t const e = s->buf[plane * 4 + a].alloc_zero(plane_len(s, plane) + 8);
            if (cudaSuccess != e)
            {
                delete s;
                return fail("scg_st_create: device allocation failed: %s", cudaGetErrorString(e));
            }
        }
        s->P.x[plane] = s->buf[plane * 4 + 0].p;
        s->P.so0[plane] = s->buf[plane * 4 + 1].p;
        s->P.so1[plane] = s->buf[plane * 4 + 2].p;
        s->P.cfl[plane] = s->buf[plane * 4 + 3].p;
    }
    int64_t const stage = s->nx < STAGE_DOUBLES ? s->nx : STAGE_DOUBLES;
    if (cudaSuccess != s->stage.alloc(stage) || cudaSuccess != s->red.alloc(3 * 1024))
    {
        delete s;
        return fail("scg_st_create: device allocation failed");
    }
    scg_st_set_time_increment(s, time_increment);
    *out = s;
    return 0;
}

int scg_st_destroy(scg_st * s)
{
    if (s)
    {
        cudaStreamSynchronize(stream());
        delete s;
    }
    return 0;
}

int scg_st_set_time_increment(scg_st * s, double time_increment)
{
    // Field::set_time_increment, core.cpp:111-116
    s->dt = time_increment;
    s->hdt = 0.5 * time_increment;
    s->qdt = 0.25 * time_increment;
    return 0;
}

int scg_st_grid_uniform(scg_st * s, double xmin, double xmax)
{
    SCG_REQUIRE(xmin < xmax, "Grid::Grid(xmin=%g, xmax=%g, ncelm=%lld) invalid arguments: xmin >= xmax", xmin, xmax,
                static_cast<long long>(s->n));
    // Host-side restatement of Grid::Grid + init_from_array (core.cpp:14-99).  The CE boundaries are a
    // sequential floating-point recurrence (x[i] = x[i-1] + h), which has no bit-exact parallel form,
    // so this one-time setup stays on the host and the result is uploaded.
    int64_t const n = s->n;
    std::vector<double> x(static_cast<size_t>(s->nx));
    double const xspace = (xmax - xmin) / static_cast<double>(n);
    double run = xmin;
    x[2] = xmin;
    for (int64_t i = 1; i < n; ++i)
    {
        run = run + xspace;
        x[2 + 2 * i] = run;
    }
    x[2 + 2 * n] = xmax;
    for (int64_t i = 0; i < n; ++i) { x[3 + 2 * i] = (x[2 + 2 * i] + x[4 + 2 * i]) / 2; }
    for (int64_t it = 1; it <= 2; ++it)
    {
        x[2 - it] = x[2] + x[2] - x[2 + it];
        x[s->nx - 3 + it] = x[s->nx - 3] + x[s->nx - 3] - x[s->nx - 3 - it];
    }
    int const rc = copy_in(s, x.data(), s->P.x[0], s->P.x[1]);
    if (rc) { return rc; }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_st_upload(scg_st * s, const double * x, const double * so0, const double * so1, const double * cfl)
{
    if (x) { if (int rc = copy_in(s, x, s->P.x[0], s->P.x[1])) { return rc; } }
    if (so0) { if (int rc = copy_in(s, so0, s->P.so0[0], s->P.so0[1])) { return rc; } }
    if (so1) { if (int rc = copy_in(s, so1, s->P.so1[0], s->P.so1[1])) { return rc; } }
    if (cfl) { if (int rc = copy_in(s, cfl, s->P.cfl[0], s->P.cfl[1])) { return rc; } }
    SCG_CUDA(cudaStreamSynchronize(stream()));
    return 0;
}

int scg_st_download(scg_st * s, double * x, double * so0, double * so1, double * cfl)
{
    if (x) { if (int rc = copy_out(s, x, s->P.x[0], s->P.x[1])) { return rc; } }
    if (so0) { if (int rc = copy_out(s, so0, s->P.so0[0], s->P.so0[1])) { return rc; } }
    if (so1) { if (int rc = copy_out(s, so1, s->P.so1[0], s->P.so1[1])) { return rc; } }
    if (cfl) { if (int rc = copy_out(s, cfl, s->P.cfl[0], s->P.cfl[1])) { return rc; } }
    return 0;
}

int scg_st_phase(scg_st * s, int phase, int odd_plane, int alpha)
{
    SCG_REQUIRE(alpha >= 0 && alpha <= 2, "scg_st_phase: alpha must be 0, 1 or 2");
    int const T = odd_plane ? 0 : 1; // CEs of the even plane advance the odd plane and vice versa
    switch (phase)
    {
    case SCG_PH_SO0: launch_half(s, T, F_SO0, alpha); break;
    case SCG_PH_SO1: launch_half(s, T, F_SO1, alpha); break;
    case SCG_PH_CFL:
    {
        int const plane = odd_plane ? 1 : 0;
        int64_t const tbeg = plane ? 0 : 1;
        int64_t const tend = s->n + 1;
        unsigned const grid = static_cast<unsigned>(ceil_div(tend - tbeg + 1, 256));
        if (0 == s->eq) { st_cfl_kernel<0><<<grid, 256, 0, stream()>>>(s->P.plane(1 - plane), s->P.plane(plane), plane, tbeg, tend, s->hdt); }
        else { st_cfl_kernel<1><<<grid, 256, 0, stream()>>>(s->P.plane(1 - plane), s->P.plane(plane), plane, tbeg, tend, s->hdt); }
        count_launch();
        break;
    }
    case SCG_PH_BC0:
        st_bc_kernel<<<1, 32, 0, stream()>>>(s->P.so0[1], s->n);
        count_launch();
        break;
    case SCG_PH_BC1:
        st_bc_kernel<<<1, 32, 0, stream()>>>(s->P.so1[1], s->n);
        count_launch();
        break;
    default: return fail("scg_st_phase: unknown phase %d", phase);
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_st_march(scg_st * s, int alpha, int64_t steps)
{
    SCG_REQUIRE(alpha >= 0 && alpha <= 2, "scg_st_march: alpha must be 0, 1 or 2");
    if (steps <= 0) { return 0; } // a non-positive count is a no-op, like the reference's `for (it = 0; it < steps; ++it)`
    int const bc = s->halo_mode ? 0 : (F_BC0 | F_BC1);
    for (int64_t it = 0; it < steps; ++it)
    {
        launch_half(s, 1, F_SO0 | F_CFL | F_SO1 | bc, alpha); // march_half1_alpha, core.hpp:1006-1015
        launch_half(s, 0, F_SO0 | F_CFL | F_SO1, alpha);      // march_half2_alpha, core.hpp:1017-1025
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_st_set_halo_mode(scg_st * s, int on)
{
    SCG_REQUIRE(s, "scg_st_set_halo_mode: null solver");
    s->halo_mode = 0 != on;
    return 0;
}

int scg_st_march_half(scg_st * s, int half, int alpha)
{
    SCG_REQUIRE(s, "scg_st_march_half: null solver");
    SCG_REQUIRE(1 == half || 2 == half, "scg_st_march_half: half must be 1 or 2");
    SCG_REQUIRE(alpha >= 0 && alpha <= 2, "scg_st_march_half: alpha must be 0, 1 or 2");
    int const bc = s->halo_mode ? 0 : (F_BC0 | F_BC1);
    if (1 == half) { launch_half(s, 1, F_SO0 | F_CFL | F_SO1 | bc, alpha); } // march_half1_alpha, core.hpp:1006-1015
    else { launch_half(s, 0, F_SO0 | F_CFL | F_SO1, alpha); }                // march_half2_alpha, core.hpp:1017-1025
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_st_reduce(scg_st * s, double * cfl_min, double * cfl_max, double * so0_sum)
{
    int const nblk = 512;
    st_reduce_kernel<<<nblk, 256, 0, stream()>>>(s->P, s->n, s->red.p);
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
    if (so0_sum) { *so0_sum = sm; }
    return 0;
}

int scg_st_halo_pack(scg_st * s, int odd_plane, double * dev_send_left, double * dev_send_right)
{
    SCG_REQUIRE(odd_plane, "scg_st_halo_pack: only the odd plane has ghosts (core.hpp:982-1004)");
    s->halo_mode = true;
    st_halo_pack_kernel<<<1, 32, 0, stream()>>>(s->P, s->n, dev_send_left, dev_send_right);
    count_launch();
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_st_halo_unpack(scg_st * s, int odd_plane, const double * dev_recv_left, const double * dev_recv_right)
{
    SCG_REQUIRE(odd_plane, "scg_st_halo_unpack: only the odd plane has ghosts (core.hpp:982-1004)");
    s->halo_mode = true;
    if (0 == s->eq) { st_halo_unpack_kernel<0><<<1, 32, 0, stream()>>>(s->P, s->n, dev_recv_left, dev_recv_right, s->hdt); }
    else { st_halo_unpack_kernel<1><<<1, 32, 0, stream()>>>(s->P, s->n, dev_recv_left, dev_recv_right, s->hdt); }
    count_launch();
    SCG_CUDA(cudaGetLastError());
    return 0;
}


// One slab of the periodic ring: march_alpha over the library's communicator.  Per step: half 1 on the
// slab, {so0, so1} of its first / last odd-plane point to the left / right ring neighbour (two 16-byte messages each way,
// one ncclGroup on the communication stream), the neighbours' points into the ghosts (which also evaluates the ghosts'
// cfl), half 2 -- the periodic copy of treat_boundary_so0/so1 (core.hpp:982-1004) done across ranks.  Stream-ordered.
int scg_st_march_ring(scg_st * s, int alpha, int64_t steps, int left_rank, int right_rank)
{
    SCG_REQUIRE(s, "scg_st_march_ring: null solver");
    SCG_REQUIRE(alpha >= 0 && alpha <= 2, "scg_st_march_ring: alpha must be 0, 1 or 2");
    SCG_REQUIRE(comm_ready(), "scg_st_march_ring: no communicator (call scg_comm_init first)");
    if (steps <= 0) { return 0; }
    if (!s->ring.p)
    {
        SCG_CUDA(s->ring.alloc_zero(8));
        SCG_CUDA(cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
        SCG_CUDA(cudaEventCreateWithFlags(&s->ev_recv, cudaEventDisableTiming));
        SCG_CUDA(cudaStreamSynchronize(stream()));
    }
    s->halo_mode = true;
    double * const b = s->ring.p;
    for (int64_t it = 0; it < steps; ++it)
    {
        if (int rc = scg_st_march_half(s, 1, alpha)) { return rc; }
        if (int rc = scg_st_halo_pack(s, 1, b, b + 2)) { return rc; }
        SCG_CUDA(cudaEventRecord(s->ev_packed, stream()));
        SCG_CUDA(cudaStreamWaitEvent(comm_stream(), s->ev_packed, 0));
        if (int rc = comm_group_start()) { return rc; }
        // with two ranks both messages go to the same peer: NCCL matches sends and receives per peer in issue order,
        // so the order here (left message first, right message second) must mirror the peer's (its right, then left)
        if (int rc = comm_send(b, 2, left_rank)) { return rc; }
        if (int rc = comm_send(b + 2, 2, right_rank)) { return rc; }
        if (int rc = comm_recv(b + 6, 2, right_rank)) { return rc; }
        if (int rc = comm_recv(b + 4, 2, left_rank)) { return rc; }
        if (int rc = comm_group_end()) { return rc; }
        SCG_CUDA(cudaEventRecord(s->ev_recv, comm_stream()));
        SCG_CUDA(cudaStreamWaitEvent(stream(), s->ev_recv, 0));
        // my left ghost is the left neighbour's LAST odd-plane point (its "right" message), and vice versa
        if (int rc = scg_st_halo_unpack(s, 1, b + 4, b + 6)) { return rc; }
        if (int rc = scg_st_march_half(s, 2, alpha)) { return rc; }
    }
    SCG_CUDA(cudaGetLastError());
    return 0;
}

} // extern "C"
