/*
 * cese_oracle.c -- TEST INFRASTRUCTURE ONLY (see cese_oracle.h).
 *
 * Plain-C restatement of the reference's CESE march, written against flat
 * arrays and explicit indices (the reference walks Selm/Celm cursor objects
 * and SimpleArray accessors).  Each function cites the reference lines whose
 * arithmetic it restates; operation order inside every expression is kept so
 * the result is bit-identical to the reference's non-FMA x86-64 build.
 */
#include "cese_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

/* ====================================================================== */
/* 1D scalar solvers                                                      */
/* ====================================================================== */

/* Grid::Grid(xmin,xmax,ncelm) + Grid::init_from_array, spacetime/core.cpp:14-99 */
static void grid_from_xloc(double const * xloc, int64_t ncelm, double * x)
{
    int64_t const nx = 2 * ncelm + 5;
    for (int64_t i = 0; i <= ncelm; ++i) { x[2 + 2 * i] = xloc[i]; }           /* core.cpp:77-80 */
    for (int64_t i = 0; i < ncelm; ++i)                                        /* core.cpp:82-86 */
    {
        int64_t const k = 3 + 2 * i;
        x[k] = (x[k - 1] + x[k + 1]) / 2;
    }
    for (int64_t it = 1; it <= 2; ++it)                                        /* core.cpp:88-98 */
    {
        x[2 - it] = x[2] + x[2] - x[2 + it];
        x[nx - 3 + it] = x[nx - 3] + x[nx - 3] - x[nx - 3 - it];
    }
}

int orc_grid_fill(double xmin, double xmax, int64_t ncelm, double * x)
{
    if (ncelm < 1 || !(xmin < xmax)) { return -1; }
    /* core.cpp:36-43: CE boundaries by cumulative addition, last one pinned to xmax.
     * They are staged in the even slots of x itself (no scratch allocation). */
    double const xspace = (xmax - xmin) / (double)ncelm;
    double prev = xmin;
    x[2] = xmin;
    for (int64_t i = 1; i < ncelm; ++i)
    {
        prev = prev + xspace;
        x[2 + 2 * i] = prev;
    }
    x[2 + 2 * ncelm] = xmax;
    int64_t const nx = 2 * ncelm + 5;
    for (int64_t i = 0; i < ncelm; ++i)
    {
        int64_t const k = 3 + 2 * i;
        x[k] = (x[k - 1] + x[k + 1]) / 2;
    }
    for (int64_t it = 1; it <= 2; ++it)
    {
        x[2 - it] = x[2] + x[2] - x[2 + it];
        x[nx - 3 + it] = x[nx - 3] + x[nx - 3] - x[nx - 3 - it];
    }
    return 0;
}

int orc_grid_fill_xloc(double const * xloc, int64_t nloc, double * x)
{
    if (nloc < 2) { return -1; }
    for (int64_t i = 0; i + 1 < nloc; ++i)
    {
        if (xloc[i] >= xloc[i + 1]) { return -1; }
    }
    grid_from_xloc(xloc, nloc - 1, x);
    return 0;
}

/* Per-SE quantities at coordinate index j.  Selm::{dxneg,dxpos,xctr} core.hpp:479-483. */
typedef struct se1
{
    double x, xneg, xpos, xctr, dxneg, dxpos, u, ux;
} se1;

static se1 se_load(double const * x, double const * so0, double const * so1, int64_t j)
{
    se1 s;
    s.x = x[j];
    s.xneg = x[j - 1];
    s.xpos = x[j + 1];
    s.xctr = (s.xneg + s.xpos) / 2;
    s.dxneg = s.x - s.xneg;
    s.dxpos = s.xpos - s.x;
    s.u = so0[j];
    s.ux = so1[j];
    return s;
}

/* linear_scalar.hpp:45-55 and inviscid_burgers.hpp:51-64 (identical) */
static double se_xn(se1 const * s)
{
    double const displacement = 0.5 * (s->x + s->xneg) - s->xctr;
    return s->dxneg * (s->u + displacement * s->ux);
}
static double se_xp(se1 const * s)
{
    double const displacement = 0.5 * (s->x + s->xpos) - s->xctr;
    return s->dxpos * (s->u + displacement * s->ux);
}
/* linear_scalar.hpp:66-73 / inviscid_burgers.hpp:83-92 */
static double se_tp(int eq, se1 const * s, double hdt, double qdt)
{
    double const displacement = s->x - s->xctr;
    double ret;
    if (0 == eq)
    {
        ret = s->u;
        ret += displacement * s->ux;
        ret -= qdt * s->ux;
    }
    else
    {
        double const u_2 = s->u * s->u;
        ret = 0.5 * u_2;
        ret += displacement * s->u * s->ux;
        ret -= qdt * u_2 * s->ux;
    }
    return hdt * ret;
}
/* linear_scalar.hpp:75-81 / inviscid_burgers.hpp:97-103 (same formula in both) */
static double se_so0p(se1 const * s, double hdt)
{
    double ret = s->u;
    ret += (s->x - s->xctr) * s->ux;
    ret -= hdt * s->ux;
    return ret;
}

static double pow_alpha(double v, int alpha) /* math/math.hpp:23-46 */
{
    if (0 == alpha) { return 1.0; }
    if (1 == alpha) { return v; }
    return v * v;
}

/* Target range of a half step: odd_plane==0 -> CE ic in [0,ncelm) sits at k=3+2ic;
 * odd_plane==1 -> ic in [-1,ncelm) sits at k=4+2ic (core.hpp:183-191,946-980). */
static void st_so0(int eq, int64_t ncelm, double const * x, double * so0, double const * so1,
                   double hdt, double qdt, int odd)
{
    int64_t const kbeg = odd ? 2 : 3;
    int64_t const kend = odd ? 2 * ncelm + 2 : 2 * ncelm + 1;
    for (int64_t k = kbeg; k <= kend; k += 2)
    {
        /* Celm::calc_so0, core.hpp:645-654 */
        se1 const sn = se_load(x, so0, so1, k - 1);
        se1 const sp = se_load(x, so0, so1, k + 1);
        double const flux_ll = se_xp(&sn) + se_tp(eq, &sn, hdt, qdt);
        double const flux_ur = se_xn(&sp) - se_tp(eq, &sp, hdt, qdt);
        so0[k] = (flux_ll + flux_ur) / (x[k + 1] - x[k - 1]);
    }
}

static void st_so1(int eq, int64_t ncelm, double const * x, double const * so0, double * so1,
                   double hdt, int odd, int alpha)
{
    (void)eq;
    int64_t const kbeg = odd ? 2 : 3;
    int64_t const kend = odd ? 2 * ncelm + 2 : 2 * ncelm + 1;
    for (int64_t k = kbeg; k <= kend; k += 2)
    {
        /* Celm::calc_so1_alpha, core.hpp:656-672 */
        se1 const sn = se_load(x, so0, so1, k - 1);
        se1 const sp = se_load(x, so0, so1, k + 1);
        double const upn = se_so0p(&sn, hdt);
        double const upp = se_so0p(&sp, hdt);
        double const utp = so0[k];
        double const duxn = (utp - upn) / sn.dxpos;
        double const duxp = (upp - utp) / sp.dxneg;
        double const fan = pow_alpha(fabs(duxn), alpha);
        double const fap = pow_alpha(fabs(duxp), alpha);
        so1[k] = (fap * duxn + fan * duxp) / (fap + fan + DBL_MIN);
    }
}

static void st_cfl(int eq, int64_t ncelm, double const * x, double const * so0, double * cfl,
                   double hdt, int odd)
{
    /* SolverBase::update_cfl, core.hpp:958-967: SE ic in [odd?-1:0, ncelm+1) at k=2+2ic+odd */
    int64_t const kbeg = odd ? 1 : 2;
    int64_t const kend = odd ? 2 * ncelm + 3 : 2 * ncelm + 2;
    for (int64_t k = kbeg; k <= kend; k += 2)
    {
        double const dxneg = x[k] - x[k - 1];
        double const dxpos = x[k + 1] - x[k];
        double const hdx = (dxpos < dxneg) ? dxpos : dxneg; /* std::min(dxneg, dxpos) */
        if (0 == eq) { cfl[k] = hdt / hdx; }                 /* linear_scalar.hpp:83-87 */
        else { cfl[k] = fabs(so0[k]) * hdt / hdx; }          /* inviscid_burgers.hpp:105-109 */
    }
}

static void st_bc(int64_t ncelm, double * a)
{
    /* periodic ghosts on the odd plane, core.hpp:982-1004 */
    a[1] = a[2 * ncelm + 1];
    a[2 * ncelm + 3] = a[3];
}

int orc_st_phase(int eq, int64_t ncelm, double const * x, double * so0, double * so1, double * cfl,
                 double dt, int phase, int odd, int alpha)
{
    double const hdt = 0.5 * dt;  /* core.cpp:113-115 */
    double const qdt = 0.25 * dt;
    if (alpha < 0 || alpha > 2) { return -1; }
    switch (phase)
    {
    case 0: st_so0(eq, ncelm, x, so0, so1, hdt, qdt, odd); break;
    case 1: st_cfl(eq, ncelm, x, so0, cfl, hdt, odd); break;
    case 2: st_so1(eq, ncelm, x, so0, so1, hdt, odd, alpha); break;
    case 3: st_bc(ncelm, so0); break;
    case 4: st_bc(ncelm, so1); break;
    default: return -1;
    }
    return 0;
}

int orc_st_march(int eq, int64_t ncelm, double const * x, double * so0, double * so1, double * cfl,
                 double dt, int alpha, int64_t steps)
{
    double const hdt = 0.5 * dt;
    double const qdt = 0.25 * dt;
    if (alpha < 0 || alpha > 2) { return -1; }
    for (int64_t it = 0; it < steps; ++it)
    {
        /* march_half1_alpha, core.hpp:1006-1015 */
        st_so0(eq, ncelm, x, so0, so1, hdt, qdt, 0);
        st_bc(ncelm, so0);
        st_cfl(eq, ncelm, x, so0, cfl, hdt, 1);
        st_so1(eq, ncelm, x, so0, so1, hdt, 0, alpha);
        st_bc(ncelm, so1);
        /* march_half2_alpha, core.hpp:1017-1025 */
        st_so0(eq, ncelm, x, so0, so1, hdt, qdt, 1);
        st_cfl(eq, ncelm, x, so0, cfl, hdt, 0);
        st_so1(eq, ncelm, x, so0, so1, hdt, 1, alpha);
    }
    return 0;
}

/* ====================================================================== */
/* 1D Euler                                                               */
/* ====================================================================== */

#define E1D_TINY 1.e-100 /* Euler1DCore.hpp:47,182 */

typedef struct e1d_se
{
    double ga, x, xneg, xpos, xctr;
    double u[3], ux[3], f[3], ut[3], ft[3], up[3];
} e1d_se;

/* Euler1DKernel::set_value + derive, Euler1DCore.hpp:194-244 */
static void e1d_derive(e1d_se * s, int64_t ic, double const * gamma, double const * coord,
                       double const * so0, double const * so1, double hdt)
{
    double jac[3][3];
    double const tiny = E1D_TINY;
    s->ga = gamma[ic];
    s->x = coord[ic];
    s->xneg = coord[ic - 1];
    s->xpos = coord[ic + 1];
    s->xctr = (s->xpos + s->xneg) * 0.5;
    for (int i = 0; i < 3; ++i)
    {
        s->u[i] = so0[ic * 3 + i];
        s->ux[i] = so1[ic * 3 + i];
    }
    double const ga = s->ga;
    double const * u = s->u;
    double const * ux = s->ux;
    jac[0][0] = 0.0;
    jac[0][1] = 1.0;
    jac[0][2] = 0.0;
    jac[1][0] = (ga - 3.0) / 2.0 * u[1] * u[1] / (u[0] * u[0] + tiny);
    jac[1][1] = -(ga - 3.0) * u[1] / (u[0] + tiny);
    jac[1][2] = ga - 1.0;
    jac[2][0] = (ga - 1.0) * u[1] * u[1] * u[1] / (u[0] * u[0] * u[0] + tiny) - ga * u[1] * u[2] / (u[0] * u[0] + tiny);
    jac[2][1] = ga * u[2] / (u[0] + tiny) - 3.0 / 2.0 * (ga - 1.0) * u[1] * u[1] / (u[0] * u[0] + tiny);
    jac[2][2] = ga * u[1] / (u[0] + tiny);

    s->f[0] = u[1];
    s->f[1] = (ga - 1.0) * u[2] + (3.0 - ga) / 2.0 * u[1] * u[1] / (u[0] + tiny);
    s->f[2] = ga * u[1] * u[2] / (u[0] + tiny) - (ga - 1.0) / 2.0 * u[1] * u[1] * u[1] / (u[0] * u[0] + tiny);

    for (int i = 0; i < 3; ++i)
    {
        s->ut[i] = -jac[i][0] * ux[0] - jac[i][1] * ux[1] - jac[i][2] * ux[2];
    }
    for (int i = 0; i < 3; ++i)
    {
        s->ft[i] = jac[i][0] * s->ut[0] + jac[i][1] * s->ut[1] + jac[i][2] * s->ut[2];
    }
    for (int i = 0; i < 3; ++i)
    {
        s->up[i] = u[i] + (s->x - s->xctr) * ux[i] + hdt * s->ut[i];
    }
}

/* calc_flux_ll / calc_flux_lr, Euler1DCore.hpp:246-266; sign=+1 for ll, -1 for lr */
static void e1d_flux(e1d_se const * s, double hdt, double qdt, int lr, double * r)
{
    double const xo = lr ? s->xneg : s->xpos;
    double const deltax = lr ? (s->x - s->xneg) : (s->xpos - s->x);
    double const dxmid = 0.5 * (s->x + xo) - s->xctr;
    double const dxctr = s->x - s->xctr;
    for (int i = 0; i < 3; ++i)
    {
        double const sp = deltax * (s->u[i] + dxmid * s->ux[i]);
        double const tm = hdt * (s->f[i] - (dxctr * s->ut[i]) + (qdt * s->ft[i]));
        r[i] = lr ? (sp - tm) : (sp + tm);
    }
}

static void e1d_range(int64_t ncoord, int odd, int64_t * start, int64_t * stop)
{
    /* Euler1DCore.cpp:73-74,99-100: BOUND_COUNT == 2 */
    *start = 2 - (odd ? 1 : 0);
    *stop = ncoord - 2 - (odd ? 0 : 1);
}

static void e1d_so0(int64_t ncoord, double const * coord, double const * gamma, double * so0,
                    double const * so1, double hdt, double qdt, int odd)
{
    /* Euler1DCore::march_half_so0, Euler1DCore.cpp:96-130 */
    int64_t start, stop;
    e1d_range(ncoord, odd, &start, &stop);
    for (int64_t ic = start; ic < stop; ic += 2)
    {
        e1d_se sn, sp;
        double fll[3], flr[3];
        e1d_derive(&sn, ic, gamma, coord, so0, so1, hdt);
        e1d_derive(&sp, ic + 2, gamma, coord, so0, so1, hdt);
        e1d_flux(&sn, hdt, qdt, 0, fll);
        e1d_flux(&sp, hdt, qdt, 1, flr);
        double const dx = coord[ic + 2] - coord[ic];
        for (int i = 0; i < 3; ++i) { so0[(ic + 1) * 3 + i] = (fll[i] + flr[i]) / dx; }
    }
}

static void e1d_cfl(int64_t ncoord, double const * coord, double const * gamma, double const * so0,
                    double * cfl, double dt, int odd)
{
    /* Euler1DCore::update_cfl, Euler1DCore.cpp:70-94 */
    int64_t start, stop;
    e1d_range(ncoord, odd, &start, &stop);
    double const hdt = dt / 2;
    for (int64_t it = start; it < stop; it += 2)
    {
        double const ga = gamma[it];
        double wspd = so0[it * 3 + 1];
        wspd *= wspd;
        double const ke = wspd / (2.0 * so0[it * 3 + 0]);
        double pr = (ga - 1.0) * (so0[it * 3 + 2] - ke);
        pr = (pr + fabs(pr)) / 2.0;
        wspd = sqrt(ga * pr / so0[it * 3 + 0]) + sqrt(wspd) / so0[it * 3 + 0];
        double const dxpos = coord[it + 1] - coord[it];
        double const dxneg = coord[it] - coord[it - 1];
        cfl[it] = hdt * wspd / (dxpos < dxneg ? dxpos : dxneg);
    }
}

static void e1d_so1(int64_t ncoord, double const * coord, double const * gamma, double const * so0,
                    double * so1, double hdt, int odd, int alpha)
{
    /* Euler1DCore::march_half_so1_alpha, Euler1DCore.hpp:284-320 */
    int64_t start, stop;
    e1d_range(ncoord, odd, &start, &stop);
    for (int64_t ic = start; ic < stop; ic += 2)
    {
        e1d_se sn, sp;
        e1d_derive(&sn, ic, gamma, coord, so0, so1, hdt);
        e1d_derive(&sp, ic + 2, gamma, coord, so0, so1, hdt);
        for (int iv = 0; iv < 3; ++iv)
        {
            double const utp = so0[(ic + 1) * 3 + iv];
            double const duxn = (utp - sn.up[iv]) / (sn.xpos - sn.x);
            double const duxp = (sp.up[iv] - utp) / (sp.x - sp.xneg);
            double const fan = pow_alpha(fabs(duxn), alpha);
            double const fap = pow_alpha(fabs(duxp), alpha);
            so1[(ic + 1) * 3 + iv] = (fap * duxn + fan * duxp) / (fap + fan + E1D_TINY);
        }
    }
}

static void e1d_bc(int64_t ncoord, double * a)
{
    /* treat_boundary_so0/so1, Euler1DCore.cpp:132-172 */
    for (int i = 0; i < 3; ++i)
    {
        a[1 * 3 + i] = a[2 * 3 + i];
        a[(ncoord - 2) * 3 + i] = a[(ncoord - 3) * 3 + i];
    }
}

int orc_e1d_phase(int64_t ncoord, double const * coord, double const * gamma, double * so0, double * so1,
                  double * cfl, double dt, int phase, int odd, int alpha)
{
    double const hdt = dt / 2.0; /* Euler1DCore.hpp:186-191 */
    double const qdt = hdt / 2.0;
    if (alpha < 1 || alpha > 2) { return -1; }
    switch (phase)
    {
    case 0: e1d_so0(ncoord, coord, gamma, so0, so1, hdt, qdt, odd); break;
    case 1: e1d_cfl(ncoord, coord, gamma, so0, cfl, dt, odd); break;
    case 2: e1d_so1(ncoord, coord, gamma, so0, so1, hdt, odd, alpha); break;
    case 3: e1d_bc(ncoord, so0); break;
    case 4: e1d_bc(ncoord, so1); break;
    default: return -1;
    }
    return 0;
}

int orc_e1d_march(int64_t ncoord, double const * coord, double const * gamma, double * so0, double * so1,
                  double * cfl, double dt, int alpha, int64_t steps)
{
    double const hdt = dt / 2.0;
    double const qdt = hdt / 2.0;
    if (alpha < 1 || alpha > 2) { return -1; }
    for (int64_t it = 0; it < steps; ++it)
    {
        /* march_alpha, Euler1DCore.hpp:322-349 */
        e1d_so0(ncoord, coord, gamma, so0, so1, hdt, qdt, 0);
        e1d_cfl(ncoord, coord, gamma, so0, cfl, dt, 0);
        e1d_so1(ncoord, coord, gamma, so0, so1, hdt, 0, alpha);
        e1d_bc(ncoord, so0);
        e1d_bc(ncoord, so1);
        e1d_so0(ncoord, coord, gamma, so0, so1, hdt, qdt, 1);
        e1d_cfl(ncoord, coord, gamma, so0, cfl, dt, 1);
        e1d_so1(ncoord, coord, gamma, so0, so1, hdt, 1, alpha);
    }
    return 0;
}

/* ====================================================================== */
/* multidim Euler                                                         */
/* ====================================================================== */

#define MD_TINY 1.e-60          /* euler_march.cpp:30,89; euler_solution.cpp:108 */
#define MD_ALMOST_ZERO 1.e-200  /* euler_gradient.cpp:28 */
#define CLMFC 6
#define FCMND 4

typedef struct md_jaco
{
    double jacos[5][5][3];
    double fcn[5][3];
} md_jaco;

/* detail::EulerJacobian<2>::update, euler_march.cpp:35-81 */
static void jaco2(md_jaco * j, double gamma, double const * sol)
{
    memset(j, 0, sizeof(*j));
    double const ga = gamma;
    double const ga1 = ga - 1;
    double const ga3 = ga - 3;
    double const ga1h = ga1 / 2;
    double const u1 = sol[0] + MD_TINY;
    double const u2 = sol[1];
    double const u3 = sol[2];
    double const u4 = sol[3];
    double const rho2 = u1 * u1;
    double const v1 = u2 / u1;
    double const v1o2 = v1 * v1;
    double const v2 = u3 / u1;
    double const v2o2 = v2 * v2;
    double const ke2 = (u2 * u2 + u3 * u3) / u1;
    double const g1ke2 = ga1 * ke2;
    double const vs = ke2 / u1;
    double const gretot = ga * u4;
    double const getot = gretot / u1;
    double const pr = ga1 * u4 - ga1h * ke2;

    j->fcn[0][0] = u2;              j->fcn[0][1] = u3;
    j->fcn[1][0] = pr + u2 * v1;    j->fcn[1][1] = u2 * v2;
    j->fcn[2][0] = u3 * v1;         j->fcn[2][1] = pr + u3 * v2;
    j->fcn[3][0] = (pr + u4) * v1;  j->fcn[3][1] = (pr + u4) * v2;

    j->jacos[0][1][0] = 1;
    j->jacos[0][2][1] = 1;

    j->jacos[1][0][0] = -v1o2 + ga1h * vs;  j->jacos[1][0][1] = -v1 * v2;
    j->jacos[1][1][0] = -ga3 * v1;          j->jacos[1][1][1] = v2;
    j->jacos[1][2][0] = -ga1 * v2;          j->jacos[1][2][1] = v1;
    j->jacos[1][3][0] = ga1;

    j->jacos[2][0][0] = -v2 * v1;           j->jacos[2][0][1] = -v2o2 + ga1h * vs;
    j->jacos[2][1][0] = v2;                 j->jacos[2][1][1] = -ga1 * v1;
    j->jacos[2][2][0] = v1;                 j->jacos[2][2][1] = -ga3 * v2;
    j->jacos[2][3][1] = ga1;

    j->jacos[3][0][0] = (-gretot + g1ke2) * u2 / rho2;
    j->jacos[3][0][1] = (-gretot + g1ke2) * u3 / rho2;
    j->jacos[3][1][0] = getot - ga1h * (vs + 2 * v1o2);
    j->jacos[3][1][1] = -ga1 * v1 * v2;
    j->jacos[3][2][0] = -ga1 * v2 * v1;
    j->jacos[3][2][1] = getot - ga1h * (vs + 2 * v2o2);
    j->jacos[3][3][0] = ga * v1;
    j->jacos[3][3][1] = ga * v2;
}

/* detail::EulerJacobian<3>::update, euler_march.cpp:94-157 */
static void jaco3(md_jaco * j, double gamma, double const * sol)
{
    memset(j, 0, sizeof(*j));
    double const ga = gamma;
    double const ga1 = ga - 1;
    double const ga3 = ga - 3;
    double const ga1h = ga1 / 2;
    double const u1 = sol[0] + MD_TINY;
    double const u2 = sol[1];
    double const u3 = sol[2];
    double const u4 = sol[3];
    double const u5 = sol[4];
    double const rho2 = u1 * u1;
    double const v1 = u2 / u1;
    double const v1o2 = v1 * v1;
    double const v2 = u3 / u1;
    double const v2o2 = v2 * v2;
    double const v3 = u4 / u1;
    double const v3o2 = v3 * v3;
    double const ke2 = (u2 * u2 + u3 * u3 + u4 * u4) / u1;
    double const g1ke2 = ga1 * ke2;
    double const vs = ke2 / u1;
    double const gretot = ga * u5;
    double const getot = gretot / u1;
    double const pr = ga1 * u5 - ga1h * ke2;

    double (*f)[3] = j->fcn;
    f[0][0] = u2;             f[0][1] = u3;             f[0][2] = u4;
    f[1][0] = pr + u2 * v1;   f[1][1] = u2 * v2;        f[1][2] = u2 * v3;
    f[2][0] = u3 * v1;        f[2][1] = pr + u3 * v2;   f[2][2] = u3 * v3;
    f[3][0] = u4 * v1;        f[3][1] = u4 * v2;        f[3][2] = pr + u4 * v3;
    f[4][0] = (pr + u5) * v1; f[4][1] = (pr + u5) * v2; f[4][2] = (pr + u5) * v3;

    double (*J)[5][3] = j->jacos;
    J[0][1][0] = 1;
    J[0][2][1] = 1;
    J[0][3][2] = 1;

    J[1][0][0] = -v1o2 + ga1h * vs; J[1][0][1] = -v1 * v2; J[1][0][2] = -v1 * v3;
    J[1][1][0] = -ga3 * v1;         J[1][1][1] = v2;       J[1][1][2] = v3;
    J[1][2][0] = -ga1 * v2;         J[1][2][1] = v1;
    J[1][3][0] = -ga1 * v3;                                J[1][3][2] = v1;
    J[1][4][0] = ga1;

    J[2][0][0] = -v2 * v1; J[2][0][1] = -v2o2 + ga1h * vs; J[2][0][2] = -v2 * v3;
    J[2][1][0] = v2;       J[2][1][1] = -ga1 * v1;
    J[2][2][0] = v1;       J[2][2][1] = -ga3 * v2;         J[2][2][2] = v3;
                           J[2][3][1] = -ga1 * v3;         J[2][3][2] = v2;
                           J[2][4][1] = ga1;

    J[3][0][0] = -v3 * v1; J[3][0][1] = -v3 * v2; J[3][0][2] = -v3o2 + ga1h * vs;
    J[3][1][0] = v3;                              J[3][1][2] = -ga1 * v1;
                           J[3][2][1] = v3;       J[3][2][2] = -ga1 * v2;
    J[3][3][0] = v1;       J[3][3][1] = v2;       J[3][3][2] = -ga3 * v3;
                                                  J[3][4][2] = ga1;

    J[4][0][0] = (-gretot + g1ke2) * u2 / rho2;
    J[4][0][1] = (-gretot + g1ke2) * u3 / rho2;
    J[4][0][2] = (-gretot + g1ke2) * u4 / rho2;
    J[4][1][0] = getot - ga1h * (vs + 2 * v1o2); J[4][1][1] = -ga1 * v1 * v2; J[4][1][2] = -ga1 * v1 * v3;
    J[4][2][0] = -ga1 * v2 * v1; J[4][2][1] = getot - ga1h * (vs + 2 * v2o2); J[4][2][2] = -ga1 * v2 * v3;
    J[4][3][0] = -ga1 * v3 * v1; J[4][3][1] = -ga1 * v3 * v2; J[4][3][2] = getot - ga1h * (vs + 2 * v3o2);
    J[4][4][0] = ga * v1; J[4][4][1] = ga * v2; J[4][4][2] = ga * v3;
}

static void jaco(md_jaco * j, int ndim, double gamma, double const * sol)
{
    if (2 == ndim) { jaco2(j, gamma, sol); }
    else { jaco3(j, gamma, sol); }
}

static int32_t fcrcl(orc_mesh const * m, int32_t ifc, int32_t icl) /* StaticMesh.hpp:355-358 */
{
    return m->fccls[ifc * 4 + 0] + m->fccls[ifc * 4 + 1] - icl;
}

/* Neighbour solution point: CE centroid for body cells, mesh cell centroid for
 * ghost cells (euler_march.cpp:229-234, GradientElement.cpp:176). */
static double solpt(orc_md const * md, int32_t jcl, int d)
{
    int const ndim = md->mesh->ndim;
    return (jcl >= 0) ? md->cecnd[jcl * (CLMFC + 1) * ndim + d] : md->mesh->clcnd[jcl * ndim + d];
}

/* EulerCore::prepare_ce_2d, euler_ce.cpp:78-176 */
static void md_prepare_ce_2d(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    for (int32_t icl = 0; icl < m->ncell; ++icl)
    {
        int32_t const clnfc = m->clfcs[icl * 7];
        double volc = 0.0, cc0 = 0.0, cc1 = 0.0;
        double const crdi0 = m->clcnd[icl * 2 + 0];
        double const crdi1 = m->clcnd[icl * 2 + 1];
        for (int32_t ifl = 1; ifl <= clnfc; ++ifl)
        {
            int32_t const ifc = m->clfcs[icl * 7 + ifl];
            int32_t const jcl = fcrcl(m, ifc, icl);
            double const crde0 = m->clcnd[jcl * 2 + 0];
            double const crde1 = m->clcnd[jcl * 2 + 1];
            int32_t const ind0 = m->fcnds[ifc * 5 + 1];
            int32_t const ind1 = m->fcnds[ifc * 5 + 2];
            double const crd10 = m->ndcrd[ind0 * 2 + 0];
            double const crd11 = m->ndcrd[ind0 * 2 + 1];
            double const crd20 = m->ndcrd[ind1 * 2 + 0];
            double const crd21 = m->ndcrd[ind1 * 2 + 1];

            double const cndi0 = (crd10 + crd20 + crdi0) / 3.0;
            double const cndi1 = (crd11 + crd21 + crdi1) / 3.0;
            double const voli = fabs((crd10 - crdi0) * (crd21 - crdi1) - (crd11 - crdi1) * (crd20 - crdi0)) / 2.0;
            double const cnde0 = (crd10 + crd20 + crde0) / 3.0;
            double const cnde1 = (crd11 + crd21 + crde1) / 3.0;
            double const cross = (crd10 - crde0) * (crd21 - crde1) - (crd11 - crde1) * (crd20 - crde0);
            double const vole = fabs(cross) / 2.0;

            double const volb = voli + vole;
            md->cevol[icl * 7 + ifl] = volb;
            md->cecnd[icl * 14 + ifl * 2 + 0] = (cndi0 * voli + cnde0 * vole) / volb;
            md->cecnd[icl * 14 + ifl * 2 + 1] = (cndi1 * voli + cnde1 * vole) / volb;
            volc += volb;
            cc0 += (cndi0 * voli + cnde0 * vole);
            cc1 += (cndi1 * voli + cnde1 * vole);

            double const sgn = (cross > 0.0) ? 1.0 : -1.0;
            int64_t const s0 = ((int64_t)icl * 24 + (ifl - 1) * FCMND) * 2;
            md->sfcnd[s0 + 0] = (crde0 + crd10) / 2.0;
            md->sfcnd[s0 + 1] = (crde1 + crd11) / 2.0;
            double const a_dx = crde0 - crd10;
            double const a_dy = crde1 - crd11;
            md->sfnml[s0 + 0] = -a_dy * sgn;
            md->sfnml[s0 + 1] = a_dx * sgn;
            md->sfcnd[s0 + 2] = (crd20 + crde0) / 2.0;
            md->sfcnd[s0 + 3] = (crd21 + crde1) / 2.0;
            double const b_dx = crd20 - crde0;
            double const b_dy = crd21 - crde1;
            md->sfnml[s0 + 2] = -b_dy * sgn;
            md->sfnml[s0 + 3] = b_dx * sgn;
        }
        md->cevol[icl * 7] = volc;
        md->cecnd[icl * 14 + 0] = cc0 / volc;
        md->cecnd[icl * 14 + 1] = cc1 / volc;
    }
}

/* EulerCore::prepare_ce_3d, euler_ce.cpp:178-347 */
static void md_prepare_ce_3d(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    for (int32_t icl = 0; icl < m->ncell; ++icl)
    {
        int32_t const clnfc = m->clfcs[icl * 7];
        double volc = 0.0;
        double cc[3] = {0.0, 0.0, 0.0};
        double const * ci = &m->clcnd[icl * 3];
        for (int32_t ifl = 1; ifl <= clnfc; ++ifl)
        {
            int32_t const ifc = m->clfcs[icl * 7 + ifl];
            int32_t const jcl = fcrcl(m, ifc, icl);
            int32_t const fcnnd = m->fcnds[ifc * 5];
            double const * ce = &m->clcnd[jcl * 3];
            double const * fc = &m->fccnd[ifc * 3];
            double crd[FCMND + 2][3];
            for (int32_t inf = 1; inf <= fcnnd; ++inf)
            {
                int32_t const ind = m->fcnds[ifc * 5 + inf];
                for (int d = 0; d < 3; ++d) { crd[inf][d] = m->ndcrd[ind * 3 + d]; }
            }
            for (int d = 0; d < 3; ++d) { crd[fcnnd + 1][d] = crd[1][d]; }

            double volb = 0.0;
            double bc[3] = {0.0, 0.0, 0.0};
            double const voe = (m->fccls[ifc * 4] == icl) ? 0.5 : -0.5;
            for (int32_t inf = 1; inf <= fcnnd; ++inf)
            {
                double du[3], dv[3], dt_[3], wi[3], we[3];
                for (int d = 0; d < 3; ++d)
                {
                    du[d] = crd[inf][d] - fc[d];
                    dv[d] = crd[inf + 1][d] - fc[d];
                }
                dt_[0] = du[1] * dv[2] - du[2] * dv[1];
                dt_[1] = du[2] * dv[0] - du[0] * dv[2];
                dt_[2] = du[0] * dv[1] - du[1] * dv[0];
                for (int d = 0; d < 3; ++d)
                {
                    wi[d] = ci[d] - fc[d];
                    we[d] = ce[d] - fc[d];
                }
                double const voli = fabs(dt_[0] * wi[0] + dt_[1] * wi[1] + dt_[2] * wi[2]) / 6.0;
                double const vole = fabs(dt_[0] * we[0] + dt_[1] * we[1] + dt_[2] * we[2]) / 6.0;
                volb += voli + vole;
                for (int d = 0; d < 3; ++d)
                {
                    double const cndi = (crd[inf][d] + crd[inf + 1][d] + fc[d] + ci[d]) / 4.0;
                    double const cnde = (crd[inf][d] + crd[inf + 1][d] + fc[d] + ce[d]) / 4.0;
                    bc[d] += cndi * voli + cnde * vole;
                }
                int64_t const s = ((int64_t)icl * 24 + (ifl - 1) * FCMND + (inf - 1)) * 3;
                double su[3], sv[3];
                for (int d = 0; d < 3; ++d)
                {
                    md->sfcnd[s + d] = (ce[d] + crd[inf][d] + crd[inf + 1][d]) / 3.0;
                    su[d] = crd[inf][d] - ce[d];
                    sv[d] = crd[inf + 1][d] - ce[d];
                }
                md->sfnml[s + 0] = (su[1] * sv[2] - su[2] * sv[1]) * voe;
                md->sfnml[s + 1] = (su[2] * sv[0] - su[0] * sv[2]) * voe;
                md->sfnml[s + 2] = (su[0] * sv[1] - su[1] * sv[0]) * voe;
            }
            md->cevol[icl * 7 + ifl] = volb;
            for (int d = 0; d < 3; ++d)
            {
                md->cecnd[icl * 21 + ifl * 3 + d] = bc[d] / volb;
                cc[d] += bc[d];
            }
            volc += volb;
        }
        md->cevol[icl * 7] = volc;
        for (int d = 0; d < 3; ++d) { md->cecnd[icl * 21 + d] = cc[d] / volc; }
    }
}

/* calc_solt_impl, euler_march.cpp:161-191 */
static void md_solt(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    int const ndim = m->ndim;
    int const neq = ndim + 2;
    md_jaco jc;
    for (int32_t icl = -m->ngstcell; icl < m->ncell; ++icl)
    {
        jaco(&jc, ndim, md->gamma[icl], &md->so0c[icl * neq]);
        for (int ieq = 0; ieq < neq; ++ieq)
        {
            double val = 0.0;
            for (int idm = 0; idm < ndim; ++idm)
            {
                for (int jeq = 0; jeq < neq; ++jeq)
                {
                    val += jc.jacos[ieq][jeq][idm] * md->so1c[(icl * neq + jeq) * ndim + idm];
                }
            }
            md->so0t[icl * neq + ieq] = -val;
        }
    }
}

/* calc_soln_impl, euler_march.cpp:197-303 */
static void md_soln(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    int const ndim = m->ndim;
    int const neq = ndim + 2;
    double const qdt = md->dt * 0.25;
    double const hdt = md->dt * 0.5;
    md_jaco jc;
    for (int32_t icl = 0; icl < m->ncell; ++icl)
    {
        int32_t const clnfc = m->clfcs[icl * 7];
        double acc[5] = {0, 0, 0, 0, 0};
        for (int32_t ifl = 1; ifl <= clnfc; ++ifl)
        {
            int32_t const ifc = m->clfcs[icl * 7 + ifl];
            int32_t const jcl = fcrcl(m, ifc, icl);
            double jcecnd[3] = {0, 0, 0}, bcnd[3] = {0, 0, 0};
            for (int d = 0; d < ndim; ++d)
            {
                jcecnd[d] = solpt(md, jcl, d);
                bcnd[d] = md->cecnd[icl * 7 * ndim + ifl * ndim + d];
            }
            double const bvol = md->cevol[icl * 7 + ifl];
            double const * ju = &md->so0c[jcl * neq];
            double const * jg = &md->so1c[jcl * neq * ndim];
            double const * jt = &md->so0t[jcl * neq];
            for (int ieq = 0; ieq < neq; ++ieq)
            {
                double fusp = ju[ieq];
                for (int d = 0; d < ndim; ++d) { fusp += (bcnd[d] - jcecnd[d]) * jg[ieq * ndim + d]; }
                acc[ieq] += fusp * bvol;
            }
            jaco(&jc, ndim, md->gamma[jcl], ju);
            int32_t const fcnnd = m->fcnds[ifc * 5];
            for (int32_t inf = 0; inf < fcnnd; ++inf)
            {
                int64_t const s = ((int64_t)icl * 24 + (ifl - 1) * FCMND + inf) * ndim;
                double usfc[5];
                for (int ieq = 0; ieq < neq; ++ieq)
                {
                    usfc[ieq] = qdt * jt[ieq];
                    for (int d = 0; d < ndim; ++d)
                    {
                        usfc[ieq] += (md->sfcnd[s + d] - jcecnd[d]) * jg[ieq * ndim + d];
                    }
                }
                for (int ieq = 0; ieq < neq; ++ieq)
                {
                    double dot = 0.0;
                    for (int d2 = 0; d2 < ndim; ++d2)
                    {
                        double dfcn = jc.fcn[ieq][d2];
                        for (int jeq = 0; jeq < neq; ++jeq) { dfcn += jc.jacos[ieq][jeq][d2] * usfc[jeq]; }
                        dot += dfcn * md->sfnml[s + d2];
                    }
                    acc[ieq] -= hdt * dot;
                }
            }
        }
        double const cvol = md->cevol[icl * 7];
        for (int ieq = 0; ieq < neq; ++ieq) { md->so0n[icl * neq + ieq] = acc[ieq] / cvol; }
    }
}

/* EulerCore::calc_cfl, euler_solution.cpp:106-160 */
static void md_cfl(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    int const ndim = m->ndim;
    int const neq = ndim + 2;
    double const hdt = md->dt / 2.0;
    for (int32_t icl = 0; icl < m->ncell; ++icl)
    {
        int32_t const clnfc = m->clfcs[icl * 7];
        double dist = DBL_MAX;
        double const * cc = &md->cecnd[icl * 7 * ndim];
        for (int32_t ifl = 1; ifl <= clnfc; ++ifl)
        {
            double d2 = 0.0;
            for (int d = 0; d < ndim; ++d)
            {
                double const diff = cc[ifl * ndim + d] - cc[d];
                d2 += diff * diff;
            }
            double const r = sqrt(d2);
            dist = (r < dist) ? r : dist; /* std::min(dist, r) */
        }
        double const ga = md->gamma[icl];
        double const ga1 = ga - 1.0;
        double * u = &md->so0n[icl * neq];
        double const density = u[0];
        double momsq = 0.0;
        for (int d = 0; d < ndim; ++d) { momsq += u[1 + d] * u[1 + d]; }
        double const ke = momsq / (2.0 * density);
        double const energy = u[neq - 1];
        double const pr = ga1 * (energy - ke);
        double const pr_adj = (pr + fabs(pr)) / 2.0;
        double const wspd = sqrt(ga * pr_adj / density) + sqrt(momsq) / density;
        double const cflo = hdt * wspd / dist;
        md->cflo[icl] = cflo;
        md->cflc[icl] = (cflo - 1.0) * pr_adj / (pr_adj + MD_TINY) + 1.0;
        u[neq - 1] = pr_adj / ga1 + ke + MD_TINY;
    }
}

/* GradientElementTypeGroup tables, GradientElement.cpp:16-79 (1-based face ids) */
typedef struct fge_type
{
    int nfge;
    int faces[8][3];
} fge_type;

static fge_type const FGE[9] = {
    {0, {{0}}},
    {0, {{0}}},
    {0, {{0}}},
    /* 3 quadrilateral */ {4, {{1, 2, -1}, {2, 3, -1}, {3, 4, -1}, {4, 1, -1}}},
    /* 4 triangle      */ {3, {{1, 2, -1}, {2, 3, -1}, {3, 1, -1}}},
    /* 5 hexahedron    */ {8, {{2, 3, 5}, {6, 3, 2}, {4, 3, 6}, {5, 3, 4}, {5, 1, 2}, {2, 1, 6}, {6, 1, 4}, {4, 1, 5}}},
    /* 6 tetrahedron   */ {4, {{3, 1, 2}, {2, 1, 4}, {4, 1, 3}, {2, 4, 3}}},
    /* 7 prism         */ {6, {{5, 2, 4}, {3, 2, 5}, {4, 2, 3}, {4, 1, 5}, {5, 1, 3}, {3, 1, 4}}},
    /* 8 pyramid       */ {6, {{1, 5, 2}, {2, 5, 3}, {3, 5, 4}, {4, 5, 1}, {1, 3, 4}, {3, 1, 2}}},
};

static double det_n(double a[3][3], int ndim) /* GradientElement.cpp:215-226 */
{
    if (2 == ndim) { return a[0][0] * a[1][1] - a[0][1] * a[1][0]; }
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1])
         - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
         + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}

static void adj_n(double a[3][3], int ndim, double r[3][3]) /* GradientElement.cpp:228-249 */
{
    memset(r, 0, 9 * sizeof(double));
    if (2 == ndim)
    {
        r[0][0] = a[1][1];
        r[0][1] = -a[0][1];
        r[1][0] = -a[1][0];
        r[1][1] = a[0][0];
        return;
    }
    r[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    r[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    r[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    r[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    r[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    r[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    r[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    r[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    r[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
}

/* EulerCore::calc_dsoln, euler_gradient.cpp:23-130, with the GradientElement
 * constructor (GradientElement.cpp:137-213) and calc_gge_centroid (:81-133)
 * expanded in place. */
static void md_dsoln(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    int const ndim = m->ndim;
    int const neq = ndim + 2;
    double const hdt = md->dt * 0.5;
    for (int32_t icl = 0; icl < m->ncell; ++icl)
    {
        double const acfl = fabs(md->cflc[icl]);
        double const sgm0 = md->sigma0 / acfl;
        double const tau = md->taumin + acfl * md->tauscale;
        fge_type const * ge = &FGE[m->cltpn[icl]];
        int const nfge = ge->nfge;
        double const ofg1 = 1.0 / nfge;
        int32_t const clnfc = m->clfcs[icl * 7];

        /* --- gradient element geometry --- */
        int32_t rcl[CLMFC];
        double gp[CLMFC][3], jd[CLMFC][3], idis[CLMFC][3], jdis[CLMFC][3];
        double icecnd[3] = {0, 0, 0};
        memset(gp, 0, sizeof(gp));
        memset(jd, 0, sizeof(jd));
        for (int d = 0; d < ndim; ++d) { icecnd[d] = md->cecnd[icl * 7 * ndim + d]; }
        for (int32_t ifl = 0; ifl < clnfc; ++ifl)
        {
            int32_t const ifc = m->clfcs[icl * 7 + ifl + 1];
            int32_t const jcl = fcrcl(m, ifc, icl);
            rcl[ifl] = jcl;
            for (int d = 0; d < ndim; ++d)
            {
                double const mid = md->cecnd[icl * 7 * ndim + (ifl + 1) * ndim + d];
                double const jce = solpt(md, jcl, d);
                gp[ifl][d] = mid + tau * (jce - mid);
                jd[ifl][d] = gp[ifl][d] - jce;
            }
        }
        double avg[3] = {0, 0, 0};
        for (int32_t ifl = 0; ifl < clnfc; ++ifl)
        {
            for (int d = 0; d < ndim; ++d) { avg[d] += gp[ifl][d]; }
        }
        for (int d = 0; d < ndim; ++d) { avg[d] /= clnfc; }
        double cnd[3] = {0, 0, 0};
        double voc = 0.0;
        for (int isub = 0; isub < nfge; ++isub)
        {
            double subcnd[3] = {avg[0], avg[1], avg[2]};
            double dst[3][3];
            memset(dst, 0, sizeof(dst));
            for (int ivx = 0; ivx < ndim; ++ivx)
            {
                int const fl = ge->faces[isub][ivx] - 1;
                for (int d = 0; d < ndim; ++d)
                {
                    subcnd[d] += gp[fl][d];
                    dst[ivx][d] = gp[fl][d] - avg[d];
                }
            }
            for (int d = 0; d < ndim; ++d) { subcnd[d] /= (double)(ndim + 1); }
            double const vob = det_n(dst, ndim);
            voc += vob;
            for (int d = 0; d < ndim; ++d) { cnd[d] += subcnd[d] * vob; }
        }
        for (int d = 0; d < ndim; ++d) { cnd[d] /= voc; }
        for (int32_t ifl = 0; ifl < clnfc; ++ifl)
        {
            for (int d = 0; d < ndim; ++d)
            {
                idis[ifl][d] = gp[ifl][d] - cnd[d];
                jdis[ifl][d] = jd[ifl][d] + icecnd[d] - cnd[d];
            }
        }

        /* --- per-FGE gradients and W-1/2 weights --- */
        double grad[8][5][3], widv[8][5], wacc[5] = {0, 0, 0, 0, 0}, sigma_max[5];
        memset(grad, 0, sizeof(grad));
        for (int ifge = 0; ifge < nfge; ++ifge)
        {
            double udf[5][3];
            double dst[3][3], adj[3][3];
            memset(udf, 0, sizeof(udf));
            memset(dst, 0, sizeof(dst));
            for (int ivx = 0; ivx < ndim; ++ivx)
            {
                int const ifl = ge->faces[ifge][ivx] - 1;
                int32_t const jcl = rcl[ifl];
                for (int ieq = 0; ieq < neq; ++ieq)
                {
                    double val = md->so0c[jcl * neq + ieq] + hdt * md->so0t[jcl * neq + ieq] - md->so0n[icl * neq + ieq];
                    for (int d = 0; d < ndim; ++d) { val += jdis[ifl][d] * md->so1c[(jcl * neq + ieq) * ndim + d]; }
                    udf[ieq][ivx] = val;
                }
                for (int d = 0; d < ndim; ++d) { dst[ivx][d] = idis[ifl][d]; }
            }
            adj_n(dst, ndim, adj);
            double const det = det_n(dst, ndim);
            for (int ieq = 0; ieq < neq; ++ieq)
            {
                double sq = 0.0;
                for (int i = 0; i < ndim; ++i)
                {
                    double r = 0.0;
                    for (int j = 0; j < ndim; ++j) { r += adj[i][j] * udf[ieq][j]; }
                    double const g = r / det;
                    grad[ifge][ieq][i] = g;
                }
                for (int d = 0; d < ndim; ++d) { sq += grad[ifge][ieq][d] * grad[ifge][ieq][d]; }
                double const wgt = 1.0 / sqrt(sq + MD_ALMOST_ZERO);
                wacc[ieq] += wgt;
                widv[ifge][ieq] = wgt;
            }
        }

        /* --- W-3/4 limiter --- */
        double wpa[5][2];
        memset(wpa, 0, sizeof(wpa));
        for (int ifge = 0; ifge < nfge; ++ifge)
        {
            for (int ieq = 0; ieq < neq; ++ieq)
            {
                double const wgt = widv[ifge][ieq] / wacc[ieq] - ofg1;
                widv[ifge][ieq] = wgt;
                wpa[ieq][0] = fmax(wpa[ieq][0], wgt);
                wpa[ieq][1] = fmin(wpa[ieq][1], wgt);
            }
        }
        for (int ieq = 0; ieq < neq; ++ieq)
        {
            double const sm = fmin((1.0 - ofg1) / (wpa[ieq][0] + MD_ALMOST_ZERO),
                                   -ofg1 / (wpa[ieq][1] - MD_ALMOST_ZERO));
            sigma_max[ieq] = fmin(sm, sgm0);
        }
        double * out = &md->so1n[icl * neq * ndim];
        for (int k = 0; k < neq * ndim; ++k) { out[k] = 0.0; }
        for (int isub = 0; isub < nfge; ++isub)
        {
            for (int ieq = 0; ieq < neq; ++ieq)
            {
                double const wgt = ofg1 + sigma_max[ieq] * widv[isub][ieq];
                for (int d = 0; d < ndim; ++d) { out[ieq * ndim + d] += wgt * grad[isub][ieq][d]; }
            }
        }
    }
}

/* normal_matrix<2>/<3>, euler_bc.cpp:90-128 */
void orc_normal_matrix(int ndim, double const * n, double * out)
{
    if (2 == ndim)
    {
        out[0] = n[0];
        out[1] = n[1];
        out[2] = -n[1];
        out[3] = n[0];
        return;
    }
    double axis[3] = {0.0, 0.0, 0.0};
    double const ax = fabs(n[0]), ay = fabs(n[1]), az = fabs(n[2]);
    if (ax <= ay && ax <= az) { axis[0] = 1.0; }
    else if (ay <= az) { axis[1] = 1.0; }
    else { axis[2] = 1.0; }
    double t1[3] = {n[1] * axis[2] - n[2] * axis[1], n[2] * axis[0] - n[0] * axis[2], n[0] * axis[1] - n[1] * axis[0]};
    double const len = sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
    t1[0] = t1[0] / len;
    t1[1] = t1[1] / len;
    t1[2] = t1[2] / len;
    double const t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
    for (int d = 0; d < 3; ++d)
    {
        out[d] = n[d];
        out[3 + d] = t1[d];
        out[6 + d] = t2[d];
    }
}

static void matvec(int n, double const * m, double const * v, double * r) /* euler_bc.cpp:38-51 */
{
    for (int i = 0; i < n; ++i)
    {
        r[i] = 0.0;
        for (int j = 0; j < n; ++j) { r[i] += m[i * n + j] * v[j]; }
    }
}
static void transpose(int n, double const * m, double * r)
{
    for (int i = 0; i < n; ++i)
    {
        for (int j = 0; j < n; ++j) { r[i * n + j] = m[j * n + i]; }
    }
}
static void matmat(int n, double const * a, double const * b, double * r) /* euler_bc.cpp:68-83 */
{
    for (int i = 0; i < n; ++i)
    {
        for (int j = 0; j < n; ++j)
        {
            r[i * n + j] = 0.0;
            for (int k = 0; k < n; ++k) { r[i * n + j] += a[i * n + k] * b[k * n + j]; }
        }
    }
}

/* bc_soln_impl, euler_bc.cpp:143-215 */
static void md_bc_soln(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    int const ndim = m->ndim;
    int const neq = ndim + 2;
    for (int32_t ib = 0; ib < md->nbnd; ++ib)
    {
        for (int32_t k = 0; k < md->bnd_nface[ib]; ++k)
        {
            int32_t const ifc = md->bnd_faces[ib][k];
            int32_t const icl = m->fccls[ifc * 4 + 0];
            int32_t const jcl = m->fccls[ifc * 4 + 1];
            double const * ui = &md->so0n[icl * neq];
            double * uj = &md->so0n[jcl * neq];
            if (0 == md->bnd_kind[ib])
            {
                for (int ieq = 0; ieq < neq; ++ieq) { uj[ieq] = ui[ieq]; }
            }
            else if (1 == md->bnd_kind[ib])
            {
                double mat[9], inv[9], mom[3], rot[3], out[3];
                orc_normal_matrix(ndim, &m->fcnml[ifc * ndim], mat);
                transpose(ndim, mat, inv);
                uj[0] = ui[0];
                uj[neq - 1] = ui[neq - 1];
                for (int d = 0; d < ndim; ++d) { mom[d] = ui[1 + d]; }
                matvec(ndim, mat, mom, rot);
                rot[0] = -rot[0];
                matvec(ndim, inv, rot, out);
                for (int d = 0; d < ndim; ++d) { uj[1 + d] = out[d]; }
            }
            else
            {
                double const * v = md->bnd_value[ib];
                double const rho = v[0];
                double const p = v[1 + ndim];
                double const ga = v[2 + ndim];
                double vsq = 0.0;
                for (int d = 0; d < ndim; ++d) { vsq += v[1 + d] * v[1 + d]; }
                uj[0] = rho;
                for (int d = 0; d < ndim; ++d) { uj[1 + d] = rho * v[1 + d]; }
                uj[neq - 1] = p / (ga - 1.0) + 0.5 * rho * vsq;
            }
        }
    }
}

/* bc_dsoln_impl, euler_bc.cpp:219-320 */
static void md_bc_dsoln(orc_md * md)
{
    orc_mesh const * m = md->mesh;
    int const ndim = m->ndim;
    int const neq = ndim + 2;
    for (int32_t ib = 0; ib < md->nbnd; ++ib)
    {
        for (int32_t k = 0; k < md->bnd_nface[ib]; ++k)
        {
            int32_t const ifc = md->bnd_faces[ib][k];
            int32_t const icl = m->fccls[ifc * 4 + 0];
            int32_t const jcl = m->fccls[ifc * 4 + 1];
            double const * gi = &md->so1c[icl * neq * ndim];
            double * gj = &md->so1n[jcl * neq * ndim];
            if (2 == md->bnd_kind[ib])
            {
                for (int q = 0; q < neq * ndim; ++q) { gj[q] = 0.0; }
                continue;
            }
            double mat[9], inv[9];
            orc_normal_matrix(ndim, &m->fcnml[ifc * ndim], mat);
            transpose(ndim, mat, inv);
            if (0 == md->bnd_kind[ib])
            {
                for (int ieq = 0; ieq < neq; ++ieq)
                {
                    double rot[3], out[3];
                    matvec(ndim, mat, &gi[ieq * ndim], rot);
                    rot[0] = 0.0;
                    matvec(ndim, inv, rot, out);
                    for (int d = 0; d < ndim; ++d) { gj[ieq * ndim + d] = out[d]; }
                }
                continue;
            }
            int const scal[2] = {0, neq - 1};
            for (int q = 0; q < 2; ++q)
            {
                int const ieq = scal[q];
                double rot[3], out[3];
                matvec(ndim, mat, &gi[ieq * ndim], rot);
                rot[0] = -rot[0];
                matvec(ndim, inv, rot, out);
                for (int d = 0; d < ndim; ++d) { gj[ieq * ndim + d] = out[d]; }
            }
            double mg[9], t0[9], uv[9], t1[9], out[9];
            for (int a = 0; a < ndim; ++a)
            {
                for (int b = 0; b < ndim; ++b) { mg[a * ndim + b] = gi[(1 + a) * ndim + b]; }
            }
            matmat(ndim, mat, mg, t0);
            matmat(ndim, t0, inv, uv);
            for (int it = 1; it < ndim; ++it)
            {
                uv[it] = -uv[it];
                uv[it * ndim] = -uv[it * ndim];
            }
            matmat(ndim, inv, uv, t1);
            matmat(ndim, t1, mat, out);
            for (int a = 0; a < ndim; ++a)
            {
                for (int b = 0; b < ndim; ++b) { gj[(1 + a) * ndim + b] = out[a * ndim + b]; }
            }
        }
    }
}

static void md_update(orc_md * md) /* euler.hpp:123-127 */
{
    double * t = md->so0c;
    md->so0c = md->so0n;
    md->so0n = t;
    t = md->so1c;
    md->so1c = md->so1n;
    md->so1n = t;
}

int orc_md_phase(orc_md * md, int phase)
{
    int const ndim = md->mesh->ndim;
    if (ndim != 2 && ndim != 3) { return -1; }
    switch (phase)
    {
    case 0: md_update(md); break;
    case 1: md_solt(md); break;
    case 2: md_soln(md); break;
    case 3: md_bc_soln(md); break;
    case 4: md_cfl(md); break;
    case 5: md_dsoln(md); break;
    case 6: md_bc_dsoln(md); break;
    case 7:
        if (2 == ndim) { md_prepare_ce_2d(md); }
        else { md_prepare_ce_3d(md); }
        break;
    case 8: /* march_substep, euler_march.cpp:339-348 */
        md_update(md);
        md_solt(md);
        md_soln(md);
        md_bc_soln(md);
        md_cfl(md);
        md_dsoln(md);
        md_bc_dsoln(md);
        break;
    default: return -1;
    }
    return 0;
}

int orc_md_march(orc_md * md, int64_t steps) /* euler_march.cpp:350-359 */
{
    for (int64_t i = 0; i < steps; ++i)
    {
        for (int s = 0; s < 2; ++s)
        {
            int const rc = orc_md_phase(md, 8);
            if (rc) { return rc; }
        }
    }
    return 0;
}
