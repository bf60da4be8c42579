/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * A thin extern "C" veneer over the UNMODIFIED reference classes, so the
 * reference's own CPU march can be driven through ctypes without pybind11 or
 * the reference's CMake build.  It is compiled by oracle/Makefile together
 * with the reference sources *where they lie* under /root/reference/cpp into
 * oracle/_ref/libsolvcon_ref.so (git-ignored).  No reference source is copied
 * into this repository; this file only calls the public C++ API of:
 *
 *   solvcon::onedim::Euler1DCore              cpp/solvcon/onedim/Euler1DCore.hpp:41-145
 *   solvcon::spacetime::{Grid,LinearScalarSolver,InviscidBurgersSolver}
 *                                             cpp/solvcon/spacetime/core.hpp:98-1053
 *   solvcon::StaticMesh                       cpp/solvcon/mesh/StaticMesh.hpp:253-455
 *   solvcon::EulerCore                        cpp/solvcon/multidim/euler.hpp:56-211
 *   solvcon::GradientElement                  cpp/solvcon/multidim/GradientElement.hpp
 *   solvcon::inout::Gmsh                      cpp/solvcon/inout/gmsh.hpp:80-262
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#include <solvcon/solvcon.hpp>
#include <solvcon/onedim/Euler1DCore.hpp>
#include <solvcon/spacetime/core.hpp>
#include <solvcon/spacetime/kernel/linear_scalar.hpp>
#include <solvcon/spacetime/kernel/inviscid_burgers.hpp>
#include <solvcon/mesh/mesh.hpp>
#include <solvcon/multidim/euler.hpp>
#include <solvcon/multidim/GradientElement.hpp>
#include <solvcon/inout/gmsh.hpp>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace sc = solvcon;

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH                                    \
    }                                                \
    catch (std::exception const & e)                 \
    {                                                \
        g_err = e.what();                            \
        return -1;                                   \
    }                                                \
    return 0;

template <class S>
static void st_phase(S & s, int phase, bool odd, int alpha)
{
    switch (phase)
    {
    case 0: s.march_half_so0(odd); break;
    case 1: s.update_cfl(odd); break;
    case 2:
        if (0 == alpha) { s.template march_half_so1_alpha<0>(odd); }
        else if (1 == alpha) { s.template march_half_so1_alpha<1>(odd); }
        else if (2 == alpha) { s.template march_half_so1_alpha<2>(odd); }
        else { throw std::invalid_argument("ref_st_phase: alpha must be 0, 1 or 2"); }
        break;
    case 3: s.treat_boundary_so0(); break;
    case 4: s.treat_boundary_so1(); break;
    default: throw std::invalid_argument("ref_st_phase: bad phase");
    }
}

template <class S>
static void st_march(S & s, int alpha, size_t steps)
{
    if (0 == alpha) { s.template march_alpha<0>(steps); }
    else if (1 == alpha) { s.template march_alpha<1>(steps); }
    else if (2 == alpha) { s.template march_alpha<2>(steps); }
    else { throw std::invalid_argument("ref_st_march: alpha must be 0, 1 or 2"); }
}

extern "C" {

const char * ref_last_error() { return g_err.c_str(); }

/* ---------------------------------------------------------------- 1D Euler */

struct ref_e1d
{
    std::shared_ptr<sc::onedim::Euler1DCore> p;
};

ref_e1d * ref_e1d_create(int64_t ncoord, double dt)
{
    try
    {
        auto * h = new ref_e1d;
        h->p = sc::onedim::Euler1DCore::construct(static_cast<size_t>(ncoord), dt);
        return h;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return nullptr;
    }
}
void ref_e1d_destroy(ref_e1d * h) { delete h; }
double * ref_e1d_coord(ref_e1d * h) { return h->p->coord().data(); }
double * ref_e1d_cfl(ref_e1d * h) { return h->p->cfl().data(); }
double * ref_e1d_so0(ref_e1d * h) { return h->p->so0().data(); }
double * ref_e1d_so1(ref_e1d * h) { return h->p->so1().data(); }
double * ref_e1d_gamma(ref_e1d * h) { return h->p->gamma().data(); }

/* phase: 0 so0, 1 cfl, 2 so1, 3 bc_so0, 4 bc_so1 */
int ref_e1d_phase(ref_e1d * h, int phase, int odd, int alpha)
{
    REF_TRY
    auto & s = *h->p;
    switch (phase)
    {
    case 0: s.march_half_so0(odd != 0); break;
    case 1: s.update_cfl(odd != 0); break;
    case 2:
        if (1 == alpha) { s.march_half_so1_alpha<1>(odd != 0); }
        else if (2 == alpha) { s.march_half_so1_alpha<2>(odd != 0); }
        else { throw std::invalid_argument("ref_e1d_phase: alpha must be 1 or 2"); }
        break;
    case 3: s.treat_boundary_so0(); break;
    case 4: s.treat_boundary_so1(); break;
    default: throw std::invalid_argument("ref_e1d_phase: bad phase");
    }
    REF_CATCH
}

int ref_e1d_march(ref_e1d * h, int alpha, int64_t steps)
{
    REF_TRY
    if (1 == alpha) { h->p->march_alpha<1>(static_cast<size_t>(steps)); }
    else if (2 == alpha) { h->p->march_alpha<2>(static_cast<size_t>(steps)); }
    else { throw std::invalid_argument("ref_e1d_march: alpha must be 1 or 2"); }
    REF_CATCH
}

/* derived fields of Euler1DCore.hpp:106-117; which: 0 density, 1 velocity, 2 pressure, 3 temperature,
 * 4 internal_energy, 5 entropy */
int ref_e1d_derived(ref_e1d * h, int which, double * out)
{
    REF_TRY
    auto const & s = *h->p;
    sc::SimpleArray<double> r;
    switch (which)
    {
    case 0: r = s.density(); break;
    case 1: r = s.velocity(); break;
    case 2: r = s.pressure(); break;
    case 3: r = s.temperature(); break;
    case 4: r = s.internal_energy(); break;
    case 5: r = s.entropy(); break;
    default: throw std::invalid_argument("ref_e1d_derived: bad field");
    }
    for (size_t it = 0; it < r.size(); ++it) { out[it] = r.data()[it]; }
    REF_CATCH
}

/* --------------------------------------------------------------- 1D scalar */

struct ref_st
{
    int eq = 0; /* 0 linear scalar, 1 inviscid Burgers */
    std::shared_ptr<sc::spacetime::Grid> grid;
    std::shared_ptr<sc::spacetime::LinearScalarSolver> lin;
    std::shared_ptr<sc::spacetime::InviscidBurgersSolver> bur;
};

ref_st * ref_st_create(int eq, double xmin, double xmax, int64_t ncelm, double dt)
{
    try
    {
        auto * h = new ref_st;
        h->eq = eq;
        h->grid = sc::spacetime::Grid::construct(xmin, xmax, static_cast<size_t>(ncelm));
        if (0 == eq) { h->lin = sc::spacetime::LinearScalarSolver::construct(h->grid, dt); }
        else { h->bur = sc::spacetime::InviscidBurgersSolver::construct(h->grid, dt); }
        return h;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return nullptr;
    }
}

/* Grid from explicit CE boundary locations (Grid(xloc), core.hpp:129). */
ref_st * ref_st_create_xloc(int eq, double const * xloc, int64_t nloc, double dt)
{
    try
    {
        auto * h = new ref_st;
        h->eq = eq;
        sc::SimpleArray<double> arr(sc::small_vector<ssize_t>{static_cast<ssize_t>(nloc)});
        for (int64_t i = 0; i < nloc; ++i) { arr[i] = xloc[i]; }
        h->grid = sc::spacetime::Grid::construct(arr);
        if (0 == eq) { h->lin = sc::spacetime::LinearScalarSolver::construct(h->grid, dt); }
        else { h->bur = sc::spacetime::InviscidBurgersSolver::construct(h->grid, dt); }
        return h;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return nullptr;
    }
}
void ref_st_destroy(ref_st * h) { delete h; }
int64_t ref_st_nx(ref_st * h) { return static_cast<int64_t>(h->grid->xsize()); }
double * ref_st_x(ref_st * h) { return h->grid->xcoord().data(); }
double * ref_st_so0(ref_st * h) { return h->eq ? h->bur->so0().data() : h->lin->so0().data(); }
double * ref_st_so1(ref_st * h) { return h->eq ? h->bur->so1().data() : h->lin->so1().data(); }
double * ref_st_cfl(ref_st * h) { return h->eq ? h->bur->cfl().data() : h->lin->cfl().data(); }

int ref_st_phase(ref_st * h, int phase, int odd, int alpha)
{
    REF_TRY
    if (h->eq) { st_phase(*h->bur, phase, odd != 0, alpha); }
    else { st_phase(*h->lin, phase, odd != 0, alpha); }
    REF_CATCH
}

int ref_st_march(ref_st * h, int alpha, int64_t steps)
{
    REF_TRY
    if (h->eq) { st_march(*h->bur, alpha, static_cast<size_t>(steps)); }
    else { st_march(*h->lin, alpha, static_cast<size_t>(steps)); }
    REF_CATCH
}

/* -------------------------------------------------------------------- mesh */

struct ref_mesh
{
    std::shared_ptr<sc::StaticMesh> p;
};

ref_mesh * ref_mesh_create(int ndim, int64_t nnode, int64_t nface, int64_t ncell)
{
    try
    {
        auto * h = new ref_mesh;
        h->p = sc::StaticMesh::construct(static_cast<uint8_t>(ndim),
                                         static_cast<uint32_t>(nnode),
                                         static_cast<uint32_t>(nface),
                                         static_cast<uint32_t>(ncell));
        return h;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return nullptr;
    }
}
// Gmsh(data).to_block(), cpp/solvcon/inout/gmsh.cpp:10-117 (the reference's reader, unmodified)
ref_mesh * ref_gmsh_to_block(char const * data, int64_t nbytes)
{
    try
    {
        sc::inout::Gmsh gmsh(std::string(data, static_cast<size_t>(nbytes)));
        auto * h = new ref_mesh;
        h->p = gmsh.to_block();
        return h;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return nullptr;
    }
}

void ref_mesh_destroy(ref_mesh * h) { delete h; }

int ref_mesh_build_interior(ref_mesh * h, int do_metric, int do_edge)
{
    REF_TRY
    h->p->build_interior(do_metric != 0, do_edge != 0);
    REF_CATCH
}
int ref_mesh_add_bc(ref_mesh * h, char const * name, int32_t const * faces, int64_t n)
{
    REF_TRY
    std::vector<int32_t> v(faces, faces + n);
    h->p->add_bc(std::string(name), v);
    REF_CATCH
}
int ref_mesh_build_boundary(ref_mesh * h)
{
    REF_TRY
    h->p->build_boundary();
    REF_CATCH
}
int ref_mesh_build_ghost(ref_mesh * h)
{
    REF_TRY
    h->p->build_ghost();
    REF_CATCH
}

/* counts: 0 ndim 1 nnode 2 nface 3 ncell 4 nbound 5 ngstnode 6 ngstface 7 ngstcell 8 nedge */
int64_t ref_mesh_count(ref_mesh * h, int which)
{
    auto & m = *h->p;
    switch (which)
    {
    case 0: return m.ndim();
    case 1: return m.nnode();
    case 2: return m.nface();
    case 3: return m.ncell();
    case 4: return m.nbound();
    case 5: return m.ngstnode();
    case 6: return m.ngstface();
    case 7: return m.ngstcell();
    case 8: return m.nedge();
    default: return -1;
    }
}

/*
 * Table access by name.  Returns the pointer to the first stored element
 * (ghost rows included), writes rows (incl. ghost), cols and nghost.
 * is_real: 1 for double tables, 0 for int32 tables.
 */
void * ref_mesh_table(ref_mesh * h, char const * name, int64_t * rows, int64_t * cols, int64_t * nghost, int * is_real)
{
    auto & m = *h->p;
    std::string const n(name);
#define REF_TBL(NAME, REAL)                                                        \
    if (n == #NAME)                                                                \
    {                                                                              \
        auto & a = m.NAME();                                                       \
        *rows = a.shape().empty() ? 0 : a.shape()[0];                              \
        *cols = a.shape().size() > 1 ? a.shape()[1] : 1;                           \
        *nghost = a.nghost();                                                      \
        *is_real = REAL;                                                           \
        return static_cast<void *>(a.data());                                      \
    }
    REF_TBL(ndcrd, 1)
    REF_TBL(fccnd, 1)
    REF_TBL(fcnml, 1)
    REF_TBL(fcara, 1)
    REF_TBL(clcnd, 1)
    REF_TBL(clvol, 1)
    REF_TBL(fctpn, 0)
    REF_TBL(cltpn, 0)
    REF_TBL(clgrp, 0)
    REF_TBL(fcnds, 0)
    REF_TBL(fccls, 0)
    REF_TBL(clnds, 0)
    REF_TBL(clfcs, 0)
    REF_TBL(ednds, 0)
    REF_TBL(bndfcs, 0)
#undef REF_TBL
    g_err = "ref_mesh_table: unknown table " + n;
    return nullptr;
}

/* ---------------------------------------------------------------- multidim */

struct ref_md
{
    std::shared_ptr<sc::EulerCore> p;
};

ref_md * ref_md_create(ref_mesh * mesh, double dt)
{
    try
    {
        auto * h = new ref_md;
        h->p = sc::EulerCore::construct(mesh->p, dt);
        return h;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return nullptr;
    }
}
void ref_md_destroy(ref_md * h) { delete h; }

int ref_md_set_params(ref_md * h, double sigma0, double taumin, double tauscale)
{
    h->p->set_sigma0(sigma0);
    h->p->set_taumin(taumin);
    h->p->set_tauscale(tauscale);
    return 0;
}

int ref_md_init_solution(ref_md * h, double gamma, double rho, double const * v, double p)
{
    REF_TRY
    std::array<double, 3> vel = {v[0], v[1], v[2]};
    h->p->init_solution(gamma, rho, vel, p);
    REF_CATCH
}

/* kind: 0 nonrefl, 1 slipwall, 2 inlet (EulerBC, euler.hpp:27-32) */
int ref_md_add_bc(ref_md * h, int kind, int32_t const * faces, int64_t n, double const * value, int nvalue)
{
    REF_TRY
    std::vector<int32_t> f(faces, faces + n);
    std::vector<double> v(value, value + nvalue);
    h->p->add_bc(static_cast<sc::EulerBC>(kind), f, v);
    REF_CATCH
}
int ref_md_clear_bc(ref_md * h)
{
    h->p->clear_bc();
    return 0;
}

/* phase: 0 update 1 solt 2 soln 3 bc_soln 4 cfl 5 dsoln 6 bc_dsoln 7 prepare_ce 8 march_substep */
int ref_md_phase(ref_md * h, int phase)
{
    REF_TRY
    auto & s = *h->p;
    switch (phase)
    {
    case 0: s.update(); break;
    case 1: s.calc_solt(); break;
    case 2: s.calc_soln(); break;
    case 3: s.bc_soln(); break;
    case 4: s.calc_cfl(); break;
    case 5: s.calc_dsoln(); break;
    case 6: s.bc_dsoln(); break;
    case 7: s.prepare_ce(); break;
    case 8: s.march_substep(); break;
    default: throw std::invalid_argument("ref_md_phase: bad phase");
    }
    REF_CATCH
}

int ref_md_march(ref_md * h, int64_t steps)
{
    REF_TRY
    h->p->march(static_cast<int32_t>(steps));
    REF_CATCH
}

/* Array access by name; pointer to the first stored element (ghost rows first). */
// GradientElement(mesh, cecnd, icl, tau) of the reference (GradientElement.cpp:137-213) flattened for ctypes:
// rcl[6], idis[6][3], jdis[6][3], faces[8][3] (1-based), dmat[8][3][3], and grads[8][3] = solve_gradient(ifge, udf)
int ref_md_gradient_element(ref_md * h, ref_mesh * mesh, int32_t icl, double tau, double const * udf, int32_t * meta,
                            int32_t * rcl, double * idis, double * jdis, int32_t * faces, double * dmat, double * grads)
{
    try
    {
        sc::GradientElement ge(*mesh->p, h->p->cecnd(), icl, tau);
        int const nd = ge.ndim();
        meta[0] = ge.icl();
        meta[1] = nd;
        meta[2] = ge.clnfc();
        meta[3] = ge.nfge();
        idis[18] = ge.nfge_inverse(); // one spare slot behind idis[6][3]
        for (int f = 0; f < ge.clnfc(); ++f)
        {
            rcl[f] = ge.rcl(f);
            for (int d = 0; d < nd; ++d)
            {
                idis[f * 3 + d] = ge.idis(f, d);
                jdis[f * 3 + d] = ge.jdis(f, d);
            }
        }
        for (int g = 0; g < ge.nfge(); ++g)
        {
            auto const & tf = ge.faces(g);
            auto const dm = ge.displacement_matrix(g);
            sc::GradientElement::ge_vector_type u = {0, 0, 0};
            for (int d = 0; d < nd; ++d) { u[d] = udf[d]; }
            auto const gr = ge.solve_gradient(g, u);
            for (int v = 0; v < nd; ++v)
            {
                faces[g * 3 + v] = tf[v];
                grads[g * 3 + v] = gr[v];
                for (int d = 0; d < nd; ++d) { dmat[(g * 3 + v) * 3 + d] = dm[v][d]; }
            }
        }
        return 0;
    }
    catch (std::exception const & e)
    {
        g_err = e.what();
        return 1;
    }
}

double * ref_md_array(ref_md * h, char const * name, int64_t * size, int64_t * nghost)
{
    auto & s = *h->p;
    std::string const n(name);
#define REF_ARR(NAME)                       \
    if (n == #NAME)                         \
    {                                       \
        auto & a = s.NAME();                \
        *size = static_cast<int64_t>(a.size()); \
        *nghost = a.nghost();               \
        return a.data();                    \
    }
    REF_ARR(cevol)
    REF_ARR(cecnd)
    REF_ARR(sfcnd)
    REF_ARR(sfnml)
    REF_ARR(so0c)
    REF_ARR(so0n)
    REF_ARR(so0t)
    REF_ARR(so1c)
    REF_ARR(so1n)
    REF_ARR(stm)
    REF_ARR(cflo)
    REF_ARR(cflc)
    REF_ARR(gamma)
#undef REF_ARR
    g_err = "ref_md_array: unknown array " + n;
    return nullptr;
}

int ref_md_normal_matrix(ref_md * h, int32_t ifc, double * out /* ndim*ndim */)
{
    REF_TRY
    auto const m = h->p->get_normal_matrix(ifc);
    size_t k = 0;
    for (auto const & row : m)
    {
        for (double v : row) { out[k++] = v; }
    }
    REF_CATCH
}

} /* extern "C" */
