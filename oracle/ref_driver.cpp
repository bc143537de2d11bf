// TEST INFRASTRUCTURE ONLY (oracle/): extern "C" driver around the REFERENCE's own
// Maestro::MakeEdgeScal / MakeRhoXFlux / MakeRhoHFlux / CelltoEdge / Makew0, compiled from
// /root/reference/Source/*.cpp against oracle/shim (see oracle/Makefile) into
// oracle/_ref/libmaestro_ref{2,3}d.so.  It aliases the caller's flat arrays as single-box
// MultiFabs and calls the reference member functions unchanged.  Used only by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by
// the product path.
#include <Maestro.H>

#include "../include/mxadv.h"

// The same driver also fronts the PRODUCT's overlay (maestroex_b200/overlay/*.cpp, the
// `Maestro::` members implemented on libmxadv.so) when compiled with -DDRV_OVERLAY: exports
// are then named ovl_* and tests/test_gpu_overlay.py runs reference and overlay side by side.
#ifdef DRV_OVERLAY
#define DRV(name) ovl_##name
#else
#define DRV(name) ref_##name
#endif

using namespace amrex;

// ---- definitions of the globals the shim header declares -------------------------------
namespace maestro {
int ppm_type = 1, slope_order = 4, ppm_trace_forces = 0, bds_type = 0;
int species_pred_type = 1, enthalpy_pred_type = 1, reflux_type = 1;
bool spherical = false, use_exact_base_state = false, evolve_base_state = true,
     average_base_state = false;
int maestro_verbose = 0;
bool do_2d_planar_octant = false, do_planar_invsq_grav = false, add_pb = false;
Real base_cutoff_density = -1.0, dpdt_factor = 0.0, planar_invsq_mass = 0.0,
     grav_const = -1.5e10;
bool do_sponge = false, do_eos_h_above_cutoff = true;
int s0_interp_type = 3, w0_interp_type = 2, s0mac_interp_type = 1, w0mac_interp_type = 1;
Real buoyancy_cutoff_factor = 5.0;
Real anelastic_cutoff_density = -1.0, burning_cutoff_density_lo = -1.0, burning_cutoff_density_hi = 1.e100;
}  // namespace maestro
int Rho = 0, RhoH = 1, Temp = 2, Pi = 3, FirstSpec = 4, NumSpec = 3, Nscal = 7;

#ifndef DRV_OVERLAY
#if (AMREX_SPACEDIM == 2)
IntVect Maestro::nodal_flag(1, 1);
#else
IntVect Maestro::nodal_flag(1, 1, 1);
#endif
#if (AMREX_SPACEDIM == 2)
// Source/MaestroBaseStateGeometry.cpp defines it for 3-D only; MakeCCtoRadii is never called in 2-D
void Maestro::InitBaseStateMapSphr(const int, const MFIter&, const GpuArray<Real, AMREX_SPACEDIM>,
                                   const GpuArray<Real, AMREX_SPACEDIM>) {
    Abort("ref_driver: InitBaseStateMapSphr is 3-D only");
}
#endif
#else
// the overlay test library: one radial level, where the reference's RestrictBase / FillGhostBase
// (loops over radial levels >= 1) do nothing
void Maestro::RestrictBase(BaseState<Real>&, const bool) const {}
void Maestro::RestrictBase(const BaseStateArray<Real>&, const bool) const {}
void Maestro::FillGhostBase(BaseState<Real>&, const bool) const {}
void Maestro::FillGhostBase(const BaseStateArray<Real>&, const bool) const {}
#endif
#if (AMREX_SPACEDIM == 2)
IntVect Maestro::nodal_flag_x(1, 0);
IntVect Maestro::nodal_flag_y(0, 1);
IntVect Maestro::nodal_flag_z(0, 0);
#else
IntVect Maestro::nodal_flag_x(1, 0, 0);
IntVect Maestro::nodal_flag_y(0, 1, 0);
IntVect Maestro::nodal_flag_z(0, 0, 1);
#endif

namespace {

constexpr int DIM = AMREX_SPACEDIM;

IntVect iv(const int* v) {
#if (AMREX_SPACEDIM == 2)
    return IntVect(v[0], v[1]);
#else
    return IntVect(v[0], v[1], v[2]);
#endif
}

Box valid_box(const MxGeom* g) { return Box(iv(g->valid_lo), iv(g->valid_hi)); }

// alias `f` as a MultiFab whose valid box is `g`'s valid box made nodal in `nodal_dir`
// (-1 = cell centred); the ghost width is read off the fab's low corner.
void alias(MultiFab& mf, const MxGeom* g, const MxFab* f, int nodal_dir) {
    Box vb = valid_box(g);
    if (nodal_dir >= 0) vb.surroundingNodes(nodal_dir);
    const int ng = g->valid_lo[0] - f->lo[0];
    for (int d = 0; d < DIM; ++d) {
        if (vb.lo[d] - f->lo[d] != ng || f->hi[d] - vb.hi[d] != ng) {
            Abort("ref_driver: fab box is not the valid box grown uniformly");
        }
    }
    mf.alias(vb, f->ncomp, ng, f->p);
}

void setup(Maestro& m, const MxGeom* g) {
    if (g->dim != DIM) Abort("ref_driver: wrong dimensionality for this build");
    m.finest_level = 0;
    Box dom(iv(g->dom_lo), iv(g->dom_hi));
    m.geom.resize(1);
    m.geom[0] = Geometry(dom, g->dx);
    m.grids.resize(1);
    m.grids[0] = BoxArray(valid_box(g));
    m.dmap.resize(1);
    m.flux_reg_s.resize(2);
}

void fill_base(BaseState<Real>& b, const MxBase* src) {
    if (src == nullptr || src->p == nullptr) {
        b.define(1, 1, 1);
        return;
    }
    b.define(src->nlev, src->len, src->nvar);
    auto a = b.array();
    for (int i = 0; i < src->nlev * src->len * src->nvar; ++i) a(i) = src->p[i];
}

void alias_faces(Vector<std::array<MultiFab, DIM>>& v, const MxGeom* g, const MxFab* f) {
    v.resize(1);
    if (f == nullptr) return;
    for (int d = 0; d < DIM; ++d) alias(v[0][d], g, &f[d], d);
}

void set_flux_globals(const MxFluxParams* p) {
    maestro::species_pred_type = p->species_pred_type;
    maestro::enthalpy_pred_type = p->enthalpy_pred_type;
    maestro::spherical = p->spherical != 0;
    maestro::use_exact_base_state = p->use_exact_base_state != 0;
    maestro::evolve_base_state = p->evolve_base_state != 0;
    maestro::reflux_type = 1;
    Rho = p->rho_comp;
    RhoH = p->rhoh_comp;
    FirstSpec = p->first_spec;
    NumSpec = p->num_spec;
}

}  // namespace

extern "C" {

int DRV(dim)(void) { return DIM; }

int DRV(num_threads)(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void DRV(set_num_threads)(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// Maestro::MakeEdgeScal, Source/MaestroMakeEdgeScalars.cpp:6
int DRV(make_edge_scal)(const MxGeom* g, const MxFab* state, const MxFab* sedge,
                       const MxFab* umac, const MxFab* force, const MxEdgeParams* p) {
    Maestro m;
    setup(m, g);
    m.dt = p->dt;
    m.rel_eps = p->rel_eps;
    maestro::ppm_type = p->ppm_type;
    maestro::slope_order = p->slope_order;
    maestro::ppm_trace_forces = p->ppm_trace_forces;
    maestro::reflux_type = 1;

    Vector<MultiFab> st(1), fo(1);
    alias(st[0], g, state, -1);
    alias(fo[0], g, force, -1);
    Vector<std::array<MultiFab, DIM>> se, um;
    alias_faces(se, g, sedge);
    alias_faces(um, g, umac);

    Vector<BCRec> bcs(p->nbccomp);
    for (int n = 0; n < p->nbccomp; ++n) {
        for (int d = 0; d < DIM; ++d) {
            bcs[n].setLo(d, p->bc[6 * n + d]);
            bcs[n].setHi(d, p->bc[6 * n + 3 + d]);
        }
    }
    m.MakeEdgeScal(st, se, um, fo, p->is_vel != 0, bcs, p->nbccomp, p->start_scomp,
                   p->start_bccomp, p->num_comp, p->is_conservative != 0);
    return 0;
}

// Maestro::MakeRhoXFlux, Source/MaestroMakeFlux.cpp:18
int DRV(make_rhoX_flux)(const MxGeom* g, const MxFab* sflux, const MxFab* etarhoflux,
                       const MxFab* sedge, const MxFab* umac, const MxBase* rho0_old,
                       const MxBase* rho0_edge_old, const MxFab* r0mac_old,
                       const MxBase* rho0_new, const MxBase* rho0_edge_new,
                       const MxFab* r0mac_new, const MxBase* rho0_predicted_edge,
                       const MxBase* w0, const MxFluxParams* p) {
    Maestro m;
    setup(m, g);
    set_flux_globals(p);
    if (p->lev != 0) Abort("ref_driver: single level only");

    // `state` is only used to drive the MFIter: give it the valid box, no data
    Vector<MultiFab> st(1), eta(1);
    st[0].define(m.grids[0], m.dmap[0], 1, 0);
    alias(eta[0], g, etarhoflux, DIM - 1);
    Vector<std::array<MultiFab, DIM>> sf, se, um, r0o, r0n;
    alias_faces(sf, g, sflux);
    alias_faces(se, g, sedge);
    alias_faces(um, g, umac);
    alias_faces(r0o, g, r0mac_old);
    alias_faces(r0n, g, r0mac_new);

    BaseState<Real> b_r0o, b_r0eo, b_r0n, b_r0en, b_r0pe;
    fill_base(b_r0o, rho0_old);
    fill_base(b_r0eo, rho0_edge_old);
    fill_base(b_r0n, rho0_new);
    fill_base(b_r0en, rho0_edge_new);
    fill_base(b_r0pe, rho0_predicted_edge);
    fill_base(m.w0, w0);

    m.MakeRhoXFlux(st, sf, eta, se, um, b_r0o, b_r0eo, r0o, b_r0n, b_r0en, r0n, b_r0pe,
                   p->start_comp, p->num_comp);
    return 0;
}

// Maestro::MakeRhoHFlux, Source/MaestroMakeFlux.cpp:467
int DRV(make_rhoh_flux)(const MxGeom* g, const MxFab* sflux, const MxFab* sedge,
                       const MxFab* umac, const MxBase* rho0_old,
                       const MxBase* rho0_edge_old, const MxFab* r0mac_old,
                       const MxBase* rho0_new, const MxBase* rho0_edge_new,
                       const MxFab* r0mac_new, const MxBase* rhoh0_old,
                       const MxBase* rhoh0_edge_old, const MxFab* rh0mac_old,
                       const MxBase* rhoh0_new, const MxBase* rhoh0_edge_new,
                       const MxFab* rh0mac_new, const MxFab* h0mac_old,
                       const MxFab* h0mac_new, const MxFluxParams* p) {
    Maestro m;
    setup(m, g);
    set_flux_globals(p);
    if (p->lev != 0) Abort("ref_driver: single level only");

    Vector<MultiFab> st(1);
    st[0].define(m.grids[0], m.dmap[0], 1, 0);
    Vector<std::array<MultiFab, DIM>> sf, se, um, w0mac, r0o, r0n, rh0o, rh0n, h0o, h0n;
    alias_faces(sf, g, sflux);
    alias_faces(se, g, sedge);
    alias_faces(um, g, umac);
    w0mac.resize(1);
    alias_faces(r0o, g, r0mac_old);
    alias_faces(r0n, g, r0mac_new);
    alias_faces(rh0o, g, rh0mac_old);
    alias_faces(rh0n, g, rh0mac_new);
    alias_faces(h0o, g, h0mac_old);
    alias_faces(h0n, g, h0mac_new);

    BaseState<Real> b_r0o, b_r0eo, b_r0n, b_r0en, b_rh0o, b_rh0eo, b_rh0n, b_rh0en;
    fill_base(b_r0o, rho0_old);
    fill_base(b_r0eo, rho0_edge_old);
    fill_base(b_r0n, rho0_new);
    fill_base(b_r0en, rho0_edge_new);
    fill_base(b_rh0o, rhoh0_old);
    fill_base(b_rh0eo, rhoh0_edge_old);
    fill_base(b_rh0n, rhoh0_new);
    fill_base(b_rh0en, rhoh0_edge_new);

    m.MakeRhoHFlux(st, sf, se, um, w0mac, b_r0o, b_r0eo, r0o, b_r0n, b_r0en, r0n, b_rh0o,
                   b_rh0eo, rh0o, b_rh0n, b_rh0en, rh0n, h0o, h0n);
    return 0;
}

// Maestro::CelltoEdge, Source/MaestroCelltoEdge.cpp:5 — one level, one chunk [r_lo, r_hi]
int DRV(cell_to_edge)(const double* cell, double* edge, int nr, int r_lo, int r_hi) {
    Maestro m;
    maestro::spherical = false;
    // base_geom tables: BaseStateArray views over local storage
    int nr_tab[1] = {nr};
    int chunks[1] = {1};
    int rs[2] = {0, r_lo};
    int re[2] = {0, r_hi};
    m.base_geom.finest_radial_level = 0;
    m.base_geom.nr.init(nr_tab, 1, 1, 1);
    m.base_geom.numdisjointchunks.init(chunks, 1, 1, 1);
    m.base_geom.r_start_coord.init(rs, 1, 2, 1);
    m.base_geom.r_end_coord.init(re, 1, 2, 1);

    BaseState<Real> c(1, nr, 1), e(1, nr + 1, 1);
    auto ca = c.array();
    auto ea = e.array();
    for (int i = 0; i < nr; ++i) ca(i) = cell[i];
    for (int i = 0; i <= nr; ++i) ea(i) = edge[i];
    m.CelltoEdge(c, e);
    for (int i = 0; i <= nr; ++i) edge[i] = ea(i);
    return 0;
}

namespace {
// one radial level, one chunk [0, nr_fine-1]: the base_geom tables Makew0* reads
struct W0Geom {
    int nr_tab[1], chunks[1], rs[2], re[2], cut[1];
    double dr_tab[1];
    std::vector<double> rcc, redge;
    void attach(Maestro& m, const MxW0Params* p, const double* r_cc, const double* r_edge) {
        const int nr = p->nr_fine;
        nr_tab[0] = nr;
        chunks[0] = 1;
        rs[0] = 0; rs[1] = 0;
        re[0] = 0; re[1] = nr - 1;
        cut[0] = p->base_cutoff_density_coord;
        dr_tab[0] = p->dr;
        rcc.assign(nr, 0.0);
        redge.assign(nr + 1, 0.0);
        for (int r = 0; r < nr; ++r) rcc[r] = r_cc ? r_cc[r] : (r + 0.5) * p->dr;
        for (int r = 0; r <= nr; ++r) redge[r] = r_edge ? r_edge[r] : r * p->dr;
        auto& bg = m.base_geom;
        bg.max_radial_level = 0;
        bg.finest_radial_level = 0;
        bg.nr_fine = nr;
        bg.dr_fine = p->dr;
        bg.nr_irreg = 0;
        bg.nr.init(nr_tab, 1, 1, 1);
        bg.dr.init(dr_tab, 1, 1, 1);
        bg.r_cc_loc.init(rcc.data(), 1, nr, 1);
        bg.r_edge_loc.init(redge.data(), 1, nr + 1, 1);
        bg.base_cutoff_density_coord.init(cut, 1, 1, 1);
        bg.numdisjointchunks.init(chunks, 1, 1, 1);
        bg.r_start_coord.init(rs, 1, 2, 1);
        bg.r_end_coord.init(re, 1, 2, 1);
    }
};

void fill1(BaseState<Real>& b, const double* src, int len) {
    b.define(1, len, 1);
    auto a = b.array();
    for (int i = 0; i < len; ++i) a(i) = src ? src[i] : 0.0;
}

int run_w0(int sphr, const MxW0Params* p, const double* w0_old, const double* Sbar,
           const double* rho0_old, const double* rho0_new, const double* p0_old,
           const double* p0_new, const double* g1o, const double* g1n, const double* pmp,
           const double* etarho_cc, const double* etarho_ec, const double* r_cc,
           const double* r_edge, double* w0, double* w0_force, double* p0bdot) {
    Maestro m;
    W0Geom wg;
    wg.attach(m, p, r_cc, r_edge);
    const int nr = p->nr_fine;
    maestro::spherical = sphr != 0;
    maestro::use_exact_base_state = sphr == 2;  // Makew0SphrIrreg
    maestro::do_planar_invsq_grav = false;
    maestro::do_2d_planar_octant = false;
    maestro::maestro_verbose = 0;
    maestro::add_pb = p->add_pb != 0;
    maestro::grav_const = p->grav_const;
    maestro::dpdt_factor = p->dpdt_factor;
    maestro::base_cutoff_density = p->base_cutoff_density;
    m.dt = p->dt;
    BaseState<Real> b_w0o, b_force, b_S, b_r0o, b_r0n, b_p0o, b_p0n, b_g1o, b_g1n, b_pmp;
    fill1(b_w0o, w0_old, nr + 1);
    fill1(b_force, w0_force, nr);
    fill1(b_S, Sbar, nr);
    fill1(b_r0o, rho0_old, nr);
    fill1(b_r0n, rho0_new, nr);
    fill1(b_p0o, p0_old, nr);
    fill1(b_p0n, p0_new, nr);
    fill1(b_g1o, g1o, nr);
    fill1(b_g1n, g1n, nr);
    fill1(b_pmp, pmp, nr);
    fill1(m.w0, w0, nr + 1);
    fill1(m.etarho_cc, etarho_cc, nr);
    fill1(m.etarho_ec, etarho_ec, nr + 1);
    m.Makew0(b_w0o, b_force, b_S, b_r0o, b_r0n, b_p0o, b_p0n, b_g1o, b_g1n, b_pmp, p->dt_in,
             p->dtold_in, p->is_predictor != 0);
    auto wa = m.w0.array();
    auto fa = b_force.array();
    for (int r = 0; r <= nr; ++r) w0[r] = wa(r);
    for (int r = 0; r < nr; ++r) w0_force[r] = fa(r);
    if (p0bdot && p->add_pb) *p0bdot = m.p0bdot;
    return 0;
}
}  // namespace

// Maestro::Makew0 -> Makew0Planar, Source/MaestroMakew0.cpp:5,60
int DRV(make_w0_planar)(const MxW0Params* p, const double* w0_old, const double* Sbar,
                       const double* p0_old, const double* p0_new, const double* g1o,
                       const double* g1n, const double* pmp, const double* etarho_cc,
                       double* w0, double* w0_force, double* p0bdot) {
    return run_w0(0, p, w0_old, Sbar, nullptr, nullptr, p0_old, p0_new, g1o, g1n, pmp,
                  etarho_cc, nullptr, nullptr, nullptr, w0, w0_force, p0bdot);
}

// Maestro::Makew0 -> Makew0Sphr, Source/MaestroMakew0.cpp:5,537
int DRV(make_w0_sphr)(const MxW0Params* p, const double* w0_old, const double* Sbar,
                     const double* rho0_old, const double* rho0_new, const double* p0_old,
                     const double* p0_new, const double* g1o, const double* g1n,
                     const double* pmp, const double* etarho_cc, const double* etarho_ec,
                     const double* r_cc, const double* r_edge, double* w0, double* w0_force) {
    return run_w0(1, p, w0_old, Sbar, rho0_old, rho0_new, p0_old, p0_new, g1o, g1n, pmp,
                  etarho_cc, etarho_ec, r_cc, r_edge, w0, w0_force, nullptr);
}

// Maestro::Makew0 -> Makew0SphrIrreg (use_exact_base_state), Source/MaestroMakew0.cpp:5,742
int DRV(make_w0_sphr_irreg)(const MxW0Params* p, const double* w0_old, const double* Sbar,
                           const double* rho0_old, const double* rho0_new, const double* p0_old,
                           const double* p0_new, const double* g1o, const double* g1n,
                           const double* pmp, const double* etarho_cc, const double* etarho_ec,
                           const double* r_cc, const double* r_edge, double* w0, double* w0_force) {
    return run_w0(2, p, w0_old, Sbar, rho0_old, rho0_new, p0_old, p0_new, g1o, g1n, pmp,
                  etarho_cc, etarho_ec, r_cc, r_edge, w0, w0_force, nullptr);
}

// Maestro::UpdateScal, Source/MaestroAdvection.cpp:86 (one level, one box).  The reference
// build applies its EOS (here the shim's stand-in, oracle/shim/Maestro.H) to the zones at or
// below base_cutoff_density itself and reports nothing; the overlay build does the same
// through mx_update_scal's eos_mask.
int DRV(update_scal)(const MxGeom* g, const MxFab* sold, const MxFab* snew, const MxFab* sflux,
                    const MxFab* force, const MxFab* p0_cart, const MxUpdateParams* p) {
    Maestro m;
    setup(m, g);
    m.dt = p->dt;
    maestro::base_cutoff_density = p->base_cutoff_density;
    maestro::do_eos_h_above_cutoff = p->do_eos_h_above_cutoff != 0;
    maestro::reflux_type = 1;
    Rho = p->rho_comp;
    RhoH = p->rhoh_comp;
    Temp = p->temp_comp;
    FirstSpec = p->first_spec;
    NumSpec = p->num_spec;
    Nscal = sold->ncomp;
    Vector<MultiFab> so(1), sn(1), fo(1), p0(1);
    alias(so[0], g, sold, -1);
    alias(sn[0], g, snew, -1);
    alias(fo[0], g, force, -1);
    if (p0_cart) alias(p0[0], g, p0_cart, -1);
    Vector<std::array<MultiFab, DIM>> sf;
    alias_faces(sf, g, sflux);
    m.UpdateScal(so, sn, sf, fo, p->start_comp, p->num_comp, p0);
    return 0;
}

#ifndef DRV_OVERLAY
// ==== the passes either side of the path: the reference's own
// MaestroConvert.cpp / MaestroFill3dData.cpp / MaestroForce.cpp / MaestroAdvection.cpp ========
namespace {
// alias a fab whose ghost widths differ per side / direction: valid box = the fab's own box
// shrunk by ng (uniform); used for fabs the driver owns the geometry of
struct SphrGeom {
    int nr_tab[1];
    double dr_tab[1];
    void attach(Maestro& m, const MxGeom* g, const MxSphr* S) {
        maestro::spherical = S && S->spherical;
        maestro::use_exact_base_state = false;
        if (!S) return;
        maestro::s0_interp_type = S->s0_interp_type;
        maestro::w0_interp_type = S->w0_interp_type;
        maestro::s0mac_interp_type = S->s0mac_interp_type;
        maestro::w0mac_interp_type = S->w0mac_interp_type;
        if (!S->spherical) return;
        auto& bg = m.base_geom;
        bg.max_radial_level = 0;
        bg.finest_radial_level = 0;
        bg.nr_fine = S->nr_fine;
        bg.dr_fine = S->dr_fine;
        nr_tab[0] = S->nr_fine;
        dr_tab[0] = S->dr_fine;
        bg.nr.init(nr_tab, 1, 1, 1);
        bg.dr.init(dr_tab, 1, 1, 1);
        bg.r_cc_loc.init(const_cast<double*>(S->r_cc_loc), 1, S->nr_fine, 1);
        bg.r_edge_loc.init(const_cast<double*>(S->r_edge_loc), 1, S->nr_fine + 1, 1);
        for (int d = 0; d < 3; ++d) m.center[d] = S->center[d];
        m.geom[0].setProbLo(S->prob_lo);
        // only read with use_exact_base_state, but its Array4 is formed unconditionally
        m.cell_cc_to_r.resize(1);
        m.cell_cc_to_r[0].define(m.grids[0], m.dmap[0], 1, 0);
    }
};
void set_bcs_f(Maestro& m, int ncomp, int code_hi_last) {
    // bcs_f: "first-order extrap" records; only the entries the tested members read matter
    m.bcs_f.resize(ncomp);
    for (auto& b : m.bcs_f)
        for (int d = 0; d < DIM; ++d) {
            b.setLo(d, BCType::foextrap);
            b.setHi(d, d == DIM - 1 ? code_hi_last : BCType::foextrap);
        }
    m.bcs_u.resize(DIM);
    m.bcs_s.resize(ncomp);
}
// alias with an explicit valid box (fabs whose ghost widths the MxGeom does not describe)
void alias_ng(MultiFab& mf, const MxGeom* g, const MxFab* f, int nodal_dir, int ng) {
    Box vb = valid_box(g);
    if (nodal_dir >= 0) vb.surroundingNodes(nodal_dir);
    for (int d = 0; d < DIM; ++d)
        if (vb.lo[d] - f->lo[d] != ng || f->hi[d] - vb.hi[d] != ng)
            Abort("ref_driver: fab box is not the valid box grown by the expected ghost width");
    mf.alias(vb, f->ncomp, ng, f->p);
}
}  // namespace

// Maestro::ConvertRhoXToX / ConvertRhoHToH, Source/MaestroConvert.cpp:64-114 (the FillPatch /
// AverageDown tail is a no-op in the shim)
int DRV(convert_by_rho)(const MxGeom* g, const MxFab* scal, int rho, int c0, int nc, int flag) {
    Maestro m;
    setup(m, g);
    set_bcs_f(m, scal->ncomp, BCType::foextrap);
    Vector<MultiFab> s(1);
    alias(s[0], g, scal, -1);
    Rho = rho;
    if (nc == 1 && c0 != FirstSpec) {
        RhoH = c0;
        m.ConvertRhoHToH(s, flag != 0);
    } else {
        FirstSpec = c0;
        NumSpec = nc;
        m.ConvertRhoXToX(s, flag != 0);
    }
    return 0;
}

// Maestro::Put1dArrayOnCart (per-level body), Source/MaestroFill3dData.cpp:48
int DRV(put_1d_array_on_cart)(const MxGeom* g, const MxBase* s0, int lev, const MxFab* s0_cart,
                              int edge, int vec, const MxSphr* S) {
    Maestro m;
    setup(m, g);
    SphrGeom sg;
    sg.attach(m, g, S);
    BaseState<Real> b;
    fill_base(b, s0);
    MultiFab mf;
    alias(mf, g, s0_cart, -1);
    if (lev != 0 && !maestro::spherical) Abort("ref_driver: one level only");
    m.Put1dArrayOnCart(0, b, mf, edge != 0, vec != 0);
    return 0;
}

// Maestro::PutInPertForm, Source/MaestroConvert.cpp:6
int DRV(put_in_pert_form)(const MxGeom* g, const MxFab* scal, const MxBase* s0, int lev, int comp,
                          int flag, const MxSphr* S) {
    Maestro m;
    setup(m, g);
    SphrGeom sg;
    sg.attach(m, g, S);
    set_bcs_f(m, scal->ncomp, BCType::foextrap);
    BaseState<Real> b;
    fill_base(b, s0);
    Vector<MultiFab> s(1);
    alias(s[0], g, scal, -1);
    (void)lev;
    m.PutInPertForm(s, b, comp, 0, m.bcs_f, flag != 0);
    return 0;
}

// Maestro::ModifyScalForce, Source/MaestroForce.cpp:261.  The shim's FillPatch is a no-op, so
// the ghost cell of the reference's s0_edge_cart temporary above the top zone is filled here
// the way FillPatch(bcs_f) would (one box covering the domain in the last direction): the
// reference's Put1dArrayOnCart is wrapped for that one call.
int DRV(modify_scal_force)(const MxGeom* g, const MxFab* scal_force, const MxFab* state,
                           const MxFab* umac, const MxBase* s0_edge, int lev, const MxFab* s0_cart,
                           const MxFab* w0_cart, const double* w0, int comp, int fullform,
                           int bcf_hi, const MxFab* /*s0_edge_cart: built inside the reference*/,
                           const MxSphr* S) {
    Maestro m;
    setup(m, g);
    SphrGeom sg;
    sg.attach(m, g, S);
    set_bcs_f(m, state->ncomp, bcf_hi);
    (void)lev;
    Vector<MultiFab> f(1), st(1), s0c(1);
    alias(f[0], g, scal_force, -1);
    alias(st[0], g, state, -1);
    alias(s0c[0], g, s0_cart, -1);
    Vector<std::array<MultiFab, DIM>> um;
    alias_faces(um, g, umac);
    BaseState<Real> be;
    fill_base(be, s0_edge);
    m.w0_cart.resize(1);
    if (w0_cart) alias(m.w0_cart[0], g, w0_cart, -1);
    if (S && S->spherical) fill1(m.w0, w0, S->nr_fine + 1);
    m.fillpatch_hi_code = bcf_hi;
    m.ModifyScalForce(f, st, um, be, s0c, comp, m.bcs_s, fullform);
    return 0;
}

// Maestro::Addw0, Source/MaestroFill3dData.cpp:357
int DRV(addw0)(const MxGeom* g, const MxFab* uedge, const MxFab* w0mac, const MxBase* w0, int lev,
               double mult, int spherical) {
    Maestro m;
    setup(m, g);
    maestro::spherical = spherical != 0;
    (void)lev;
    Vector<std::array<MultiFab, DIM>> ue(1), wm;
    // uedge carries one ghost cell (FillUmacGhost's), w0mac too
    for (int d = 0; d < DIM; ++d) alias(ue[0][d], g, &uedge[d], d);
    alias_faces(wm, g, spherical ? w0mac : nullptr);
    if (w0) fill_base(m.w0, w0);
    m.sold.resize(1);
    m.sold[0].define(m.grids[0], m.dmap[0], 1, 0);
    m.Addw0(ue, wm, mult);
    return 0;
}

#if (AMREX_SPACEDIM == 3)
// Maestro::MakeW0mac, Source/MaestroFill3dData.cpp:464
int DRV(make_w0mac)(const MxGeom* g, const MxFab* w0mac, const MxFab* w0_cart, const double* w0,
                    const MxSphr* S) {
    Maestro m;
    setup(m, g);
    SphrGeom sg;
    sg.attach(m, g, S);
    Vector<std::array<MultiFab, DIM>> wm;
    alias_faces(wm, g, w0mac);
    m.w0_cart.resize(1);
    if (w0_cart)
        alias(m.w0_cart[0], g, w0_cart, -1);
    else
        m.w0_cart[0].define(m.grids[0], m.dmap[0], 3, 2);
    fill1(m.w0, w0, S->nr_fine + 1);
    m.MakeW0mac(wm);
    return 0;
}

// Maestro::MakeS0mac, Source/MaestroFill3dData.cpp:712.  s0mac_interp_type 1 reads the 2-ghost
// s0_cart temporary after FillPatch(bcs_f): the shim's FillPatch hook fills it by first-order
// extrapolation (the reference's bcs_f fill at outflow faces).
int DRV(make_s0mac)(const MxGeom* g, const MxBase* s0, const MxFab* s0mac,
                    const MxFab* /*s0_cart: built inside the reference*/, const MxSphr* S) {
    Maestro m;
    setup(m, g);
    SphrGeom sg;
    sg.attach(m, g, S);
    set_bcs_f(m, 1, BCType::foextrap);
    Vector<std::array<MultiFab, DIM>> sm;
    alias_faces(sm, g, s0mac);
    BaseState<Real> b;
    fill_base(b, s0);
    m.fillpatch_hi_code = BCType::foextrap;
    m.MakeS0mac(b, sm);
    return 0;
}
#endif

// Maestro::UpdateVel, Source/MaestroAdvection.cpp:299
int DRV(update_vel)(const MxGeom* g, const MxFab* uold, const MxFab* unew, const MxFab* umac,
                    const MxFab* uedge, const MxFab* force, const MxFab* sponge, const MxFab* w0_cart,
                    const MxFab* w0mac, double dt, int do_sponge, int spherical) {
    Maestro m;
    setup(m, g);
    maestro::spherical = spherical != 0;
    maestro::do_sponge = do_sponge != 0;
    m.dt = dt;
    m.uold.resize(1);
    m.unew.resize(1);
    alias(m.uold[0], g, uold, -1);
    alias(m.unew[0], g, unew, -1);
    Vector<MultiFab> fo(1), sp(1);
    alias(fo[0], g, force, -1);
    if (sponge)
        alias(sp[0], g, sponge, -1);
    else
        sp[0].define(m.grids[0], m.dmap[0], 1, 0);
    Vector<std::array<MultiFab, DIM>> um, ue, wm;
    alias_faces(um, g, umac);
    alias_faces(ue, g, uedge);
    alias_faces(wm, g, spherical ? w0mac : nullptr);
    m.w0_cart.resize(1);
    if (w0_cart)
        alias(m.w0_cart[0], g, w0_cart, -1);
    else
        m.w0_cart[0].define(m.grids[0], m.dmap[0], DIM, 1);
    m.UpdateVel(um, ue, fo, sp, wm);
    return 0;
}

// Maestro::MakeUtrans (Source/MaestroMakeUtrans.cpp:6) and Maestro::VelPred
// (Source/MaestroVelPred.cpp:5)
namespace {
void set_vel(Maestro& m, const MxVelPredParams* p) {
    m.dt = p->dt;
    m.rel_eps = p->rel_eps;
    maestro::ppm_type = p->ppm_type;
    maestro::slope_order = p->slope_order;
    maestro::ppm_trace_forces = p->ppm_trace_forces;
    maestro::spherical = p->spherical != 0;
    m.bcs_u.resize(DIM);
    for (int n = 0; n < DIM; ++n) {
        for (int d = 0; d < DIM; ++d) {
            m.bcs_u[n].setLo(d, p->bc[6 * n + d]);
            m.bcs_u[n].setHi(d, p->bc[6 * n + 3 + d]);
        }
    }
    for (int d = 0; d < DIM; ++d) {
        m.phys_bc.setLo(d, p->phys_bc_lo[d]);
        m.phys_bc.setHi(d, p->phys_bc_hi[d]);
    }
}
}  // namespace

int DRV(make_utrans)(const MxGeom* g, const MxFab* utilde, const MxFab* ufull, const MxFab* utrans,
                     const MxFab* w0_cart, const MxFab* w0mac, const MxVelPredParams* p) {
    Maestro m;
    setup(m, g);
    set_vel(m, p);
    Vector<MultiFab> ut(1), uf(1);
    alias(ut[0], g, utilde, -1);
    alias(uf[0], g, ufull, -1);
    Vector<std::array<MultiFab, DIM>> tr, wm;
    alias_faces(tr, g, utrans);
    alias_faces(wm, g, p->spherical ? w0mac : nullptr);
    m.w0_cart.resize(1);
    if (w0_cart)
        alias(m.w0_cart[0], g, w0_cart, -1);
    else
        m.w0_cart[0].define(m.grids[0], m.dmap[0], DIM, 1);
    m.MakeUtrans(ut, uf, tr, wm);
    return 0;
}

int DRV(vel_pred)(const MxGeom* g, const MxFab* utilde, const MxFab* ufull, const MxFab* utrans,
                  const MxFab* umac, const MxFab* w0mac, const MxFab* w0_cart, const MxFab* force,
                  const MxVelPredParams* p) {
    Maestro m;
    setup(m, g);
    set_vel(m, p);
    Vector<MultiFab> ut(1), uf(1), fo(1);
    alias(ut[0], g, utilde, -1);
    alias(uf[0], g, ufull, -1);
    alias(fo[0], g, force, -1);
    Vector<std::array<MultiFab, DIM>> tr, um, wm;
    alias_faces(tr, g, utrans);
    alias_faces(um, g, umac);
    alias_faces(wm, g, p->spherical ? w0mac : nullptr);
    m.w0_cart.resize(1);
    if (w0_cart)
        alias(m.w0_cart[0], g, w0_cart, -1);
    else
        m.w0_cart[0].define(m.grids[0], m.dmap[0], DIM, 1);
    m.VelPred(ut, uf, tr, um, wm, fo);
    return 0;
}
#endif  // !DRV_OVERLAY

#ifndef DRV_OVERLAY
// Maestro::Makew0 -> Makew0Planar with a MULTILEVEL base state (Source/MaestroMakew0.cpp:60-272)
// incl. the reference's own RestrictBase / FillGhostBase (MaestroBaseStateGeometry.cpp)
int DRV(make_w0_planar_ml)(const MxW0Params* p, const MxRadialGeom* rg, const double* w0_old,
                          const double* Sbar, const double* p0_old, const double* p0_new,
                          const double* g1o, const double* g1n, const double* pmp,
                          const double* etarho_cc, double* w0, double* w0_force, double* p0bdot) {
    Maestro m;
    const int nlev = rg->max_radial_level + 1, nf = rg->finest_radial_level + 1, nrf = rg->nr_fine;
    auto& bg = m.base_geom;
    bg.max_radial_level = rg->max_radial_level;
    bg.finest_radial_level = rg->finest_radial_level;
    bg.nr_fine = nrf;
    bg.dr_fine = rg->dr[rg->max_radial_level];
    bg.nr_irreg = 0;
    bg.nr.init(const_cast<int*>(rg->nr), nlev, 1, 1);
    bg.dr.init(const_cast<double*>(rg->dr), nlev, 1, 1);
    bg.base_cutoff_density_coord.init(const_cast<int*>(rg->base_cutoff_density_coord), nlev, 1, 1);
    bg.numdisjointchunks.init(const_cast<int*>(rg->numdisjointchunks), nf, 1, 1);
    bg.r_start_coord.init(const_cast<int*>(rg->r_start_coord), nf, rg->chunk_len, 1);
    bg.r_end_coord.init(const_cast<int*>(rg->r_end_coord), nf, rg->chunk_len, 1);
    maestro::spherical = false;
    maestro::use_exact_base_state = false;
    maestro::do_planar_invsq_grav = false;
    maestro::do_2d_planar_octant = false;
    maestro::maestro_verbose = 0;
    maestro::add_pb = p->add_pb != 0;
    maestro::grav_const = p->grav_const;
    maestro::dpdt_factor = p->dpdt_factor;
    maestro::base_cutoff_density = p->base_cutoff_density;
    m.dt = p->dt;
    auto fill = [&](BaseState<Real>& b, const double* src, int len) {
        b.define(nlev, len, 1);
        auto a = b.array();
        for (int i = 0; i < nlev * len; ++i) a(i) = src ? src[i] : 0.0;
    };
    BaseState<Real> b_w0o, b_force, b_S, b_r0, b_p0o, b_p0n, b_g1o, b_g1n, b_pmp;
    fill(b_w0o, w0_old, nrf + 1);
    fill(b_force, nullptr, nrf);
    fill(b_S, Sbar, nrf);
    fill(b_r0, nullptr, nrf);
    fill(b_p0o, p0_old, nrf);
    fill(b_p0n, p0_new, nrf);
    fill(b_g1o, g1o, nrf);
    fill(b_g1n, g1n, nrf);
    fill(b_pmp, pmp, nrf);
    fill(m.w0, nullptr, nrf + 1);
    fill(m.etarho_cc, etarho_cc, nrf);
    fill(m.etarho_ec, nullptr, nrf + 1);
    m.Makew0(b_w0o, b_force, b_S, b_r0, b_r0, b_p0o, b_p0n, b_g1o, b_g1n, b_pmp, p->dt_in, p->dtold_in,
             p->is_predictor != 0);
    auto wa = m.w0.array();
    auto fa = b_force.array();
    for (int i = 0; i < nlev * (nrf + 1); ++i) w0[i] = wa(i);
    for (int i = 0; i < nlev * nrf; ++i) w0_force[i] = fa(i);
    if (p0bdot && p->add_pb) *p0bdot = m.p0bdot;
    return 0;
}
#endif

}  // extern "C"
