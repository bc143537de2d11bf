// Overlay for Source/MaestroMakew0.cpp: Maestro::Makew0 (Source/Maestro.H:1185-1196) and the
// members that file defines, on libmxadv.so (mx_make_w0_planar / mx_make_w0_sphr).
// (mx_make_w0_planar / mx_make_w0_planar_ml / mx_make_w0_sphr / mx_make_w0_sphr_irreg).
// Makew0PlanarVarg (do_planar_invsq_grav / do_2d_planar_octant) is not on the accelerated path
// and aborts here (reference convention for illegal states; no CPU fallback).
#include <MxadvAmrex.H>

using namespace amrex;

namespace {
MxW0Params w0_params(const BaseStateGeometry& bg, Real dt_member, Real dt_in, Real dtold_in,
                     bool is_predictor) {
    MxW0Params p;
    p.nr_fine = bg.nr_fine;
    p.base_cutoff_density_coord = bg.base_cutoff_density_coord(0);
    p.is_predictor = is_predictor ? 1 : 0;
    p.add_pb = add_pb ? 1 : 0;
    p.dr = bg.dr(0);
    p.dt = dt_member;
    p.dt_in = dt_in;
    p.dtold_in = dtold_in;
    p.grav_const = grav_const;
    p.dpdt_factor = dpdt_factor;
    p.base_cutoff_density = base_cutoff_density;
    return p;
}
// level 0 of a one-component BaseState is the first `len` entries ((lev*len + r)*nvar)
const Real* row0(const BaseState<Real>& b) { return b.const_array().dataPtr(); }
Real* row0(BaseState<Real>& b) { return b.array().dataPtr(); }
}  // namespace

void Maestro::Makew0(const BaseState<Real>& w0_old, BaseState<Real>& w0_force,
                     const BaseState<Real>& Sbar_in, const BaseState<Real>& rho0_old_in,
                     const BaseState<Real>& rho0_new_in, const BaseState<Real>& p0_old_in,
                     const BaseState<Real>& p0_new_in, const BaseState<Real>& gamma1bar_old_in,
                     const BaseState<Real>& gamma1bar_new_in,
                     const BaseState<Real>& p0_minus_peosbar, const Real dt_in,
                     const Real dtold_in, const bool is_predictor) {
    BL_PROFILE_VAR("Maestro::Makew0()", Makew0);

    w0_force.setVal(0.);  // MaestroMakew0.cpp:18

    if (!spherical) {
        if (do_planar_invsq_grav || do_2d_planar_octant) {
            Makew0PlanarVarg(w0_old, w0_force, Sbar_in, rho0_old_in, rho0_new_in, p0_old_in,
                             p0_new_in, gamma1bar_old_in, gamma1bar_new_in, p0_minus_peosbar,
                             dt_in, dtold_in);
        } else {
            Makew0Planar(w0_old, w0_force, Sbar_in, p0_old_in, p0_new_in, gamma1bar_old_in,
                         gamma1bar_new_in, p0_minus_peosbar, dt_in, dtold_in, is_predictor);
        }
    } else {
        if (use_exact_base_state) {
            Makew0SphrIrreg(w0_old, w0_force, Sbar_in, rho0_old_in, rho0_new_in, p0_old_in,
                            p0_new_in, gamma1bar_old_in, gamma1bar_new_in, p0_minus_peosbar,
                            dt_in, dtold_in);
        } else {
            Makew0Sphr(w0_old, w0_force, Sbar_in, rho0_old_in, rho0_new_in, p0_old_in, p0_new_in,
                       gamma1bar_old_in, gamma1bar_new_in, p0_minus_peosbar, dt_in, dtold_in);
        }
    }

    if (maestro_verbose >= 2) {  // :45-57
        const auto w0_arr = w0.const_array();
        for (auto n = 0; n <= base_geom.finest_radial_level; ++n) {
            Real max_w0 = 0.0;
            for (auto r = base_geom.r_start_coord(n, 1); r <= base_geom.r_end_coord(n, 1) + 1; ++r) {
                max_w0 = amrex::max(max_w0, amrex::Math::abs(w0_arr(n, r)));
            }
            Print() << "... max CFL of w0: " << max_w0 * dt_in / base_geom.dr(n) << std::endl;
        }
        Print() << std::endl;
    }
}

void Maestro::Makew0Planar(const BaseState<Real>& w0_old, BaseState<Real>& w0_force,
                           const BaseState<Real>& Sbar_in, const BaseState<Real>& p0_old_in,
                           const BaseState<Real>& p0_new_in,
                           const BaseState<Real>& gamma1bar_old_in,
                           const BaseState<Real>& gamma1bar_new_in,
                           const BaseState<Real>& p0_minus_peosbar, const Real dt_in,
                           const Real dtold_in, const bool is_predictor) {
    BL_PROFILE_VAR("Maestro::Makew0Planar()", Makew0Planar);
    const MxW0Params p = w0_params(base_geom, dt, dt_in, dtold_in, is_predictor);
    Real pb = 0.0;
    if (base_geom.max_radial_level != 0) {
        // multilevel base state: level loop, restriction / offset, RestrictBase / FillGhostBase
        // all inside the library (mx_make_w0_planar_ml); the BaseState arrays are passed whole
        auto nr = base_geom.nr;
        auto dr = base_geom.dr;
        auto cut = base_geom.base_cutoff_density_coord;
        auto ndc = base_geom.numdisjointchunks;
        auto rs = base_geom.r_start_coord;
        auto re = base_geom.r_end_coord;
        MxRadialGeom rg;
        rg.max_radial_level = base_geom.max_radial_level;
        rg.finest_radial_level = base_geom.finest_radial_level;
        rg.nr_fine = base_geom.nr_fine;
        rg.dr = dr.dataPtr();
        rg.nr = nr.dataPtr();
        rg.base_cutoff_density_coord = cut.dataPtr();
        rg.numdisjointchunks = ndc.dataPtr();
        rg.r_start_coord = rs.dataPtr();
        rg.r_end_coord = re.dataPtr();
        rg.chunk_len = rs.length();
#ifdef AMREX_USE_CUDA
        Gpu::streamSynchronize();  // the 1-D arrays are managed memory read on the host side of the call
#endif
        mxadv::check(mx_make_w0_planar_ml(&p, &rg, row0(w0_old), row0(Sbar_in), row0(p0_old_in),
                                          row0(p0_new_in), row0(gamma1bar_old_in), row0(gamma1bar_new_in),
                                          row0(p0_minus_peosbar), row0(etarho_cc), row0(w0),
                                          row0(w0_force), &pb, mxadv::stream()),
                     "Makew0Planar (multilevel)");
        if (add_pb) {
            p0bdot = pb;
            p0b = p0bdot * dt;
        }
        return;
    }
    mxadv::check(mx_make_w0_planar(&p, row0(w0_old), row0(Sbar_in), row0(p0_old_in),
                                   row0(p0_new_in), row0(gamma1bar_old_in),
                                   row0(gamma1bar_new_in), row0(p0_minus_peosbar),
                                   row0(etarho_cc), row0(w0), row0(w0_force), &pb,
                                   mxadv::stream()),
                 "Makew0Planar");
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();  // w0 / w0_force (managed memory) are read on the host next
#endif
    if (add_pb) {  // :175-180
        p0bdot = pb;
        p0b = p0bdot * dt;
    }
    // RestrictBase / FillGhostBase (:238-239, :270-271) act on radial levels >= 1 only
}

void Maestro::Makew0Sphr(const BaseState<Real>& w0_old, BaseState<Real>& w0_force,
                         const BaseState<Real>& Sbar_in, const BaseState<Real>& rho0_old_in,
                         const BaseState<Real>& rho0_new_in, const BaseState<Real>& p0_old_in,
                         const BaseState<Real>& p0_new_in,
                         const BaseState<Real>& gamma1bar_old_in,
                         const BaseState<Real>& gamma1bar_new_in,
                         const BaseState<Real>& p0_minus_peosbar, const Real dt_in,
                         const Real dtold_in) {
    BL_PROFILE_VAR("Maestro::Makew0Sphr()", Makew0Sphr);
    const MxW0Params p = w0_params(base_geom, dt, dt_in, dtold_in, true);
    // level 0 of base_geom.r_cc_loc / r_edge_loc (BaseStateArray views, nvar == 1)
    mxadv::check(mx_make_w0_sphr(&p, row0(w0_old), row0(Sbar_in), row0(rho0_old_in),
                                 row0(rho0_new_in), row0(p0_old_in), row0(p0_new_in),
                                 row0(gamma1bar_old_in), row0(gamma1bar_new_in),
                                 row0(p0_minus_peosbar), row0(etarho_cc), row0(etarho_ec),
                                 base_geom.r_cc_loc.dataPtr(), base_geom.r_edge_loc.dataPtr(),
                                 row0(w0), row0(w0_force), mxadv::stream()),
                 "Makew0Sphr");
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();  // w0 / w0_force (managed memory) are read on the host next
#endif
}

void Maestro::Makew0PlanarVarg(const BaseState<Real>&, BaseState<Real>&, const BaseState<Real>&,
                               const BaseState<Real>&, const BaseState<Real>&,
                               const BaseState<Real>&, const BaseState<Real>&,
                               const BaseState<Real>&, const BaseState<Real>&,
                               const BaseState<Real>&, const Real, const Real) {
    Abort("libmxadv: Makew0PlanarVarg (do_planar_invsq_grav / do_2d_planar_octant) is not supported");
}

void Maestro::Makew0SphrIrreg(const BaseState<Real>& w0_old, BaseState<Real>& w0_force,
                              const BaseState<Real>& Sbar_in, const BaseState<Real>& rho0_old_in,
                              const BaseState<Real>& rho0_new_in, const BaseState<Real>& p0_old_in,
                              const BaseState<Real>& p0_new_in,
                              const BaseState<Real>& gamma1bar_old_in,
                              const BaseState<Real>& gamma1bar_new_in,
                              const BaseState<Real>& p0_minus_peosbar, const Real dt_in,
                              const Real dtold_in) {
    BL_PROFILE_VAR("Maestro::Makew0SphrIrreg()", Makew0SphrIrreg);
    const MxW0Params p = w0_params(base_geom, dt, dt_in, dtold_in, true);
    mxadv::check(mx_make_w0_sphr_irreg(&p, row0(w0_old), row0(Sbar_in), row0(rho0_old_in),
                                       row0(rho0_new_in), row0(p0_old_in), row0(p0_new_in),
                                       row0(gamma1bar_old_in), row0(gamma1bar_new_in),
                                       row0(p0_minus_peosbar), row0(etarho_cc), row0(etarho_ec),
                                       base_geom.r_cc_loc.dataPtr(), base_geom.r_edge_loc.dataPtr(),
                                       row0(w0), row0(w0_force), mxadv::stream()),
                 "Makew0SphrIrreg");
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();  // w0 / w0_force (managed memory) are read on the host next
#endif
}

// Only the host variants called these (MaestroMakew0.cpp:322-330,:469,:905);
// they stay defined so that the set of symbols this translation unit provides is unchanged.
void Maestro::Tridiag(const BaseStateArray<Real>&, const BaseStateArray<Real>&,
                      const BaseStateArray<Real>&, const BaseStateArray<Real>&,
                      const BaseStateArray<Real>&, const int) {
    Abort("libmxadv: the host Tridiag is not part of the accelerated path (solve is on the device)");
}

void Maestro::ProlongBasetoUniform(const BaseState<Real>&, BaseState<Real>&) const {
    Abort("libmxadv: ProlongBasetoUniform is only used by the unsupported Makew0PlanarVarg");
}
