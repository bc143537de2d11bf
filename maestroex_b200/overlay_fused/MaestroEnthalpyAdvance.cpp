// Fused overlay for Source/MaestroEnthalpyAdvance.cpp: same member signature, same forcing /
// perturbational-form bookkeeping (the reference's own members, called in the reference's
// order), but MakeEdgeScal(RhoH) + MakeRhoHFlux (Source/MaestroEnthalpyAdvance.cpp:141-146,
// :229/:312) are ONE library call per box, mx_enthalpy_edge_flux (include/mxadv.h).
//
// Supported: 3-D, plane-parallel and spherical base state (spherical: mx_enthalpy_edge_flux_sph
// with the r0mac / rh0mac / h0mac face arrays the reference builds with MakeS0mac, :166-306),
// enthalpy_pred_type = predict_rhoh, predict_rhohprime (the default) or predict_h — the types
// whose edge state feeds the flux directly.  The temperature-predicting types need
// HfromRhoTedge (EOS on faces) between the two halves: they abort here (no fallback inside this
// file; build with the unfused overlay for them).
#include <MxadvAmrex.H>

using namespace amrex;

void Maestro::EnthalpyAdvance(int which_step, Vector<MultiFab>& scalold,
                              Vector<MultiFab>& scalnew,
                              Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sedge,
                              Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sflux,
                              Vector<MultiFab>& scal_force,
                              Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
                              const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac,
                              const Vector<MultiFab>& thermal) {
    BL_PROFILE_VAR("Maestro::EnthalpyAdvance()", EnthalpyAdvance);

#if (AMREX_SPACEDIM != 3)
    Abort("libmxadv fused EnthalpyAdvance: 3-D only (use the unfused overlay in 2-D)");
#endif
    const bool pert = enthalpy_pred_type == predict_rhohprime;
    const bool spec_h = enthalpy_pred_type == predict_h;
    const bool cons = enthalpy_pred_type == predict_rhoh;
    if (!pert && !spec_h && !cons) {
        Abort("libmxadv fused EnthalpyAdvance: enthalpy_pred_type must be 0, 1 or 2");
    }
    const int nlev = finest_level + 1;
    const int nr1 = base_geom.nr_fine + 1, nrl = base_geom.max_radial_level + 1;

    BaseState<Real> h0_old(nrl, base_geom.nr_fine), h0_new(nrl, base_geom.nr_fine);  // spherical
    BaseState<Real> rho0_edge_old(nrl, nr1), rho0_edge_new(nrl, nr1);
    BaseState<Real> rhoh0_edge_old(nrl, nr1), rhoh0_edge_new(nrl, nr1);
    if (!spherical) {
        CelltoEdge(rho0_old, rho0_edge_old);
        CelltoEdge(rho0_new, rho0_edge_new);
        CelltoEdge(rhoh0_old, rhoh0_edge_old);
        CelltoEdge(rhoh0_new, rhoh0_edge_new);
    }

    if (spec_h) ConvertRhoHToH(scalold, true);

    // forcing at time n (velocity without w0)
    Vector<MultiFab> rhoh0_old_cart(nlev);
    for (int lev = 0; lev < nlev; ++lev) {
        scal_force[lev].setVal(0.);
        rhoh0_old_cart[lev].define(grids[lev], dmap[lev], 1, 1);
        rhoh0_old_cart[lev].setVal(0.);
    }
    Addw0(umac, w0mac, -1.);
    MakeRhoHForce(scal_force, 1, thermal, umac, 1, 1);
    if (pert) {
        Put1dArrayOnCart(rhoh0_old, rhoh0_old_cart, false, false, bcs_s, RhoH);
        ModifyScalForce(scal_force, scalold, umac, rhoh0_edge_old, rhoh0_old_cart, RhoH, bcs_s,
                        0);
    } else if (spec_h) {
        for (int lev = 0; lev < nlev; ++lev) {
            MultiFab::Divide(scal_force[lev], scalold[lev], Rho, RhoH, 1, 1);
        }
    }
    Addw0(umac, w0mac, 1.);

    if (pert) PutInPertForm(scalold, rhoh0_old, RhoH, 0, bcs_f, true);

    // spherical: base state on faces, exactly as the reference forms it before MakeRhoHFlux
    // (Source/MaestroEnthalpyAdvance.cpp:166-306): "old" arrays always, "new" for which_step 2
    Vector<std::array<MultiFab, AMREX_SPACEDIM> > rho0mac_old(nlev), rhoh0mac_old(nlev), h0mac_old(nlev);
    Vector<std::array<MultiFab, AMREX_SPACEDIM> > rho0mac_new(nlev), rhoh0mac_new(nlev), h0mac_new(nlev);
#if (AMREX_SPACEDIM == 3)
    if (spherical) {
        auto define3 = [&](Vector<std::array<MultiFab, AMREX_SPACEDIM> >& m) {
            for (int lev = 0; lev < nlev; ++lev) {
                m[lev][0].define(convert(grids[lev], nodal_flag_x), dmap[lev], 1, 1);
                m[lev][1].define(convert(grids[lev], nodal_flag_y), dmap[lev], 1, 1);
                m[lev][2].define(convert(grids[lev], nodal_flag_z), dmap[lev], 1, 1);
            }
        };
        h0_old.copy(rhoh0_old / rho0_old);
        define3(rho0mac_old);
        define3(rhoh0mac_old);
        define3(h0mac_old);
        MakeS0mac(rho0_old, rho0mac_old);
        MakeS0mac(rhoh0_old, rhoh0mac_old);
        MakeS0mac(h0_old, h0mac_old);
        if (which_step == 2) {
            h0_new.copy(rhoh0_new / rho0_new);
            define3(rho0mac_new);
            define3(rhoh0mac_new);
            define3(h0mac_new);
            MakeS0mac(rho0_new, rho0mac_new);
            MakeS0mac(rhoh0_new, rhoh0mac_new);
            MakeS0mac(h0_new, h0mac_new);
        }
    }
#endif

    // ---- the fused pass: edge state of the predicted enthalpy quantity + (rho h) flux ------
    {
        const BaseState<Real>& r_b = (which_step == 1) ? rho0_old : rho0_new;
        const BaseState<Real>& re_b = (which_step == 1) ? rho0_edge_old : rho0_edge_new;
        const BaseState<Real>& h_b = (which_step == 1) ? rhoh0_old : rhoh0_new;
        const BaseState<Real>& he_b = (which_step == 1) ? rhoh0_edge_old : rhoh0_edge_new;
        const std::vector<int> bc = mxadv::bc_table(bcs_s);
        MxEdgeParams ep;
        ep.dt = dt;
        ep.rel_eps = rel_eps;
        ep.ppm_type = ppm_type;
        ep.slope_order = slope_order;
        ep.ppm_trace_forces = ppm_trace_forces;
        ep.is_vel = 0;
        ep.is_conservative = cons ? 1 : 0;
        ep.start_scomp = RhoH;
        ep.start_bccomp = RhoH;
        ep.num_comp = 1;
        ep.nbccomp = static_cast<int>(bcs_s.size());
        ep.bc = bc.data();
        const MxBase b_ro = mxadv::base_of(rho0_old), b_reo = mxadv::base_of(rho0_edge_old);
        const MxBase b_rn = mxadv::base_of(r_b), b_ren = mxadv::base_of(re_b);
        const MxBase b_ho = mxadv::base_of(rhoh0_old), b_heo = mxadv::base_of(rhoh0_edge_old);
        const MxBase b_hn = mxadv::base_of(h_b), b_hen = mxadv::base_of(he_b);
        for (int lev = 0; lev < nlev; ++lev) {
            MxFluxParams fp;
            fp.species_pred_type = species_pred_type;
            fp.enthalpy_pred_type = enthalpy_pred_type;
            fp.spherical = spherical ? 1 : 0;
            fp.use_exact_base_state = use_exact_base_state ? 1 : 0;
            fp.evolve_base_state = evolve_base_state ? 1 : 0;
            fp.rho_comp = Rho;
            fp.rhoh_comp = RhoH;
            fp.first_spec = FirstSpec;
            fp.num_spec = NumSpec;
            fp.start_comp = RhoH;
            fp.num_comp = 1;
            fp.lev = lev;
            for (MFIter mfi(scalold[lev]); mfi.isValid(); ++mfi) {
                const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
                const MxFab st = mxadv::fab_of(scalold[lev].const_array(mfi));
                const MxFab fo = mxadv::fab_of(scal_force[lev].const_array(mfi));
                const mxadv::Faces se = mxadv::faces_of(sedge[lev], mfi);
                const mxadv::Faces sf = mxadv::faces_of(sflux[lev], mfi);
                const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
                if (spherical) {
                    const bool two = which_step == 2;
                    const mxadv::Faces ro = mxadv::faces_of(rho0mac_old[lev], mfi);
                    const mxadv::Faces rn = mxadv::faces_of(two ? rho0mac_new[lev] : rho0mac_old[lev], mfi);
                    const mxadv::Faces rho_ = mxadv::faces_of(rhoh0mac_old[lev], mfi);
                    const mxadv::Faces rhn = mxadv::faces_of(two ? rhoh0mac_new[lev] : rhoh0mac_old[lev], mfi);
                    const mxadv::Faces ho = mxadv::faces_of(h0mac_old[lev], mfi);
                    const mxadv::Faces hn = mxadv::faces_of(two ? h0mac_new[lev] : h0mac_old[lev], mfi);
                    mxadv::check(mx_enthalpy_edge_flux_sph(&g, &st, se.f, sf.f, um.f, &fo, &ep, ro.f,
                                                           rn.f, rho_.f, rhn.f, ho.f, hn.f, &fp,
                                                           mxadv::stream()),
                                 "EnthalpyAdvance (fused edge + flux, spherical)");
                } else {
                    mxadv::check(mx_enthalpy_edge_flux(&g, &st, se.f, sf.f, um.f, &fo, &ep, &b_ro,
                                                       &b_reo, &b_rn, &b_ren, &b_ho, &b_heo, &b_hn,
                                                       &b_hen, &fp, mxadv::stream()),
                                 "EnthalpyAdvance (fused edge + flux)");
                }
            }
        }
        // AMR bookkeeping of MakeRhoHFlux (Source/MaestroMakeFlux.cpp:1040-1066)
        if (reflux_type == 2) {
            for (int lev = 0; lev < nlev; ++lev) {
                for (int d = 0; d < AMREX_SPACEDIM; ++d) {
                    if (flux_reg_s[lev + 1])
                        flux_reg_s[lev + 1]->CrseInit(sflux[lev][d], d, RhoH, RhoH, 1, -1.0 * dt * mxadv::face_area(geom[lev], d));
                    if (flux_reg_s[lev])
                        flux_reg_s[lev]->FineAdd(sflux[lev][d], d, RhoH, RhoH, 1, 1.0 * dt * mxadv::face_area(geom[lev], d));
                }
            }
        }
        if (reflux_type == 1) AverageDownFaces(sflux);
    }

    // back to (rho h)
    if (pert) PutInPertForm(scalold, rhoh0_old, RhoH, RhoH, bcs_s, false);
    if (spec_h) ConvertRhoHToH(scalold, false);

    // (rho h) force at n+1/2 and the conservative update
    for (int lev = 0; lev < nlev; ++lev) {
        scal_force[lev].setVal(0., RhoH, 1, 1);
    }
    Addw0(umac, w0mac, -1.);
    MakeRhoHForce(scal_force, 0, thermal, umac, 0, which_step);
    Addw0(umac, w0mac, 1.);

    Vector<MultiFab> p0_new_cart(nlev);
    for (int lev = 0; lev < nlev; ++lev) {
        p0_new_cart[lev].define(grids[lev], dmap[lev], 1, 1);
    }
    Put1dArrayOnCart(p0_new, p0_new_cart, false, false, bcs_f, 0);
    UpdateScal(scalold, scalnew, sflux, scal_force, RhoH, 1, p0_new_cart);
}
