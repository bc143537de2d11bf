// Fused overlay for Source/MaestroDensityAdvance.cpp: same member signature
// (Source/Maestro.H, DensityAdvance), same sequence of state conversions and forcing
// (Source/MaestroDensityAdvance.cpp:16-112, 159-177, 236-257 — all of them the reference's own
// members), but the three calls MakeEdgeScal(species) + MakeEdgeScal(rho') + MakeRhoXFlux
// (:115-129, :182/:217) are ONE library call per box, mx_density_edge_flux (include/mxadv.h):
// the edge states never make the extra trip through memory and the MAC velocities, forces and
// base state are read once for all NumSpec + 1 components.  Results are bit-identical to the
// three separate calls (tests/test_gpu_parity.py::test_fused_*, tests/test_gpu_fcompare.py).
//
// Supported: 3-D, species_pred_type = predict_rhoprime_and_X (the default) or
// predict_rho_and_X.  Anything else aborts — there is no fallback inside this file; build
// with the unfused overlay (maestroex_b200/overlay only) for those options.
#include <MxadvAmrex.H>

using namespace amrex;

void Maestro::DensityAdvance(int which_step, Vector<MultiFab>& scalold,
                             Vector<MultiFab>& scalnew,
                             Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sedge,
                             Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sflux,
                             Vector<MultiFab>& scal_force, Vector<MultiFab>& etarhoflux,
                             Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
                             const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac,
                             const BaseState<Real>& rho0_predicted_edge) {
    BL_PROFILE_VAR("Maestro::DensityAdvance()", DensityAdvance);

#if (AMREX_SPACEDIM != 3)
    Abort("libmxadv fused DensityAdvance: 3-D only (use the unfused overlay in 2-D)");
#endif
    const bool perturbational = species_pred_type == predict_rhoprime_and_X;
    if (!perturbational && species_pred_type != predict_rho_and_X) {
        Abort("libmxadv fused DensityAdvance: species_pred_type must be 1 or 3");
    }
    const int nlev = finest_level + 1;
    const int nr1 = base_geom.nr_fine + 1, nrl = base_geom.max_radial_level + 1;

    // interpolated base-state density on edges (planar); spherical uses the MAC arrays below
    BaseState<Real> rho0_edge_old(nrl, nr1), rho0_edge_new(nrl, nr1);
    if (!spherical) {
        CelltoEdge(rho0_old, rho0_edge_old);
        CelltoEdge(rho0_new, rho0_edge_new);
    }

    // forcing at time n: zero for the species, the rho' (or rho) source from ModifyScalForce,
    // evaluated with the full velocity minus w0
    Vector<MultiFab> rho0_old_cart(nlev);
    for (int lev = 0; lev < nlev; ++lev) {
        scal_force[lev].setVal(0.);
        rho0_old_cart[lev].define(grids[lev], dmap[lev], 1, 1);
        rho0_old_cart[lev].setVal(0.);
    }
    Put1dArrayOnCart(rho0_old, rho0_old_cart, false, false, bcs_s, Rho);
    Addw0(umac, w0mac, -1.);
    ModifyScalForce(scal_force, scalold, umac, rho0_edge_old, rho0_old_cart, Rho, bcs_s,
                    perturbational ? 0 : 1);
    Addw0(umac, w0mac, 1.);

    // predicted quantities: X_k and rho' (or rho)
    ConvertRhoXToX(scalold, true);
    if (perturbational) PutInPertForm(scalold, rho0_old, Rho, 0, bcs_f, true);

    // base-state density on faces for the spherical flux forms
    const BaseState<Real>& rho0_b = (which_step == 1) ? rho0_old : rho0_new;
    const BaseState<Real>& rho0_edge_b = (which_step == 1) ? rho0_edge_old : rho0_edge_new;
    Vector<std::array<MultiFab, AMREX_SPACEDIM> > r0mac_a(nlev), r0mac_b(nlev);
    if (spherical) {
        const IntVect nodal[3] = {nodal_flag_x, nodal_flag_y, nodal_flag_z};
        for (int lev = 0; lev < nlev; ++lev) {
            for (int d = 0; d < AMREX_SPACEDIM; ++d) {
                r0mac_a[lev][d].define(convert(grids[lev], nodal[d]), dmap[lev], 1, 1);
                if (which_step == 2)
                    r0mac_b[lev][d].define(convert(grids[lev], nodal[d]), dmap[lev], 1, 1);
            }
        }
        MakeS0mac(rho0_old, r0mac_a);
        if (which_step == 2) MakeS0mac(rho0_new, r0mac_b);
    }
    const auto& r0mac_2nd = (which_step == 1) ? r0mac_a : r0mac_b;

    // ---- the fused pass: edge states of X_k and rho', fluxes, rho flux, etarhoflux ----------
    {
        const std::vector<int> bc = mxadv::bc_table(bcs_s);
        MxEdgeParams ep;
        ep.dt = dt;
        ep.rel_eps = rel_eps;
        ep.ppm_type = ppm_type;
        ep.slope_order = slope_order;
        ep.ppm_trace_forces = ppm_trace_forces;
        ep.is_vel = 0;
        ep.is_conservative = 0;
        ep.start_scomp = FirstSpec;
        ep.start_bccomp = FirstSpec;
        ep.num_comp = NumSpec;
        ep.nbccomp = static_cast<int>(bcs_s.size());
        ep.bc = bc.data();
        const MxBase b_old = mxadv::base_of(rho0_old), b_eold = mxadv::base_of(rho0_edge_old);
        const MxBase b_2nd = mxadv::base_of(rho0_b), b_e2nd = mxadv::base_of(rho0_edge_b);
        const MxBase b_pred = mxadv::base_of(rho0_predicted_edge), b_w0 = mxadv::base_of(w0);
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
            fp.start_comp = FirstSpec;
            fp.num_comp = NumSpec;
            fp.lev = lev;
            for (MFIter mfi(scalold[lev]); mfi.isValid(); ++mfi) {
                const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
                const MxFab st = mxadv::fab_of(scalold[lev].const_array(mfi));
                const MxFab fo = mxadv::fab_of(scal_force[lev].const_array(mfi));
                const MxFab eta = mxadv::fab_of(etarhoflux[lev].const_array(mfi));
                const mxadv::Faces se = mxadv::faces_of(sedge[lev], mfi);
                const mxadv::Faces sf = mxadv::faces_of(sflux[lev], mfi);
                const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
                mxadv::Faces ma, mb;
                if (spherical) {
                    ma = mxadv::faces_of(r0mac_a[lev], mfi);
                    mb = mxadv::faces_of(r0mac_2nd[lev], mfi);
                }
                mxadv::check(mx_density_edge_flux(&g, &st, se.f, sf.f, &eta, um.f, &fo, &ep, Rho,
                                                  &b_old, &b_eold, ma.ptr(), &b_2nd, &b_e2nd,
                                                  mb.ptr(), &b_pred, &b_w0, &fp, mxadv::stream()),
                             "DensityAdvance (fused edge + flux)");
            }
        }
        // AMR bookkeeping of MakeRhoXFlux (Source/MaestroMakeFlux.cpp:420-462)
        if (reflux_type == 2) {
            for (int lev = 0; lev < nlev; ++lev) {
                for (int d = 0; d < AMREX_SPACEDIM; ++d) {
                    if (flux_reg_s[lev + 1]) {
                        flux_reg_s[lev + 1]->CrseInit(sflux[lev][d], d, FirstSpec, FirstSpec,
                                                      NumSpec, -1.0 * dt * mxadv::face_area(geom[lev], d));
                        flux_reg_s[lev + 1]->CrseInit(sflux[lev][d], d, Rho, Rho, 1, -1.0 * dt * mxadv::face_area(geom[lev], d));
                    }
                    if (flux_reg_s[lev]) {
                        flux_reg_s[lev]->FineAdd(sflux[lev][d], d, FirstSpec, FirstSpec, NumSpec,
                                                 1.0 * dt * mxadv::face_area(geom[lev], d));
                        flux_reg_s[lev]->FineAdd(sflux[lev][d], d, Rho, Rho, 1, 1.0 * dt * mxadv::face_area(geom[lev], d));
                    }
                }
            }
        }
        if (reflux_type == 1) AverageDownFaces(sflux);
    }

    // back to conserved form
    if (perturbational) PutInPertForm(scalold, rho0_old, Rho, Rho, bcs_s, false);
    ConvertRhoXToX(scalold, false);

    // conservative update of (rho X)_k, rho = sum
    Vector<MultiFab> p0_new_cart(nlev);
    for (int lev = 0; lev < nlev; ++lev) {
        scal_force[lev].setVal(0.);
        p0_new_cart[lev].define(grids[lev], dmap[lev], 1, 1);
    }
    Put1dArrayOnCart(p0_new, p0_new_cart, false, false, bcs_f, 0);
    UpdateScal(scalold, scalnew, sflux, scal_force, FirstSpec, NumSpec, p0_new_cart);
}
