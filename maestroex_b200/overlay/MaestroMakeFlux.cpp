// Overlay for Source/MaestroMakeFlux.cpp: Maestro::MakeRhoXFlux / MakeRhoHFlux with the
// reference signatures (Source/Maestro.H:230-245, 270-295); the face loops run in
// libmxadv.so (mx_make_rhoX_flux / mx_make_rhoh_flux).  Flux registers and
// AverageDownFaces (AMR bookkeeping, Source/MaestroMakeFlux.cpp:420-462, 1028-1066) remain
// the reference's own code and are invoked exactly as there.
#include <MxadvAmrex.H>

using namespace amrex;

namespace {

MxFluxParams flux_params(int lev, int start_comp, int num_comp) {
    MxFluxParams p;
    p.species_pred_type = species_pred_type;
    p.enthalpy_pred_type = enthalpy_pred_type;
    p.spherical = spherical ? 1 : 0;
    p.use_exact_base_state = use_exact_base_state ? 1 : 0;
    p.evolve_base_state = evolve_base_state ? 1 : 0;
    p.rho_comp = Rho;
    p.rhoh_comp = RhoH;
    p.first_spec = FirstSpec;
    p.num_spec = NumSpec;
    p.start_comp = start_comp;
    p.num_comp = num_comp;
    p.lev = lev;
    return p;
}

}  // namespace

void Maestro::MakeRhoXFlux(
    const Vector<MultiFab>& state,
    Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sflux,
    Vector<MultiFab>& etarhoflux,
    Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sedge,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
    const BaseState<Real>& rho0_old_in, const BaseState<Real>& rho0_edge_old_state,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& r0mac_old,
    const BaseState<Real>& rho0_new_in, const BaseState<Real>& rho0_edge_new_state,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& r0mac_new,
    const BaseState<Real>& rho0_predicted_edge_state, int start_comp, int num_comp) {
    BL_PROFILE_VAR("Maestro::MakeRhoXFlux()", MakeRhoXFlux);

    const MxBase r0o = mxadv::base_of(rho0_old_in);
    const MxBase r0eo = mxadv::base_of(rho0_edge_old_state);
    const MxBase r0n = mxadv::base_of(rho0_new_in);
    const MxBase r0en = mxadv::base_of(rho0_edge_new_state);
    const MxBase r0pe = mxadv::base_of(rho0_predicted_edge_state);
    const MxBase w0b = mxadv::base_of(w0);

    for (int lev = 0; lev <= finest_level; ++lev) {
        const MxFluxParams par = flux_params(lev, start_comp, num_comp);
        for (MFIter mfi(state[lev]); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const mxadv::Faces sf = mxadv::faces_of(sflux[lev], mfi);
            const mxadv::Faces se = mxadv::faces_of(sedge[lev], mfi);
            const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
            const MxFab eta = mxadv::fab_of(etarhoflux[lev].const_array(mfi));
            mxadv::Faces m_old, m_new;
            if (spherical) {
                m_old = mxadv::faces_of(r0mac_old[lev], mfi);
                m_new = mxadv::faces_of(r0mac_new[lev], mfi);
            }
            mxadv::check(mx_make_rhoX_flux(&g, sf.f, &eta, se.f, um.f, &r0o, &r0eo,
                                           m_old.ptr(), &r0n, &r0en, m_new.ptr(), &r0pe, &w0b,
                                           &par, mxadv::stream()),
                         "MakeRhoXFlux");
        }
    }

    // AMR bookkeeping, as Source/MaestroMakeFlux.cpp:420-462
    if (reflux_type == 2) {
        for (int lev = 0; lev <= finest_level; ++lev) {
            if (flux_reg_s[lev + 1]) {
                for (int i = 0; i < AMREX_SPACEDIM; ++i) {
                    flux_reg_s[lev + 1]->CrseInit(sflux[lev][i], i, start_comp, start_comp,
                                                  num_comp, -1.0 * dt * mxadv::face_area(geom[lev], i));
                    flux_reg_s[lev + 1]->CrseInit(sflux[lev][i], i, Rho, Rho, 1,
                                                  -1.0 * dt * mxadv::face_area(geom[lev], i));
                }
            }
            if (flux_reg_s[lev]) {
                for (int i = 0; i < AMREX_SPACEDIM; ++i) {
                    flux_reg_s[lev]->FineAdd(sflux[lev][i], i, start_comp, start_comp,
                                             num_comp, 1.0 * dt * mxadv::face_area(geom[lev], i));
                    flux_reg_s[lev]->FineAdd(sflux[lev][i], i, Rho, Rho, 1,
                                             1.0 * dt * mxadv::face_area(geom[lev], i));
                }
            }
        }
    }
    if (reflux_type == 1) {
        AverageDownFaces(sflux);
    }
}

void Maestro::MakeRhoHFlux(
    const Vector<MultiFab>& state,
    Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sflux,
    Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sedge,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac,
    const BaseState<Real>& rho0_old_in, const BaseState<Real>& rho0_edge_old,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& r0mac_old,
    const BaseState<Real>& rho0_new_in, const BaseState<Real>& rho0_edge_new,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& r0mac_new,
    const BaseState<Real>& rhoh0_old_in, const BaseState<Real>& rhoh0_edge_old,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& rh0mac_old,
    const BaseState<Real>& rhoh0_new_in, const BaseState<Real>& rhoh0_edge_new,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& rh0mac_new,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& h0mac_old,
    const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& h0mac_new) {
    BL_PROFILE_VAR("Maestro::MakeRhoHFlux()", MakeRhoHFlux);
    amrex::ignore_unused(w0mac);  // read by no flux form (Source/MaestroMakeFlux.cpp:467-1067)

    const MxBase r0o = mxadv::base_of(rho0_old_in), r0eo = mxadv::base_of(rho0_edge_old);
    const MxBase r0n = mxadv::base_of(rho0_new_in), r0en = mxadv::base_of(rho0_edge_new);
    const MxBase h0o = mxadv::base_of(rhoh0_old_in), h0eo = mxadv::base_of(rhoh0_edge_old);
    const MxBase h0n = mxadv::base_of(rhoh0_new_in), h0en = mxadv::base_of(rhoh0_edge_new);

    for (int lev = 0; lev <= finest_level; ++lev) {
        const MxFluxParams par = flux_params(lev, RhoH, 1);
        for (MFIter mfi(state[lev]); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const mxadv::Faces sf = mxadv::faces_of(sflux[lev], mfi);
            const mxadv::Faces se = mxadv::faces_of(sedge[lev], mfi);
            const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
            mxadv::Faces ro, rn, rho, rhn, ho, hn;
            if (spherical) {
                ro = mxadv::faces_of(r0mac_old[lev], mfi);
                rn = mxadv::faces_of(r0mac_new[lev], mfi);
                rho = mxadv::faces_of(rh0mac_old[lev], mfi);
                rhn = mxadv::faces_of(rh0mac_new[lev], mfi);
                ho = mxadv::faces_of(h0mac_old[lev], mfi);
                hn = mxadv::faces_of(h0mac_new[lev], mfi);
            }
            mxadv::check(
                mx_make_rhoh_flux(&g, sf.f, se.f, um.f, &r0o, &r0eo, ro.ptr(), &r0n, &r0en,
                                  rn.ptr(), &h0o, &h0eo, rho.ptr(), &h0n, &h0en, rhn.ptr(),
                                  ho.ptr(), hn.ptr(), &par, mxadv::stream()),
                "MakeRhoHFlux");
        }
    }

    // AMR bookkeeping, as Source/MaestroMakeFlux.cpp:1028-1066
    if (reflux_type == 2) {
        for (int lev = 0; lev <= finest_level; ++lev) {
            if (flux_reg_s[lev + 1]) {
                for (int i = 0; i < AMREX_SPACEDIM; ++i) {
                    flux_reg_s[lev + 1]->CrseInit(sflux[lev][i], i, RhoH, RhoH, 1,
                                                  -1.0 * dt * mxadv::face_area(geom[lev], i));
                }
            }
            if (flux_reg_s[lev]) {
                for (int i = 0; i < AMREX_SPACEDIM; ++i) {
                    flux_reg_s[lev]->FineAdd(sflux[lev][i], i, RhoH, RhoH, 1,
                                             1.0 * dt * mxadv::face_area(geom[lev], i));
                }
            }
        }
    }
    if (reflux_type == 1) {
        AverageDownFaces(sflux);
    }
}
