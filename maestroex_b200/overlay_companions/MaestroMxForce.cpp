// Overlay for Maestro::ModifyScalForce (Source/MaestroForce.cpp:261-467) on libmxadv.so
// (mx_modify_scal_force).  The reference's 1-ghost s0_edge_cart temporary is still built the
// reference's way (Put1dArrayOnCart + FillPatch, so that coarse-fine boundaries get AMReX's own
// interpolation) and handed to the kernel as a fab; div(w0) of the spherical branch is evaluated
// inside the library from the 1-D w0.  AverageDown / FillPatch tail unchanged.
// Build: compile the reference's MaestroForce.cpp with -DModifyScalForce=ModifyScalForce_reference
// and add this file to CEXE_sources.
#include <MxadvSphr.H>

using namespace amrex;

void Maestro::ModifyScalForce(Vector<MultiFab>& scal_force, const Vector<MultiFab>& state,
                              const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac_in,
                              const BaseState<Real>& s0_edge, const Vector<MultiFab>& s0_cart, int comp,
                              [[maybe_unused]] const Vector<BCRec>& bcs, int fullform) {
    BL_PROFILE_VAR("Maestro::ModifyScalForce()", ModifyScalForce);
    mxadv::no_exact_base_state("ModifyScalForce");

    Vector<MultiFab> s0_edge_cart(finest_level + 1);
    for (int lev = 0; lev <= finest_level; ++lev) {
        s0_edge_cart[lev].define(grids[lev], dmap[lev], 1, 1);
    }
    if (!spherical) {
        Put1dArrayOnCart(s0_edge, s0_edge_cart, false, false, bcs_f, 0);  // :275-277
    }
    const MxBase be = mxadv::base_of(s0_edge);
    const Real* w0p = w0.const_array().dataPtr();
    for (int lev = 0; lev <= finest_level; ++lev) {
        const MxSphr sp = mxadv::sphr_of(geom[lev], base_geom, center);
        for (MFIter mfi(scal_force[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const MxFab f = mxadv::fab_of(scal_force[lev].array(mfi));
            const MxFab st = mxadv::fab_of(state[lev].const_array(mfi));
            const mxadv::Faces um = mxadv::faces_of(umac_in[lev], mfi);
            const MxFab s0c = mxadv::fab_of(s0_cart[lev].const_array(mfi));
            const MxFab w0c = mxadv::fab_of(w0_cart[lev].const_array(mfi));
            const MxFab ec = mxadv::fab_of(s0_edge_cart[lev].const_array(mfi));
            mxadv::check(mx_modify_scal_force(&g, &f, &st, um.f, &be, lev, &s0c, spherical ? nullptr : &w0c,
                                              spherical ? w0p : nullptr, comp, fullform, MX_BC_FOEXTRAP,
                                              spherical ? nullptr : &ec, &sp, mxadv::stream()),
                         "ModifyScalForce");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();  // s0_edge_cart goes out of scope
#endif
    AverageDown(scal_force, comp, 1);                                                   // :462
    FillPatch(t_old, scal_force, scal_force, scal_force, comp, comp, 1, 0, bcs_f);      // :465
}
