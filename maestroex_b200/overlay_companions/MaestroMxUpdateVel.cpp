// Overlay for Maestro::UpdateVel (Source/MaestroAdvection.cpp:299-526) on libmxadv.so
// (mx_update_vel); the AverageDown / FillPatch tail is unchanged.
// Build: compile the reference's MaestroAdvection.cpp with -DUpdateVel=UpdateVel_reference (next
// to -DUpdateScal=UpdateScal_reference of maestroex_b200/overlay) and add this file to CEXE_sources.
#include <MxadvAmrex.H>

using namespace amrex;

void Maestro::UpdateVel(const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
                        const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& uedge,
                        const Vector<MultiFab>& force, const Vector<MultiFab>& sponge,
                        [[maybe_unused]] const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac) {
    BL_PROFILE_VAR("Maestro::UpdateVel()", UpdateVel);
    for (int lev = 0; lev <= finest_level; ++lev) {
        for (MFIter mfi(force[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const MxFab uo = mxadv::fab_of(uold[lev].const_array(mfi));
            const MxFab un = mxadv::fab_of(unew[lev].array(mfi));
            const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
            const mxadv::Faces ue = mxadv::faces_of(uedge[lev], mfi);
            const MxFab fo = mxadv::fab_of(force[lev].const_array(mfi));
            const MxFab sp = mxadv::fab_of(sponge[lev].const_array(mfi));
            const MxFab wc = mxadv::fab_of(w0_cart[lev].const_array(mfi));
            mxadv::Faces wm;
            if (spherical) wm = mxadv::faces_of(w0mac[lev], mfi);
            mxadv::check(mx_update_vel(&g, &uo, &un, um.f, ue.f, &fo, do_sponge ? &sp : nullptr,
                                       spherical ? nullptr : &wc, spherical ? wm.f : nullptr, dt,
                                       do_sponge ? 1 : 0, spherical ? 1 : 0, mxadv::stream()),
                         "UpdateVel");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
    AverageDown(unew, 0, AMREX_SPACEDIM);                                          // :522
    FillPatch(t_old, unew, unew, unew, 0, 0, AMREX_SPACEDIM, 0, bcs_u, 1);        // :525
}
