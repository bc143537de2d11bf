// Overlay for Maestro::MakeUtrans (Source/MaestroMakeUtrans.cpp:6-648) and Maestro::VelPred
// (Source/MaestroVelPred.cpp:5-294 with VelPredInterface / VelPredTransverse /
// VelPredVelocities) on libmxadv.so (mx_make_utrans, mx_vel_pred).  The ghost-fill / AMR tails
// are the reference's own calls in the reference's order.
// Build: these two members are the only non-static ones their translation units define that
// anything else calls, so the overlay simply shadows MaestroMakeUtrans.cpp and MaestroVelPred.cpp
// (this file is added to CEXE_sources and the two reference files are dropped from it: see
// Make.package), or, where the reference files must stay, they are compiled with
// -DMakeUtrans=MakeUtrans_reference -DVelPred=VelPred_reference.
#include <MxadvAmrex.H>

using namespace amrex;

namespace {
// phys_bc / bcs_u / runtime parameters -> MxVelPredParams (bc points into `table`)
MxVelPredParams vel_params(const BCRec& phys_bc, const Vector<BCRec>& bcs_u, Real dt, Real rel_eps,
                           int trace, std::array<int, 6 * AMREX_SPACEDIM>& table) {
    MxVelPredParams p;
    p.dt = dt;
    p.rel_eps = rel_eps;
    p.ppm_type = ppm_type;
    p.slope_order = slope_order;
    p.ppm_trace_forces = trace;
    p.spherical = spherical ? 1 : 0;
    for (int d = 0; d < 3; ++d) {
        p.phys_bc_lo[d] = d < AMREX_SPACEDIM ? phys_bc.lo(d) : 0;
        p.phys_bc_hi[d] = d < AMREX_SPACEDIM ? phys_bc.hi(d) : 0;
    }
    table.fill(0);
    for (int n = 0; n < AMREX_SPACEDIM; ++n) {
        for (int d = 0; d < AMREX_SPACEDIM; ++d) {
            table[6 * n + d] = bcs_u[n].lo(d);
            table[6 * n + 3 + d] = bcs_u[n].hi(d);
        }
    }
    p.bc = table.data();
    return p;
}
}  // namespace

void Maestro::MakeUtrans(const Vector<MultiFab>& utilde, const Vector<MultiFab>& ufull,
                         Vector<std::array<MultiFab, AMREX_SPACEDIM> >& utrans,
                         [[maybe_unused]] const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac) {
    BL_PROFILE_VAR("Maestro::MakeUtrans()", MakeUtrans);
    std::array<int, 6 * AMREX_SPACEDIM> table;
    const MxVelPredParams par = vel_params(phys_bc, bcs_u, dt, rel_eps, 0, table);
    for (int lev = 0; lev <= finest_level; ++lev) {
        for (MFIter mfi(utilde[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const MxFab ut = mxadv::fab_of(utilde[lev].const_array(mfi));
            const MxFab uf = mxadv::fab_of(ufull[lev].const_array(mfi));
            const mxadv::Faces tr = mxadv::faces_of(utrans[lev], mfi);
            const MxFab wc = mxadv::fab_of(w0_cart[lev].const_array(mfi));
            mxadv::Faces wm;
            if (spherical) wm = mxadv::faces_of(w0mac[lev], mfi);
            mxadv::check(mx_make_utrans(&g, &ut, &uf, tr.f, spherical ? nullptr : &wc,
                                        spherical ? wm.f : nullptr, &par, mxadv::stream()),
                         "MakeUtrans");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
    if (finest_level == 0) {  // :630-647
        for (int lev = 0; lev <= finest_level; ++lev) {
            for (int d = 0; d < AMREX_SPACEDIM; ++d) {
                utrans[lev][d].FillBoundary(geom[lev].periodicity());
            }
        }
        FillUmacGhost(utrans);
    } else {
        AverageDownFaces(utrans);
        FillPatchUedge(utrans);
    }
}

void Maestro::VelPred(Vector<MultiFab>& utilde, const Vector<MultiFab>& ufull,
                      const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& utrans,
                      Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
                      const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac,
                      const Vector<MultiFab>& force) {
    amrex::ignore_unused(w0mac);
    BL_PROFILE_VAR("Maestro::VelPred()", VelPred);
    std::array<int, 6 * AMREX_SPACEDIM> table;
    const MxVelPredParams par = vel_params(phys_bc, bcs_u, dt, rel_eps, ppm_trace_forces, table);
    for (int lev = 0; lev <= finest_level; ++lev) {
        for (MFIter mfi(utilde[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const MxFab ut = mxadv::fab_of(utilde[lev].const_array(mfi));
            const MxFab uf = mxadv::fab_of(ufull[lev].const_array(mfi));
            const mxadv::Faces tr = mxadv::faces_of(utrans[lev], mfi);
            const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
            const MxFab fo = mxadv::fab_of(force[lev].const_array(mfi));
            const MxFab wc = mxadv::fab_of(w0_cart[lev].const_array(mfi));
            mxadv::Faces wm;
            if (spherical) wm = mxadv::faces_of(w0mac[lev], mfi);
            mxadv::check(mx_vel_pred(&g, &ut, &uf, tr.f, um.f, spherical ? wm.f : nullptr,
                                     spherical ? nullptr : &wc, &fo, &par, mxadv::stream()),
                         "VelPred");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
    AverageDownFaces(umac);  // :293
}
