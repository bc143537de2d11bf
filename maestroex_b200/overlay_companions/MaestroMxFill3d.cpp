// Overlay for the 1-D -> 3-D mappings of Source/MaestroFill3dData.cpp on libmxadv.so:
// Put1dArrayOnCart (:9-337), Addw0 (:357-461), MakeW0mac (:464-710), MakeS0mac (:712-1188).
// The zone / face loops run on the GPU; AverageDown, FillPatch, FillBoundary, FillUmacGhost,
// AverageDownFaces, FillPatchUedge are the reference's own calls in the reference's order.
// Build: compile the reference's MaestroFill3dData.cpp with
//   -DPut1dArrayOnCart=Put1dArrayOnCart_reference -DAddw0=Addw0_reference
//   -DMakeW0mac=MakeW0mac_reference -DMakeS0mac=MakeS0mac_reference
// and add this file to CEXE_sources.
#include <MxadvSphr.H>

using namespace amrex;

void Maestro::Put1dArrayOnCart(const BaseState<Real>& s0, Vector<MultiFab>& s0_cart,
                               const bool is_input_edge_centered, const bool is_output_a_vector,
                               const Vector<BCRec>& bcs, const int sbccomp, const int variable_type) {
    BL_PROFILE_VAR("Maestro::Put1dArrayOnCart()", Put1dArrayOnCart);
    int ng = s0_cart[0].nGrow();
    if (ng > 0 && bcs.empty()) {
        Abort("Put1dArrayOnCart with ghost cells requires bcs input");
    }
    for (int lev = 0; lev <= finest_level; ++lev) {
        Put1dArrayOnCart(lev, s0, s0_cart, is_input_edge_centered, is_output_a_vector, bcs, sbccomp);
    }
    int ncomp = is_output_a_vector ? AMREX_SPACEDIM : 1;
    AverageDown(s0_cart, 0, ncomp);
    if (ng > 0) {
        FillPatch(t_old, s0_cart, s0_cart, s0_cart, 0, 0, ncomp, sbccomp, bcs, variable_type);
    }
}

void Maestro::Put1dArrayOnCart(const int lev, const BaseState<Real>& s0, Vector<MultiFab>& s0_cart,
                               const bool is_input_edge_centered, const bool is_output_a_vector,
                               const Vector<BCRec>& bcs, const int sbccomp) {
    Put1dArrayOnCart(lev, s0, s0_cart[lev], is_input_edge_centered, is_output_a_vector, bcs, sbccomp);
}

void Maestro::Put1dArrayOnCart(const int lev, const BaseState<Real>& s0, MultiFab& s0_cart,
                               const bool is_input_edge_centered, const bool is_output_a_vector,
                               [[maybe_unused]] const Vector<BCRec>& bcs,
                               [[maybe_unused]] const int sbccomp) {
    BL_PROFILE_VAR("Maestro::Put1dArrayOnCart_lev()", Put1dArrayOnCart);
    mxadv::no_exact_base_state("Put1dArrayOnCart");
    const MxSphr sp = mxadv::sphr_of(geom[lev], base_geom, center);
    const MxBase b = mxadv::base_of(s0);
    for (MFIter mfi(s0_cart, false); mfi.isValid(); ++mfi) {
        const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
        const MxFab f = mxadv::fab_of(s0_cart.array(mfi));
        mxadv::check(mx_put_1d_array_on_cart(&g, &b, lev, &f, is_input_edge_centered ? 1 : 0,
                                             is_output_a_vector ? 1 : 0, &sp, mxadv::stream()),
                     "Put1dArrayOnCart");
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
}

void Maestro::Addw0(Vector<std::array<MultiFab, AMREX_SPACEDIM> >& u_edge,
                    [[maybe_unused]] const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac,
                    const Real& mult) {
    BL_PROFILE_VAR("Maestro::Addw0()", Addw0);
    const MxBase bw = mxadv::base_of(w0);
    for (int lev = 0; lev <= finest_level; ++lev) {
        for (MFIter mfi(sold[lev]); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            mxadv::Faces ue = mxadv::faces_of(u_edge[lev], mfi);
            if (spherical) {
                const mxadv::Faces wm = mxadv::faces_of(w0mac[lev], mfi);
                mxadv::check(mx_addw0(&g, ue.f, wm.f, nullptr, lev, mult, 1, mxadv::stream()), "Addw0");
            } else {
                mxadv::check(mx_addw0(&g, ue.f, nullptr, &bw, lev, mult, 0, mxadv::stream()), "Addw0");
            }
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
    if (finest_level == 0) {  // :422-459, the reference's ghost fills
        for (int lev = 0; lev <= finest_level; ++lev) {
            for (int d = 0; d < AMREX_SPACEDIM; ++d) {
                u_edge[lev][d].FillBoundary(geom[lev].periodicity());
            }
        }
        FillUmacGhost(u_edge);
    } else {
        AverageDownFaces(u_edge);
        FillPatchUedge(u_edge);
    }
}

#if (AMREX_SPACEDIM == 3)
void Maestro::MakeW0mac(Vector<std::array<MultiFab, AMREX_SPACEDIM> >& w0mac) {
    BL_PROFILE_VAR("Maestro::MakeW0mac()", MakeW0mac);
    if (!spherical) {
        Abort("Error: only call MakeW0mac for spherical");
    }
    if (w0mac[0][0].nGrow() != 1) {
        Abort("Error: MakeW0mac assumes one ghost cell");
    }
    const Real* w0p = w0.const_array().dataPtr();
    for (int lev = 0; lev <= finest_level; ++lev) {
        const MxSphr sp = mxadv::sphr_of(geom[lev], base_geom, center);
        for (MFIter mfi(w0_cart[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const mxadv::Faces wm = mxadv::faces_of(w0mac[lev], mfi);
            const MxFab wc = mxadv::fab_of(w0_cart[lev].const_array(mfi));
            mxadv::check(mx_make_w0mac(&g, wm.f, &wc, w0p, &sp, mxadv::stream()), "MakeW0mac");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
}

void Maestro::MakeS0mac(const BaseState<Real>& s0, Vector<std::array<MultiFab, AMREX_SPACEDIM> >& s0mac) {
    BL_PROFILE_VAR("Maestro::MakeS0mac()", MakeS0mac);
    if (!spherical) {
        Abort("Error: only call MakeS0mac for spherical");
    }
    mxadv::no_exact_base_state("MakeS0mac");
    // the 2-ghost cell-centred mapping, ghost cells by the reference's own FillPatch (:721-730)
    Vector<MultiFab> s0_cart(finest_level + 1);
    if (s0mac_interp_type == 1) {
        for (int lev = 0; lev <= finest_level; ++lev) {
            s0_cart[lev].define(grids[lev], dmap[lev], 1, 2);
            s0_cart[lev].setVal(0.);
        }
        Put1dArrayOnCart(s0, s0_cart, false, false, bcs_f, 0);
    }
    if (s0mac[0][0].nGrow() != 1) {
        Abort("Error: MakeS0mac assumes one ghost cell");
    }
    const MxBase b = mxadv::base_of(s0);
    for (int lev = 0; lev <= finest_level; ++lev) {
        const MxSphr sp = mxadv::sphr_of(geom[lev], base_geom, center);
        for (MFIter mfi(sold[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const mxadv::Faces sm = mxadv::faces_of(s0mac[lev], mfi);
            MxFab sc;
            if (s0mac_interp_type == 1) sc = mxadv::fab_of(s0_cart[lev].const_array(mfi));
            mxadv::check(mx_make_s0mac(&g, &b, sm.f, s0mac_interp_type == 1 ? &sc : nullptr, &sp,
                                       mxadv::stream()),
                         "MakeS0mac");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();  // s0_cart goes out of scope
#endif
}
#endif
