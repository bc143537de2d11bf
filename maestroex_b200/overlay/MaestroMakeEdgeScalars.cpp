// Overlay for Source/MaestroMakeEdgeScalars.cpp: same member signature
// (Source/Maestro.H:913-919), work done by libmxadv.so (mx_make_edge_scal: slopes / PPM /
// predictor / transverse / final edges in one fused pass per box, include/mxadv.h).
// None of the reference's 4 x DIM + 17 temporary MultiFabs (:24-64) nor the per-component
// copies (:171-178) are allocated.
#include <MxadvAmrex.H>

using namespace amrex;

void Maestro::MakeEdgeScal(Vector<MultiFab>& state,
                           Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sedge,
                           Vector<std::array<MultiFab, AMREX_SPACEDIM> >& umac,
                           Vector<MultiFab>& force, const bool is_vel,
                           const Vector<BCRec>& bcs, int nbccomp, int start_scomp,
                           int start_bccomp, int num_comp, const bool is_conservative) {
    BL_PROFILE_VAR("Maestro::MakeEdgeScal()", MakeEdgeScal);

    const std::vector<int> bc = mxadv::bc_table(bcs);
    MxEdgeParams par;
    par.dt = dt;
    par.rel_eps = rel_eps;
    par.ppm_type = ppm_type;
    par.slope_order = slope_order;
    par.ppm_trace_forces = ppm_trace_forces;
    par.is_vel = is_vel ? 1 : 0;
    par.is_conservative = is_conservative ? 1 : 0;
    par.start_scomp = start_scomp;
    par.start_bccomp = start_bccomp;
    par.num_comp = num_comp;
    par.nbccomp = static_cast<int>(bcs.size());
    par.bc = bc.data();
    amrex::ignore_unused(nbccomp);

    for (int lev = 0; lev <= finest_level; ++lev) {
        // one library call per box: no tiling, the kernel tiles the box itself
        for (MFIter mfi(state[lev]); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(geom[lev], mfi.validbox());
            const MxFab st = mxadv::fab_of(state[lev].const_array(mfi));
            const MxFab fo = mxadv::fab_of(force[lev].const_array(mfi));
            const mxadv::Faces se = mxadv::faces_of(sedge[lev], mfi);
            const mxadv::Faces um = mxadv::faces_of(umac[lev], mfi);
            mxadv::check(mx_make_edge_scal(&g, &st, se.f, um.f, &fo, &par, mxadv::stream()),
                         "MakeEdgeScal");
        }
    }

    // we use this to prevent the creation of edge states which differ across levels
    // (Source/MaestroMakeEdgeScalars.cpp:297-301; AMR stays the reference's own code)
    if (is_vel && (reflux_type == 1 || reflux_type == 2)) {
        AverageDownFaces(sedge);
    }
}
