// Overlay for Source/MaestroCelltoEdge.cpp: Maestro::CelltoEdge (Source/Maestro.H:401-402)
// on libmxadv.so (mx_cell_to_edge), one call per radial level and disjoint chunk
// (base_geom.r_start_coord / r_end_coord), followed by the reference's own RestrictBase.
#include <MxadvAmrex.H>

using namespace amrex;

void Maestro::CelltoEdge(const BaseState<Real>& s0_cell_s, BaseState<Real>& s0_edge_s) const {
    BL_PROFILE_VAR("Maestro::CelltoEdge()", CelltoEdge);

    if (spherical) {
        Abort("Calling CelltoEdge with spherical == true");  // MaestroCelltoEdge.cpp:10-12
    }
    auto s0_cell = s0_cell_s.const_array();
    auto s0_edge = s0_edge_s.array();
    const int len_c = s0_cell.length(), len_e = s0_edge.length();

    for (int n = 0; n <= base_geom.finest_radial_level; ++n) {
        const int nr = base_geom.nr(n);
        for (int i = 1; i <= base_geom.numdisjointchunks(n); ++i) {
            const int lo = base_geom.r_start_coord(n, i);
            const int hi = base_geom.r_end_coord(n, i);
            // level n of a 1-component BaseState is contiguous: (lev*len + r)*nvar, nvar == 1
            const Real* cell = s0_cell.dataPtr() + static_cast<long>(n) * len_c;
            Real* edge = s0_edge.dataPtr() + static_cast<long>(n) * len_e;
            mxadv::check(mx_cell_to_edge(cell, edge, nr, lo, hi, mxadv::stream()), "CelltoEdge");
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();  // RestrictBase reads the edge array on the host
#endif
    // make the edge values synchronous across levels
    RestrictBase(s0_edge_s, false);
}
