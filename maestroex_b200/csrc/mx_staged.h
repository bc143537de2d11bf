// mx_staged.h — staged (general) edge-state path; see mx_staged.cu
#ifndef MX_STAGED_H
#define MX_STAGED_H
#include "mx_common.h"

// Workspace arrays, all defined on grow(valid, 1) like the reference's 1-ghost temporaries
// (Source/MaestroMakeEdgeScalars.cpp:24-64).
struct Scratch {
    double *Ip[3], *Im[3], *Ipf[3], *Imf[3];
    double *sl[3], *sr[3], *simh[3];
    double* tr[3][3];  // tr[d][t]: state on d-faces corrected by the t-transverse term
    double* divu;
    int lo[3], n[3];
    long sy, sz;
};

size_t staged_scratch_doubles(const KGeom& g);
void staged_make_edge_scal(const KGeom& g, const KEdge& e0, const MxEdgeParams* par,
                           const DFab& state, const DFab* sedge, const DFab* umac,
                           const DFab& force, double* scratch, cudaStream_t stream);
#endif
