// mx_common.h — host/device POD views shared by the kernels and the C-ABI layer.
#ifndef MX_COMMON_H
#define MX_COMMON_H

#include <cuda_runtime.h>

#include "../../include/mxadv.h"

// Device-side view of one fab (Array4 order, i fastest; AMReX_Array4.H:86-91).
struct DFab {
    double* p;
    int lo[3];
    int n[3];  // extents
    long sy, sz, sn;  // strides of j, k and component
    int ncomp;

    __host__ __device__ inline long off(int i, int j, int k) const {
        return (long)(i - lo[0]) + (long)(j - lo[1]) * sy + (long)(k - lo[2]) * sz;
    }
    __host__ __device__ inline double* comp(int c) const { return p + (long)c * sn; }
};

inline DFab make_dfab(const MxFab* f, double* dev_ptr) {
    DFab d;
    d.p = dev_ptr;
    for (int q = 0; q < 3; ++q) {
        d.lo[q] = f->lo[q];
        d.n[q] = f->hi[q] - f->lo[q] + 1;
    }
    d.sy = d.n[0];
    d.sz = (long)d.n[0] * d.n[1];
    d.sn = d.sz * d.n[2];
    d.ncomp = f->ncomp;
    return d;
}

struct KGeom {
    int dim;
    int vlo[3], vhi[3];  // valid cell box
    int dlo[3], dhi[3];  // domain
    double dx[3];
    // 1: the streaming kernel leaves the z faces at vhi[2]+1 to the box above (the slab
    // pipeline of mx_api.cu: every face plane is then written by exactly one slab)
    int skip_top_z;
};

inline KGeom make_kgeom(const MxGeom* g) {
    KGeom k;
    k.dim = g->dim;
    k.skip_top_z = 0;
    for (int d = 0; d < 3; ++d) {
        const bool in = d < g->dim;
        k.vlo[d] = in ? g->valid_lo[d] : 0;
        k.vhi[d] = in ? g->valid_hi[d] : 0;
        k.dlo[d] = in ? g->dom_lo[d] : 0;
        k.dhi[d] = in ? g->dom_hi[d] : 0;
        k.dx[d] = in ? g->dx[d] : 1.0;
    }
    return k;
}

// Per-sweep parameters (one component) for the edge-state kernels.
struct KEdge {
    double dt, eps;
    int ppm_type, slope_order, trace, is_vel, is_cons;
    int comp;  // component of state / force / sedge being swept
    int bclo[3], bchi[3];
};

// Flux parameters + base-state device pointers.
struct KFlux {
    int spt, ept, spherical, exact, evolve;
    int rho, rhoh, first_spec, num_spec, start_comp, num_comp;
    // base-state arrays, already offset to level `lev`; nvar stride
    const double *rho0_old, *rho0_new, *rho0_edge_old, *rho0_edge_new, *rho0_pred_edge, *w0;
    const double *rhoh0_old, *rhoh0_new, *rhoh0_edge_old, *rhoh0_edge_new;
    int st_rho0_old, st_rho0_new, st_rho0_edge_old, st_rho0_edge_new, st_rho0_pred_edge, st_w0;
    int st_rhoh0_old, st_rhoh0_new, st_rhoh0_edge_old, st_rhoh0_edge_new;
};

// sets mx_last_error() of the calling thread and returns `code` (mx_api.cu)
int mx_fail(int code, const char* fmt, ...);
void mx_count_launch(int n = 1);
void mx_count_tma_launch();

#endif
