// mx_cart.h — launchers of mx_cart.cu (conversions, perturbational form, rho' forcing, the
// 1-D -> 3-D base-state mappings, velocity update); device pointers only.
#ifndef MX_CART_H
#define MX_CART_H
#include "mx_common.h"

// device-side copy of MxSphr (r_cc_loc / r_edge_loc already on the device)
struct KSphr {
    int spherical, nr;
    double drf;
    double plo[3], cen[3];
    const double* rcc;
    const double* redge;
    int s0t, w0t, s0mt, w0mt;
};

void cart_convert_by_rho(const KGeom& g, const DFab& s, int rho, int c0, int nc, int flag,
                         cudaStream_t st);
void cart_copy_valid(const KGeom& g, const DFab& src, const DFab& dst, int c0, int nc, cudaStream_t st);
void cart_put_1d(const KGeom& g, const double* s0, int st0, const DFab& out, int edge, int vec,
                 int mode, int comp, const KSphr& S, cudaStream_t st);
void cart_modify_scal_force(const KGeom& g, const DFab& force, const DFab& state, const DFab* umac,
                            const double* s0e, int st_e, const DFab& s0c, const DFab& w0c,
                            const double* w0, double* divw0_ws, int comp, int fullform, int bcf_hi,
                            const DFab* s0ec, const KSphr& S, cudaStream_t st);
void cart_addw0(const KGeom& g, const DFab* ue, const DFab* w0mac, const double* w0, int st_w,
                double mult, int spherical, cudaStream_t st);
void cart_make_w0mac(const KGeom& g, const DFab* w0mac, const DFab& w0c, const double* w0,
                     const KSphr& S, cudaStream_t st);
void cart_make_s0mac(const KGeom& g, const double* s0, int st0, const DFab* s0mac, const DFab* s0c,
                     const KSphr& S, cudaStream_t st);
void cart_update_vel(const KGeom& g, const DFab& uold, const DFab& unew, const DFab* umac,
                     const DFab* uedge, const DFab& force, const DFab& sponge, const DFab& w0c,
                     const DFab* w0mac, double dt, int do_sponge, int spherical, cudaStream_t st);
#endif
