// mx_velpred.h — launchers of mx_velpred.cu (MakeUtrans, VelPred); device pointers only.
#ifndef MX_VELPRED_H
#define MX_VELPRED_H
#include "mx_common.h"

// runtime parameters of the MAC-velocity predictor (Source/MaestroMakeUtrans.cpp,
// Source/MaestroVelPred.cpp): dt, rel_eps, maestro.ppm_type / slope_order / ppm_trace_forces /
// spherical, phys_bc (amrex::PhysBCType per face) and bcs_u (BCRec per velocity component)
struct KVel {
    double dt, eps;
    int ppm_type, slope_order, trace, spherical;
    int pbc_lo[3], pbc_hi[3];
    int bclo[3][3], bchi[3][3];  // [component][direction]
    // x / h_d as a correctly rounded multiply-and-refine sequence (mx_stream.cuh: div_h)
    double h[3], inv_h[3];
    int inv_exact[3];
};

// doubles of device workspace vel_pred() needs for this box
size_t vel_pred_workspace(const KGeom& g);

void vel_make_utrans(const KGeom& g, const KVel& e, const DFab& utilde, const DFab& ufull,
                     const DFab* utrans, const DFab& w0c, const DFab* w0mac, cudaStream_t st);
void vel_pred(const KGeom& g, const KVel& e, const DFab& utilde, const DFab& ufull,
              const DFab* utrans, const DFab* umac, const DFab* w0mac, const DFab& w0c,
              const DFab& force, double* ws, cudaStream_t st);
#endif
