// mx_fluxdev.cuh — per-face flux arithmetic of MakeRhoXFlux / MakeRhoHFlux, shared by the
// streaming flux kernels (mx_flux.cu) and the fused edge+flux kernels (mx_fused.cu).
#ifndef MX_FLUXDEV_CUH
#define MX_FLUXDEV_CUH
#include "mx_common.h"

namespace mxd {

// time-centred base-state density at a face of direction d with radial index r (planar):
// Source/MaestroMakeFlux.cpp:198-199 (x,y faces) and :264-265 (vertical faces)
__device__ __forceinline__ double rho0_edge_planar(const KFlux& f, bool vertical, int r) {
    return vertical ? 0.5 * (f.rho0_edge_old[r * f.st_rho0_edge_old] +
                             f.rho0_edge_new[r * f.st_rho0_edge_new])
                    : 0.5 * (f.rho0_old[r * f.st_rho0_old] + f.rho0_new[r * f.st_rho0_new]);
}

// species flux, MaestroMakeFlux.cpp:201-220; returns false if the reference leaves the entry
__device__ __forceinline__ double rhoX_flux(int spt, double u, double rho0_edge, double ser,
                                            double se) {
    if (spt == MX_PRED_RHOPRIME_AND_X) return u * (rho0_edge + ser) * se;
    if (spt == MX_PRED_RHOX) return u * se;
    return u * ser * se;  // MX_PRED_RHO_AND_X
}

// planar (rho h) flux, MaestroMakeFlux.cpp:674-786 (and 2-D :600-669).  `ser` is only read
// for the h-type forms.  Returns false when the reference leaves sflux(RhoH) untouched.
__device__ __forceinline__ bool rhoh_flux_planar(const KFlux& f, bool vertical, int r, double u,
                                                 double ser, double seh, double* F) {
    const int ept = f.ept, spt = f.spt;
    const bool have_h =
        ept == MX_PRED_H || ept == MX_PRED_T_THEN_H || ept == MX_PRED_TPRIME_THEN_H;
    if (have_h) {
        if (spt == MX_PRED_RHOPRIME_AND_X) {
            *F = u * (rho0_edge_planar(f, vertical, r) + ser) * seh;
            return true;
        }
        if (spt == MX_PRED_RHO_AND_X || spt == MX_PRED_RHOX) {
            *F = u * ser * seh;
            return true;
        }
        return false;
    }
    if (ept == MX_PRED_RHOH) {
        *F = u * seh;
        return true;
    }
    if (ept == MX_PRED_RHOHPRIME || ept == MX_PRED_T_THEN_RHOHPRIME) {
        if (vertical) {
            const double rhoh0_edge = 0.5 * (f.rhoh0_edge_old[r * f.st_rhoh0_edge_old] +
                                             f.rhoh0_edge_new[r * f.st_rhoh0_edge_new]);
            *F = u * (seh + rhoh0_edge);  // :782-784
        } else {
            const double rhoh0_edge =
                0.5 * (f.rhoh0_old[r * f.st_rhoh0_old] + f.rhoh0_new[r * f.st_rhoh0_new]);
            *F = u * (rhoh0_edge + seh);  // :706-708
        }
        return true;
    }
    return false;
}

// the same with the time-centred base-state values already formed (per-plane tables of the
// streaming kernel): rho0e = 0.5 (rho0[_edge]_old + rho0[_edge]_new), rhoh0e likewise
__device__ __forceinline__ bool rhoh_flux_tab(const KFlux& f, bool have_h, double u, double ser,
                                              double seh, double rho0e, double rhoh0e,
                                              bool vertical, double* F) {
    const int ept = f.ept, spt = f.spt;
    if (have_h) {
        if (spt == MX_PRED_RHOPRIME_AND_X) {
            *F = u * (rho0e + ser) * seh;
            return true;
        }
        if (spt == MX_PRED_RHO_AND_X || spt == MX_PRED_RHOX) {
            *F = u * ser * seh;
            return true;
        }
        return false;
    }
    if (ept == MX_PRED_RHOH) {
        *F = u * seh;
        return true;
    }
    if (ept == MX_PRED_RHOHPRIME || ept == MX_PRED_T_THEN_RHOHPRIME) {
        *F = vertical ? u * (seh + rhoh0e) : u * (rhoh0e + seh);
        return true;
    }
    return false;
}

// spherical, use_exact_base_state = 0 (MaestroMakeFlux.cpp:871-1021): rho0e / h0e are the
// time-centred base state on this face, 0.5*old + 0.5*new of the r0mac / h0mac face arrays
__device__ __forceinline__ bool rhoh_flux_sph(const KFlux& f, double u, double ser, double seh,
                                              double rho0e, double h0e, double* F) {
    const int ept = f.ept, spt = f.spt;
    if (ept == MX_PRED_H || ept == MX_PRED_T_THEN_H || ept == MX_PRED_TPRIME_THEN_H) {
        if (spt == MX_PRED_RHOPRIME_AND_X) {
            *F = u * (rho0e + ser) * seh;
            return true;
        }
        if (spt == MX_PRED_RHO_AND_X || spt == MX_PRED_RHOX) {
            *F = u * ser * seh;
            return true;
        }
        return false;
    }
    if (ept == MX_PRED_HPRIME) {
        if (spt == MX_PRED_RHOPRIME_AND_X) {
            *F = u * (ser + rho0e) * (seh + h0e);
            return true;
        }
        return false;
    }
    if (ept == MX_PRED_RHOH) {
        *F = u * seh;
        return true;
    }
    *F = u * (rho0e * h0e + seh);
    return true;
}

__device__ __forceinline__ bool needs_sedge_rho_for_rhoh(const KFlux& f) {
    return f.ept == MX_PRED_H || f.ept == MX_PRED_T_THEN_H || f.ept == MX_PRED_TPRIME_THEN_H;
}

}  // namespace mxd
#endif
