// mx_flux.h — (rho X) and (rho h) flux kernels; see mx_flux.cu
#ifndef MX_FLUX_H
#define MX_FLUX_H
#include "mx_common.h"

// spherical face arrays (time-centred on the fly: 0.5*old + 0.5*new)
struct KMac {
    DFab r0_old[3], r0_new[3], rh0_old[3], rh0_new[3], h0_old[3], h0_new[3];
};

void flux_rhoX(const KGeom& g, const KFlux& f, const DFab* sflux, const DFab& etarho,
               const DFab* sedge, const DFab* umac, const KMac& mac, cudaStream_t stream);
void flux_rhoh(const KGeom& g, const KFlux& f, const DFab* sflux, const DFab* sedge,
               const DFab* umac, const KMac& mac, cudaStream_t stream);
// out[like.off(face)] = 0.5*r0o + 0.5*r0n over the valid faces of direction d
void mac_average(const KGeom& g, int d, const DFab& r0o, const DFab& r0n, const DFab& like,
                 double* out, cudaStream_t stream);
void cell_to_edge(const double* cell, double* edge, int nr, int r_lo, int r_hi,
                  cudaStream_t stream);
void halo_pack(const KGeom& g, const DFab& fab, int dir, int side, int ng, int comp0, int ncomp,
               double* buf, bool unpack, cudaStream_t stream);
#endif
