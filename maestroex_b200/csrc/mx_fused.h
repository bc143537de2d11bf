// mx_fused.h — fused single-pass edge-state (+flux) kernels; see mx_fused.cu
#ifndef MX_FUSED_H
#define MX_FUSED_H
#include "mx_common.h"
#include "mx_flux.h"

bool fused_edge_supported(const KGeom& g, const KEdge& e);
int fused_edge_scal(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, const DFab& state,
                    const DFab* sedge, const DFab* umac, const DFab& force, cudaStream_t stream);
bool fused_density_supported(const KGeom& g, const KEdge& e, const KFlux& f);
int fused_density(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, int rho_bccomp,
                  const KFlux& f, const DFab& state, const DFab* sedge, const DFab* sflux,
                  const DFab& etarho, const DFab* umac, const DFab& force, const KMac& mac,
                  double* const* r0e /*[3], spherical: mac_average output, else null*/,
                  cudaStream_t stream);
bool fused_enthalpy_supported(const KGeom& g, const KEdge& e, const KFlux& f);
int fused_enthalpy(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, const KFlux& f,
                   const DFab& state, const DFab* sedge, const DFab* sflux, const DFab* umac,
                   const DFab& force, double* const* r0e /*[3] spherical, else null*/,
                   double* const* h0e /*[3] spherical, else null*/, cudaStream_t stream);
#endif
