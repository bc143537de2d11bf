// mx_fused.cu — placeholder until the fused kernels land (everything routes to mx_staged.cu)
#include "mx_fused.h"
bool fused_edge_supported(const KGeom&, const KEdge&) { return false; }
int fused_edge_scal(const KGeom&, const KEdge&, const MxEdgeParams*, const DFab&, const DFab*,
                    const DFab*, const DFab&, cudaStream_t) { return -1; }
bool fused_density_supported(const KGeom&, const KEdge&, const KFlux&) { return false; }
int fused_density(const KGeom&, const KEdge&, const MxEdgeParams*, int, const KFlux&, const DFab&,
                  const DFab*, const DFab*, const DFab&, const DFab*, const DFab&, const KMac&,
                  cudaStream_t) { return -1; }
