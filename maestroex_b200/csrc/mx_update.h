// mx_update.h — Maestro::UpdateScal zone loop; see mx_update.cu
#ifndef MX_UPDATE_H
#define MX_UPDATE_H
#include "mx_common.h"

void update_scal(const KGeom& g, const DFab& sold, const DFab& snew, const DFab* sflux,
                 const DFab& force, const MxUpdateParams& p, unsigned char* mask,
                 unsigned long long* count, cudaStream_t stream);
#endif
