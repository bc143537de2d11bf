// mx_w0.h — Maestro::Makew0 kernels; see mx_w0.cu
#ifndef MX_W0_H
#define MX_W0_H
#include <cstddef>

#include "mx_common.h"

size_t w0_scratch_doubles(int nr_fine);
void w0_planar(const MxW0Params& P, const double* w0_old, const double* Sbar,
               const double* p0_old, const double* p0_new, const double* g1o, const double* g1n,
               const double* pmp, const double* etarho_cc, double* w0, double* w0_force,
               double* scratch, double* p0bdot_dev, cudaStream_t stream);
void w0_sphr(const MxW0Params& P, const double* w0_old, const double* Sbar,
             const double* rho0_old, const double* rho0_new, const double* p0_old,
             const double* p0_new, const double* g1o, const double* g1n, const double* pmp,
             const double* etarho_cc, const double* etarho_ec, const double* rcc,
             const double* re, double* w0, double* w0_force, double* scratch, int* status_dev,
             bool irreg, cudaStream_t stream);
void w0_planar_ml(const MxW0Params& P, int max_lev, int finest_lev, int nrf, int chunk_len,
                  const double* dr, const int* nr, const int* cut, const int* nchunk, const int* rs,
                  const int* re, const double* w0_old, const double* Sbar, const double* p0_old,
                  const double* p0_new, const double* g1o, const double* g1n, const double* pmp,
                  const double* etarho_cc, double* w0, double* w0_force, double* scratch,
                  double* p0bdot_dev, cudaStream_t stream);
#endif
