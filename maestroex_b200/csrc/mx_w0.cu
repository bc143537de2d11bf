// mx_w0.cu — Maestro::Makew0 on the device (Source/MaestroMakew0.cpp:5-58): the planar
// recurrence (Makew0Planar, :60-272) and the spherical tridiagonal solve (Makew0Sphr,
// :537-740, with MakeGravEdge's spherical branch Source/MaestroMakeGrav.cpp:303-322 and
// Tridiag :942-969), one radial level.
//
// The arrays have nr_fine <= a few thousand entries, so one thread block does the whole job
// in a single launch.  Everything that is independent per radius (time centring, the terms
// of the running sums, the matrix coefficients, w0_force) is evaluated by all threads; the
// running sums and the Thomas recurrence keep the reference's summation order (they must:
// the parity contract is bit-identical results) and are walked by one thread each, the
// independent chains on different warps.  Expression trees follow the reference; the library
// is compiled -fmad=false.
#include "mx_w0.h"

namespace {

constexpr int W0_THREADS = 1024;
constexpr double kGconst = 6.67428e-8;  // Source/Maestro.H:1864
constexpr double kPi = 3.14159265358979323846;  // M_PI

// w0_force, MaestroMakew0.cpp:241-268 / :727-738
__device__ __forceinline__ void w0_force_cells(const MxW0Params& P, const double* w0_old,
                                               const double* w0, double* w0_force) {
    const double dt_in = P.dt_in, dtold_in = P.dtold_in;
    const double dt_avg = 0.5 * (dt_in + dtold_in);
    for (int r = threadIdx.x; r < P.nr_fine; r += blockDim.x) {
        const double w0_old_cen = 0.5 * (w0_old[r] + w0_old[r + 1]);
        const double w0_new_cen = 0.5 * (w0[r] + w0[r + 1]);
        const double w0_avg = 0.5 * (dt_in * w0_old_cen + dtold_in * w0_new_cen) / dt_avg;
        const double div_avg =
            0.5 * (dt_in * (w0_old[r + 1] - w0_old[r]) + dtold_in * (w0[r + 1] - w0[r])) / dt_avg;
        w0_force[r] = (w0_new_cen - w0_old_cen) / dt_avg + w0_avg * div_avg / P.dr;
    }
}

__global__ void __launch_bounds__(W0_THREADS, 1)
k_w0_planar(const MxW0Params P, const double* __restrict__ w0_old,
            const double* __restrict__ Sbar, const double* __restrict__ p0_old,
            const double* __restrict__ p0_new, const double* __restrict__ g1o,
            const double* __restrict__ g1n, const double* __restrict__ pmp,
            const double* __restrict__ etarho_cc, double* __restrict__ w0,
            double* __restrict__ w0_force, double* __restrict__ scratch,
            double* __restrict__ p0bdot_out) {
    const int nr = P.nr_fine, cut = P.base_cutoff_density_coord;
    const double dr = P.dr;
    double* tS = scratch;            // Sbar(r-1) * dr
    double* tP = tS + (nr + 2);      // psi(r-1) / gamma1bar_p0_avg * dr
    double* tC = tP + (nr + 2);      // delta_chi_w0 * dr
    double* tI = tC + (nr + 2);      // (1 / gamma1bar_p0_avg) * dr
    double* int1 = tI + (nr + 2);    // int1_over_gamma1bar_p0
    __shared__ double s_pb;

    // :141-168, the per-radius terms of the recurrence
    for (int r = 1 + threadIdx.x; r <= nr; r += blockDim.x) {
        const double gavg = (g1o[r - 1] + g1n[r - 1]) * (p0_old[r - 1] + p0_new[r - 1]) / 4.0;
        double dchi = 0.0;
        if (r < cut) {
            if (P.is_predictor) {
                dchi = P.dpdt_factor * pmp[r - 1] / (g1o[r - 1] * p0_old[r - 1] * P.dt);
            } else {
                dchi += P.dpdt_factor * pmp[r - 1] / (g1n[r - 1] * p0_new[r - 1] * P.dt);
            }
        }
        const double psi = (r - 1 < cut) ? etarho_cc[r - 1] * fabs(P.grav_const) : 0.0;  // :132-139
        tS[r] = Sbar[r - 1] * dr;
        tP[r] = psi / gavg * dr;
        tC[r] = dchi * dr;
        tI[r] = (1.0 / gavg) * dr;
    }
    __syncthreads();
    // :169-172 and :165-168: two independent running sums, one warp each
    if (threadIdx.x == 0) {
        double w = 0.0;
        w0[0] = 0.0;
        for (int r = 1; r <= nr; ++r) {
            w = w + tS[r] - tP[r] - tC[r];
            w0[r] = w;
        }
    } else if (threadIdx.x == 32) {
        double s = 0.0;
        int1[0] = 0.0;
        for (int r = 1; r <= nr; ++r) {
            s = s + tI[r];
            int1[r] = s;
        }
    }
    __syncthreads();
    if (P.add_pb) {  // :175-184
        if (threadIdx.x == 0) {
            s_pb = w0[nr] / int1[nr];
            if (p0bdot_out) *p0bdot_out = s_pb;
        }
        __syncthreads();
        const double pb = s_pb;
        for (int r = 1 + threadIdx.x; r <= nr; r += blockDim.x) w0[r] -= pb * int1[r];
        __syncthreads();
    }
    w0_force_cells(P, w0_old, w0, w0_force);
}

// IRREG: Makew0SphrIrreg (use_exact_base_state, :742-940): the same algorithm with the local
// spacings dr1 = r_edge(r) - r_edge(r-1), dr2 = r_edge(r+1) - r_edge(r), dr3 = r_cc(r) - r_cc(r-1)
// in place of the uniform dr(0).
template <bool IRREG>
__global__ void __launch_bounds__(W0_THREADS, 1)
k_w0_sphr(const MxW0Params P, const double* __restrict__ w0_old,
          const double* __restrict__ Sbar, const double* __restrict__ rho0_old,
          const double* __restrict__ rho0_new, const double* __restrict__ p0_old,
          const double* __restrict__ p0_new, const double* __restrict__ g1o,
          const double* __restrict__ g1n, const double* __restrict__ pmp,
          const double* __restrict__ etarho_cc, const double* __restrict__ etarho_ec,
          const double* __restrict__ rcc, const double* __restrict__ re,
          double* __restrict__ w0, double* __restrict__ w0_force, double* __restrict__ scratch,
          int* __restrict__ status) {
    const int nr = P.nr_fine;
    const double dr0 = P.dr;
    const int L = nr + 2;
    double* g_nph = scratch;
    double* p_nph = g_nph + L;
    double* rho_nph = p_nph + L;
    double* wS = rho_nph + L;   // w0_from_Sbar
    double* tV = wS + L;        // volume-discrepancy term of the wS sum
    double* tM = tV + L;        // enclosed-mass increments
    double* grav = tM + L;      // grav_edge, first holds mencl
    double* A = grav + L;
    double* B = A + L;
    double* Cc = B + L;
    double* F = Cc + L;
    double* u = F + L;
    double* gam = u + L;
    double* tS = gam + L;       // dr0 * Sbar * r_cc^2
    const int max_cutoff = min(P.base_cutoff_density_coord, nr - 1);

    // :595-600 time-centred base state; terms of the two running sums (:608-622, MakeGrav :307-321)
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const double pn = 0.5 * (p0_old[r] + p0_new[r]);
        const double rn = 0.5 * (rho0_old[r] + rho0_new[r]);
        const double gn = 0.5 * (g1o[r] + g1n[r]);
        p_nph[r] = pn;
        rho_nph[r] = rn;
        g_nph[r] = gn;
        const double vd =
            rho0_old[r] > P.base_cutoff_density ? P.dpdt_factor * pmp[r] / P.dt_in : 0.0;
        const double dr1 = IRREG ? re[r + 1] - re[r] : dr0;  // :812 (cell r = reference index r-1)
        tS[r + 1] = dr1 * Sbar[r] * rcc[r] * rcc[r];
        // the regular routine skips the subtraction when volume_discrepancy == 0; keep the skip
        // as an explicit flag so that signed zeros come out the same (the irregular one always
        // subtracts, :813-817)
        tV[r + 1] = (IRREG || vd != 0.0) ? dr1 * vd * rcc[r] * rcc[r] / (gn * pn) : 0.0;
        A[r + 1] = (IRREG || vd != 0.0) ? 1.0 : 0.0;  // A is rebuilt below; borrowed as the flag array
        const bool dense = rn > P.base_cutoff_density;
        tM[r + 1] = dense ? 4.0 / 3.0 * kPi * (re[r + 1] - re[r]) *
                                (re[r + 1] * re[r + 1] + re[r + 1] * re[r] + re[r] * re[r]) * rn
                          : 0.0;
        B[r + 1] = dense ? 1.0 : 0.0;  // likewise borrowed
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double w = 0.0;
        wS[0] = 0.0;
        for (int r = 1; r <= nr; ++r) {
            w = w + tS[r];
            if (A[r] != 0.0) w -= tV[r];
            wS[r] = w;
        }
    } else if (threadIdx.x == 32) {
        double m = 0.0;
        grav[0] = 0.0;
        for (int r = 1; r <= nr; ++r) {
            if (B[r] != 0.0) m += tM[r];
            grav[r] = m;
        }
    }
    __syncthreads();
    for (int r = 1 + threadIdx.x; r <= nr; r += blockDim.x) {
        const double r2 = re[r] * re[r];
        wS[r] /= r2;                              // :626-629
        grav[r] = -kGconst * grav[r] / r2;        // MakeGrav :319-320
    }
    for (int r = threadIdx.x; r <= nr; r += blockDim.x) {  // :641-645
        A[r] = 0.0;
        B[r] = 0.0;
        Cc[r] = 0.0;
        F[r] = 0.0;
        u[r] = 0.0;
    }
    __syncthreads();
    // :653-681 (regular) / :843-878 (irregular) matrix rows 1 .. max_cutoff
    for (int r = 1 + threadIdx.x; r <= max_cutoff; r += blockDim.x) {
        double a, b, c, f;
        if (IRREG) {
            const double dr1 = re[r] - re[r - 1];
            const double dr2 = re[r + 1] - re[r];
            const double dr3 = rcc[r] - rcc[r - 1];
            a = g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1]);
            a /= dr1 * dr3;
            b = -(g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1] * dr1) +
                  g_nph[r] * p_nph[r] / (rcc[r] * rcc[r] * dr2)) /
                dr3;
            const double dpdr = (p_nph[r] - p_nph[r - 1]) / dr3;
            b -= 4.0 * dpdr / (re[r] * re[r] * re[r]);
            c = g_nph[r] * p_nph[r] / (rcc[r] * rcc[r]);
            c /= dr2 * dr3;
            f = 4.0 * dpdr * wS[r] / re[r] -
                grav[r] *
                    (rcc[r] * rcc[r] * etarho_cc[r] - rcc[r - 1] * rcc[r - 1] * etarho_cc[r - 1]) /
                    (dr3 * re[r] * re[r]) -
                4.0 * kPi * kGconst * 0.5 * (rho_nph[r] + rho_nph[r - 1]) * etarho_ec[r];
        } else {
            a = g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1]);
            a /= dr0 * dr0;
            b = -(g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1]) +
                  g_nph[r] * p_nph[r] / (rcc[r] * rcc[r])) /
                (dr0 * dr0);
            const double dpdr = (p_nph[r] - p_nph[r - 1]) / dr0;
            b -= 4.0 * dpdr / (re[r] * re[r] * re[r]);
            c = g_nph[r] * p_nph[r] / (rcc[r] * rcc[r]);
            c /= dr0 * dr0;
            f = 4.0 * dpdr * wS[r] / re[r] -
                grav[r] *
                    (rcc[r] * rcc[r] * etarho_cc[r] - rcc[r - 1] * rcc[r - 1] * etarho_cc[r - 1]) /
                    (dr0 * re[r] * re[r]) -
                4.0 * kPi * kGconst * 0.5 * (rho_nph[r] + rho_nph[r - 1]) * etarho_ec[r];
        }
        A[r] = a;
        B[r] = b;
        Cc[r] = c;
        F[r] = f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // boundary rows :684-694
        A[0] = 0.0;
        B[0] = 1.0;
        Cc[0] = 0.0;
        F[0] = 0.0;
        A[max_cutoff + 1] = -1.0;
        B[max_cutoff + 1] = 1.0;
        Cc[max_cutoff + 1] = 0.0;
        F[max_cutoff + 1] = 0.0;
        // Tridiag(A, B, C, F, u, max_cutoff + 2), :942-969
        const int n = max_cutoff + 2;
        int bad = 0;
        if (B[0] == 0) {
            bad = 1;  // "tridiag: CAN NOT HAVE B(0) = 0.0"
        } else {
            double bet = B[0];
            double up = F[0] / bet;
            u[0] = up;
            for (int j = 1; j < n; j++) {
                const double gj = Cc[j - 1] / bet;
                gam[j] = gj;
                bet = B[j] - A[j] * gj;
                if (bet == 0) {
                    bad = 2;  // "tridiag: TRIDIAG FAILED"
                    break;
                }
                up = (F[j] - A[j] * up) / bet;
                u[j] = up;
            }
            if (!bad) {
                double un = u[n - 1];
                for (int j = n - 2; j >= 0; --j) {
                    un = u[j] - gam[j + 1] * un;
                    u[j] = un;
                }
            }
        }
        *status = bad;
    }
    __syncthreads();
    // :700-721
    if (threadIdx.x == 0) w0[0] = wS[0];
    const double wtop =
        u[max_cutoff + 1] / (re[max_cutoff + 1] * re[max_cutoff + 1]) + wS[max_cutoff + 1];
    for (int r = 1 + threadIdx.x; r <= nr; r += blockDim.x) {
        if (r <= max_cutoff + 1) {
            w0[r] = u[r] / (re[r] * re[r]) + wS[r];
        } else {
            w0[r] = wtop * re[max_cutoff + 1] * re[max_cutoff + 1] / (re[r] * re[r]);
        }
    }
    __syncthreads();
    if (IRREG) {
        // :924-938: the local spacing is that of the cell BELOW (r_edge(r) - r_edge(r-1)); for
        // r = 0 the reference reads r_edge_loc(0,-1), out of bounds: the first cell's own width
        // is used here and tests do not compare w0_force(0)
        const double dt_in = P.dt_in, dtold_in = P.dtold_in;
        const double dt_avg = 0.5 * (dt_in + dtold_in);
        for (int r = threadIdx.x; r < nr; r += blockDim.x) {
            const double dr1 = r > 0 ? re[r] - re[r - 1] : re[1] - re[0];
            const double w0_old_cen = 0.5 * (w0_old[r] + w0_old[r + 1]);
            const double w0_new_cen = 0.5 * (w0[r] + w0[r + 1]);
            const double w0_avg = 0.5 * (dt_in * w0_old_cen + dtold_in * w0_new_cen) / dt_avg;
            const double div_avg =
                0.5 * (dt_in * (w0_old[r + 1] - w0_old[r]) + dtold_in * (w0[r + 1] - w0[r])) / dt_avg;
            w0_force[r] = (w0_new_cen - w0_old_cen) / dt_avg + w0_avg * div_avg / dr1;
        }
    } else {
        w0_force_cells(P, w0_old, w0, w0_force);
    }
}

}  // namespace

size_t w0_scratch_doubles(int nr_fine) { return (size_t)14 * (nr_fine + 2); }

void w0_planar(const MxW0Params& P, const double* w0_old, const double* Sbar,
               const double* p0_old, const double* p0_new, const double* g1o, const double* g1n,
               const double* pmp, const double* etarho_cc, double* w0, double* w0_force,
               double* scratch, double* p0bdot_dev, cudaStream_t stream) {
    k_w0_planar<<<1, W0_THREADS, 0, stream>>>(P, w0_old, Sbar, p0_old, p0_new, g1o, g1n, pmp,
                                             etarho_cc, w0, w0_force, scratch, p0bdot_dev);
    mx_count_launch();
}

void w0_sphr(const MxW0Params& P, const double* w0_old, const double* Sbar,
             const double* rho0_old, const double* rho0_new, const double* p0_old,
             const double* p0_new, const double* g1o, const double* g1n, const double* pmp,
             const double* etarho_cc, const double* etarho_ec, const double* rcc,
             const double* re, double* w0, double* w0_force, double* scratch, int* status_dev,
             bool irreg, cudaStream_t stream) {
    if (irreg)
        k_w0_sphr<true><<<1, W0_THREADS, 0, stream>>>(P, w0_old, Sbar, rho0_old, rho0_new, p0_old,
                                                     p0_new, g1o, g1n, pmp, etarho_cc, etarho_ec, rcc,
                                                     re, w0, w0_force, scratch, status_dev);
    else
        k_w0_sphr<false><<<1, W0_THREADS, 0, stream>>>(P, w0_old, Sbar, rho0_old, rho0_new, p0_old,
                                                      p0_new, g1o, g1n, pmp, etarho_cc, etarho_ec, rcc,
                                                      re, w0, w0_force, scratch, status_dev);
    mx_count_launch();
}

// ---- multilevel planar Makew0 (MaestroMakew0.cpp:60-272 with max_radial_level > 0, plus
// RestrictBase / FillGhostBase, MaestroBaseStateGeometry.cpp:236-355) ---------------------------
// The level loop is a chain of short serial recurrences coupled through restriction and offsets;
// one thread walks it (a few thousand steps), the block's other threads only take part in the
// per-cell w0_force loop.
namespace {
struct KRadial {
    int max_lev, finest_lev, nrf, chunk_len;
    const double* dr;
    const int *nr, *cut, *nchunk, *rs, *re;
};
#define RS_(n, j) (G.rs[(n) * G.chunk_len + (j)])
#define RE_(n, j) (G.re[(n) * G.chunk_len + (j)])

__device__ double ml_dmin(double a, double b) { return (b < a) ? b : a; }
__device__ double ml_slope(const double* c, int r) {
    const double del = 0.5 * (c[r + 1] - c[r - 1]);
    const double dpls = 2.0 * (c[r + 1] - c[r]);
    const double dmn = 2.0 * (c[r] - c[r - 1]);
    double slim = ml_dmin(fabs(dpls), fabs(dmn));
    slim = dpls * dmn > 0.0 ? slim : 0.0;
    return copysign(1.0, del) * ml_dmin(slim, fabs(del));
}
__device__ void ml_restrict(const KRadial& G, double* s0, int len, bool cc) {
    for (int n = G.finest_lev; n >= 1; --n)
        for (int i = 1; i <= G.nchunk[n]; ++i) {
            double *f = s0 + (long)n * len, *c = s0 + (long)(n - 1) * len;
            if (cc) {
                for (int j = RS_(n, i); j < RE_(n, i); j += 2) c[j / 2] = 0.5 * (f[j] + f[j + 1]);
            } else {
                for (int j = RS_(n, i); j <= RE_(n, i) + 1; j += 2) c[j / 2] = f[j];
            }
        }
}
__device__ void ml_fill_ghost(const KRadial& G, double* s0, int len, bool cc) {
    for (int n = G.finest_lev; n >= 1; --n)
        for (int i = 1; i <= G.nchunk[n]; ++i) {
            const int lo = RS_(n, i), hi = RE_(n, i);
            double *f = s0 + (long)n * len, *c = s0 + (long)(n - 1) * len;
            if (cc) {
                if (lo != 0) {
                    int rc = lo / 2 - 1;
                    double sl = ml_slope(c, rc);
                    f[lo - 1] = c[rc] + 0.25 * sl;
                    f[lo - 2] = c[rc] - 0.25 * sl;
                    rc = lo / 2 - 2;
                    sl = ml_slope(c, rc);
                    f[lo - 3] = c[rc] + 0.25 * sl;
                    f[lo - 4] = c[rc] - 0.25 * sl;
                }
                if (hi != G.nr[n] - 1) {
                    int rc = (hi + 1) / 2;
                    double sl = ml_slope(c, rc);
                    f[hi + 1] = c[rc] - 0.25 * sl;
                    f[hi + 2] = c[rc] + 0.25 * sl;
                    rc = (hi + 1) / 2 + 1;
                    sl = ml_slope(c, rc);
                    f[hi + 3] = c[rc] - 0.25 * sl;
                    f[hi + 4] = c[rc] + 0.25 * sl;
                }
            } else {
                if (lo != 0) {
                    f[lo - 1] = -f[lo + 1] / 3.0 + f[lo] + c[lo / 2 - 1] / 3.0;
                    f[lo - 2] = c[(lo - 2) / 2];
                }
                if (hi + 1 != G.nr[n]) {
                    f[hi + 2] = -f[hi] / 3.0 + f[hi + 1] + c[(hi + 3) / 2] / 3.0;
                    f[hi + 3] = c[(hi + 3) / 2];
                }
            }
        }
}

__global__ void __launch_bounds__(256, 1)
k_w0_planar_ml(const MxW0Params P, const KRadial G, const double* __restrict__ w0_old,
               const double* __restrict__ Sbar, const double* __restrict__ p0_old,
               const double* __restrict__ p0_new, const double* __restrict__ g1o,
               const double* __restrict__ g1n, const double* __restrict__ pmp,
               const double* __restrict__ etarho_cc, double* w0, double* w0_force, double* scratch,
               double* p0bdot_out) {
    const int nrf = G.nrf, lc = nrf, le = nrf + 1, nlev = G.max_lev + 1;
    double* psi = scratch;
    double* int1 = scratch + nrf + 1;
    for (long q = threadIdx.x; q < (long)nlev * le; q += blockDim.x) w0[q] = 0.0;       // :87
    for (long q = threadIdx.x; q < (long)nlev * lc; q += blockDim.x) w0_force[q] = 0.0;  // Makew0 :18
    __syncthreads();
#define W_(n, r) w0[(long)(n) * le + (r)]
#define C_(a, n, r) a[(long)(n) * lc + (r)]
    if (threadIdx.x == 0) {
        for (int n = 0; n <= G.max_lev; ++n) {
            for (int r = 0; r < nrf; ++r) psi[r] = 0.0;
            const int cut = G.cut[n];
            const double dr = G.dr[n];
            for (int j = 1; j <= G.nchunk[n]; ++j) {
                if (n == 0) {
                    W_(0, 0) = 0.0;
                    int1[0] = 0.0;
                } else {
                    W_(n, RS_(n, j)) = W_(n - 1, RS_(n, j) / 2);
                }
                for (int r = RS_(n, j); r <= RE_(n, j); ++r)
                    if (r < cut) psi[r] = C_(etarho_cc, n, r) * fabs(P.grav_const);
                for (int r = RS_(n, j) + 1; r <= RE_(n, j) + 1; ++r) {
                    const double gavg = (C_(g1o, n, r - 1) + C_(g1n, n, r - 1)) *
                                        (C_(p0_old, n, r - 1) + C_(p0_new, n, r - 1)) / 4.0;
                    double dchi = 0.0;
                    if (r < cut) {
                        if (P.is_predictor)
                            dchi = P.dpdt_factor * C_(pmp, n, r - 1) / (C_(g1o, n, r - 1) * C_(p0_old, n, r - 1) * P.dt);
                        else
                            dchi += P.dpdt_factor * C_(pmp, n, r - 1) / (C_(g1n, n, r - 1) * C_(p0_new, n, r - 1) * P.dt);
                    }
                    if (n == 0) int1[r] = int1[r - 1] + (1.0 / gavg) * dr;
                    W_(n, r) = W_(n, r - 1) + C_(Sbar, n, r - 1) * dr - psi[r - 1] / gavg * dr - dchi * dr;
                }
                if (n == 0 && P.add_pb) {
                    const int k = RE_(n, j) + 1;
                    const double pb = W_(n, k) / int1[k];
                    if (p0bdot_out) *p0bdot_out = pb;
                    for (int r = RS_(n, j) + 1; r <= RE_(n, j) + 1; ++r) W_(n, r) -= pb * int1[r];
                }
                if (n > 0) {
                    const double offset = W_(n, RE_(n, j) + 1) - W_(n - 1, (RE_(n, j) + 1) / 2);
                    for (int i = n - 1; i >= 0; --i) {
                        const int refrat = 1 << (n - i);
                        // QUIRK :196: the restriction writes level n, not level i
                        for (int r = RS_(n, j); r <= RE_(n, j) + 1; ++r)
                            if (r % refrat == 0) W_(n, r / refrat) = W_(n, r);
                        for (int r = (RE_(n, j) + 1) / refrat + 1; r <= G.nr[i]; ++r) W_(i, r) += offset;
                    }
                }
            }
        }
        for (int n = 1; n <= G.max_lev; ++n)
            for (int j = 1; j <= G.nchunk[n]; ++j) {
                const int lo = RE_(n, j) + 2;
                const int hi = j == G.nchunk[n] ? G.nr[n] : RS_(n, j + 1) - 1;
                for (int r = lo; r <= hi; ++r) W_(n, r) = 0.0;
            }
        ml_restrict(G, w0, le, false);
        ml_fill_ghost(G, w0, le, false);
    }
    __syncthreads();
    const double dt_in = P.dt_in, dtold_in = P.dtold_in;
    const double dt_avg = 0.5 * (dt_in + dtold_in);
    for (int n = 0; n <= G.max_lev; ++n)
        for (int j = 1; j <= G.nchunk[n]; ++j)
            for (int r = RS_(n, j) + threadIdx.x; r <= RE_(n, j); r += blockDim.x) {
                const double* wo = w0_old + (long)n * le;
                const double w0_old_cen = 0.5 * (wo[r] + wo[r + 1]);
                const double w0_new_cen = 0.5 * (W_(n, r) + W_(n, r + 1));
                const double w0_avg = 0.5 * (dt_in * w0_old_cen + dtold_in * w0_new_cen) / dt_avg;
                const double div_avg =
                    0.5 * (dt_in * (wo[r + 1] - wo[r]) + dtold_in * (W_(n, r + 1) - W_(n, r))) / dt_avg;
                w0_force[(long)n * lc + r] = (w0_new_cen - w0_old_cen) / dt_avg + w0_avg * div_avg / G.dr[n];
            }
    __syncthreads();
    if (threadIdx.x == 0) {
        ml_restrict(G, w0_force, lc, true);
        ml_fill_ghost(G, w0_force, lc, true);
    }
#undef W_
#undef C_
}
#undef RS_
#undef RE_
}  // namespace

void w0_planar_ml(const MxW0Params& P, int max_lev, int finest_lev, int nrf, int chunk_len,
                  const double* dr, const int* nr, const int* cut, const int* nchunk, const int* rs,
                  const int* re, const double* w0_old, const double* Sbar, const double* p0_old,
                  const double* p0_new, const double* g1o, const double* g1n, const double* pmp,
                  const double* etarho_cc, double* w0, double* w0_force, double* scratch,
                  double* p0bdot_dev, cudaStream_t stream) {
    KRadial G{max_lev, finest_lev, nrf, chunk_len, dr, nr, cut, nchunk, rs, re};
    k_w0_planar_ml<<<1, 256, 0, stream>>>(P, G, w0_old, Sbar, p0_old, p0_new, g1o, g1n, pmp, etarho_cc,
                                         w0, w0_force, scratch, p0bdot_dev);
    mx_count_launch();
}
