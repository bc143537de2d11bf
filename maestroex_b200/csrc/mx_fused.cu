// mx_fused.cu — fused single-pass edge-state (+ flux) kernel for sm_100a: the performance
// path for the 3-D, non-conservative scalar sweeps of MakeEdgeScal (ppm_type 0/1/2, every BC
// code except ext_dir, which MAESTROeX never applies to scalars — Source/MaestroBCFill.cpp:67).
//
// One launch replaces, per component, the reference's ~15 ParallelFor launches and 24
// intermediate MultiFab stores (Source/MaestroMakeEdgeScalars.cpp:187-289): cell data goes
// in, edge states (and, in density / enthalpy mode, fluxes) come out; every intermediate
// (parabola integrals, 1-D states, the six corner-coupled states, transverse differences)
// lives in shared memory or registers.
//
// Decomposition: a thread block owns an (x,y) tile of BX x BY valid cells (GX x GY threads,
// one per cell of the tile grown by one) and a chunk of z planes, and marches upward in z as
// a software pipeline ("2.5-D streaming"): iteration `a` loads plane a+LA of the scalar into
// a shared-memory ring, reconstructs the parabolas of plane a in all three directions (the z
// stencil lives in per-thread registers), forms 1-D states and the z-independent corner
// states of plane a, the z-coupled corner states of plane a-1, and emits the x/y edge states
// of plane a-1 and the z edge state of face a.  Components are looped inside the block so
// that velocity planes and base-state entries are re-read from L1/L2, and rho' is swept
// first so that the species fluxes can use its edge state.
//
// Index-space bookkeeping follows the reference exactly, with the tile playing the role of
// the reference's tile box ("index-space extents"): everything is
// recomputed on the tile's one-cell halo instead of being exchanged between blocks.
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <map>
#include <mutex>
#include <vector>
#include <cstdlib>
#include <cstring>

#include "mx_device.cuh"
#include "mx_fluxdev.cuh"
#include "mx_fused.h"

using namespace mxd;

namespace {

constexpr int MAXC = 40;

struct FusedArgs {
    KGeom g;
    double dt, eps, dt2;
    double c6h[3], c4h[3];  // dt/6/h_d, dt/4/h_d
    double h[3], inv_h[3];
    int inv_exact[3];  // h_d is a power of two: x / h_d == x * inv_h_d exactly
    int slope_order;
    int is_vel;  // velocity sweep: outflow clamp of the normal component at physical faces
    DFab state, force, sedge[3], umac[3], sflux[3], eta;
    DFab r0o[3], r0n[3];
    double* r0e[3];  // spherical: 0.5*(r0mac_old + r0mac_new) laid out like one sedge[d] component
    int ncomp;
    int comp[MAXC];
    int bclo[MAXC][3], bchi[MAXC][3];
    int mode;  // 0: edge states only, 1: density pass (comp[0] = Rho, then species + fluxes),
               // 2: enthalpy pass (one component + (rho h) flux)
    KFlux f;
    int do_eta;
    int zc;        // z planes per chunk
    int zs_lo, zs_hi;  // planes in the short first / last chunk (0: none), see set_zchunks
    int nzc;           // number of z chunks
    int ntx, nty;  // number of tiles
    int bcdir[3];  // some swept component has a non-interior BC in this direction
};

// z range [k0, k1) of chunk c.  Where a swept component has a physical BC on a vertical domain
// face the box touches, the four planes next to that face form their own short chunk, so that
// only those blocks run the BC instantiation of the streaming kernels (which is ~20 % slower)
// and every other chunk of the column runs the BC-free one.
__device__ __forceinline__ void chunk_range(const FusedArgs& A, int c, int& k0, int& k1) {
    const int vlo = A.g.vlo[2], vend = A.g.vhi[2] + 1;
    const int mlo = vlo + A.zs_lo, mhi = vend - A.zs_hi;
    if (A.zs_lo && c == 0) {
        k0 = vlo;
        k1 = mlo;
    } else if (A.zs_hi && c == A.nzc - 1) {
        k0 = mhi;
        k1 = vend;
    } else {
        k0 = mlo + (c - (A.zs_lo ? 1 : 0)) * A.zc;
        k1 = min(k0 + A.zc, mhi);
    }
}

template <int D>
__device__ __forceinline__ double over_h(const FusedArgs& A, double x) {
    return A.inv_exact[D] ? x * A.inv_h[D] : x / A.h[D];
}

// sm, sp -> Ip, Im (ppm) or slope (ppm_type 0; returned in ip)
template <int PPM, bool BC, int D>
__device__ __forceinline__ void reconstruct(const FusedArgs& A, const double* S /*[7]*/, int I,
                                            int bclo, int bchi, double um, double up,
                                            double& ip, double& im) {
    const bool lobc = BC && (bclo == BC_EXT_DIR || bclo == BC_HOEXTRAP);
    const bool hibc = BC && (bchi == BC_EXT_DIR || bchi == BC_HOEXTRAP);
    if (PPM == 0) {
        if (A.slope_order == 0) {
            ip = 0.0;
        } else if (A.slope_order == 2) {
            ip = slope2(S + 1, I, A.g.dlo[D], A.g.dhi[D], lobc, hibc);
        } else {
            ip = slope4(S + 1, I, A.g.dlo[D], A.g.dhi[D], lobc, hibc, D == 0);
        }
        im = 0.0;
    } else {
        double sm, sp;
        if (PPM == 1) {
            ppm1_parabola<BC>(S, I, A.g.dlo[D], A.g.dhi[D], lobc, hibc, sm, sp);
        } else {
            ppm2_parabola<BC>(S, I, A.g.dlo[D], A.g.dhi[D], lobc, hibc, D == 0, sm, sp);
        }
        const double s0 = S[3];
        const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
        const double sig_p = over_h<D>(A, fabs(up) * A.dt);
        const double sig_m = over_h<D>(A, fabs(um) * A.dt);
        ip = ppm_Ip(sm, sp, s0, s6, up, sig_p, A.eps);
        im = ppm_Im(sm, sp, s0, s6, um, sig_m, A.eps);
    }
}

// 1-D left/right states at a face (MaestroMakeEdgeScalars.cpp:697-707): for ppm_type 0 from
// the slopes of the two cells, otherwise Ip of the left cell and Im of the right cell
template <int PPM, int D>
__device__ __forceinline__ void face_states(const FusedArgs& A, double u, double sL, double sR,
                                            double ipL, double ipR, double imR, double& sl,
                                            double& sr) {
    if (PPM == 0) {
        const double q = over_h<D>(A, A.dt * u);
        sl = sL + 0.5 * (1.0 - q) * ipL;
        sr = sR - 0.5 * (1.0 + q) * ipR;
    } else {
        sl = ipL;
        sr = imR;
    }
}

template <int PPM, bool BC, int GX, int GY>
__global__ void __launch_bounds__(GX* GY, 1) k_fused(const __grid_constant__ FusedArgs A) {
    constexpr int LA = (PPM == 2) ? 3 : 2;  // stencil radius = z look-ahead
    constexpr int R = LA + 1;               // ring slots: planes a .. a+LA
    constexpr int PW = GX + 2 * LA, PH = GY + 2 * LA;
    constexpr int NG = GX * GY;
    constexpr int NT = NG;
    constexpr int BX = GX - 2, BY = GY - 2;

    extern __shared__ double sm[];
    double* ring = sm;
    double* UP = ring + R * PW * PH;    // x-face velocities of plane a: GY rows of GX+1
    double* VP = UP + GY * (GX + 1);    // y-face velocities of plane a: GY+1 rows of GX
    double* arr = VP + (GY + 1) * GX;
    double* IPX = arr + 0 * NG;
    double* IPY = arr + 1 * NG;
    double* SX = arr + 2 * NG;
    double* SY = arr + 3 * NG;
    double* T6X = arr + 4 * NG;
    double* T6Y = arr + 5 * NG;
    double* TZ = arr + 6 * NG;
    double* HXY = arr + 7 * NG;
    double* HYX = arr + 8 * NG;
    double* HXZ = arr + 9 * NG;
    double* HYZ = arr + 10 * NG;
    double* QX1 = arr + 11 * NG;
    double* QX2 = arr + 12 * NG;
    double* QY1 = arr + 13 * NG;
    double* QY2 = arr + 14 * NG;
    double* FS = arr + 15 * NG;

    const int tid = threadIdx.x;
    const int tx = tid % GX, ty = tid / GX;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    const KGeom& g = A.g;
    const int x0 = g.vlo[0] + bx * BX, y0 = g.vlo[1] + by * BY;
    const int thx = min(x0 + BX - 1, g.vhi[0]), thy = min(y0 + BY - 1, g.vhi[1]);
    const int i = x0 - 1 + tx, j = y0 - 1 + ty;
    const int k0 = g.vlo[2] + blockIdx.y * A.zc;
    const int k1 = min(k0 + A.zc, g.vhi[2] + 1);  // exclusive
    const bool last = (k1 == g.vhi[2] + 1);
    const int a0 = k0 - 1;

    const bool inG = (i <= thx + 1) && (j <= thy + 1);
    const bool vxi = inG && tx >= 1 && i <= thx;  // column with both x-faces in the predictor range
    const bool vyj = inG && ty >= 1 && j <= thy;
    const bool fx = inG && tx >= 1;  // thread owns the x-face between (i-1,j) and (i,j)
    const bool fy = inG && ty >= 1;
    // faces this block stores: the tile's low faces, plus the high face only on the last tile
    // (the neighbouring tile owns it otherwise; sflux(Rho) is accumulated, so no double owner)
    const bool ownx = fx && vyj && (i <= thx || thx == g.vhi[0]);
    const bool owny = fy && vxi && (j <= thy || thy == g.vhi[1]);
    const int rc = (ty + LA) * PW + (tx + LA);  // own cell inside a ring plane

    const DFab& st = A.state;
    const long s_col = inG ? st.off(i, j, 0) : 0;  // + k*sz gives the cell (k relative to lo)
    const long w_col = inG ? A.umac[2].off(i, j, 0) : 0;
    const long f_col = inG ? A.force.off(i, j, 0) : 0;

    for (int n = 0; n < A.ncomp; ++n) {
        const int comp = A.comp[n];
        const double* __restrict__ sp_ = st.comp(comp);
        const double* __restrict__ fp_ = A.force.comp(comp);
        const int bclx = A.bclo[n][0], bchx = A.bchi[n][0];
        const int bcly = A.bclo[n][1], bchy = A.bchi[n][1];
        const int bclz = A.bclo[n][2], bchz = A.bchi[n][2];
        const bool flux1 = (A.mode == 1 && n > 0);
        const bool flux2 = (A.mode == 2);
        const bool first_spec = (n == 1);

        auto load_ring = [&](int k, int slot) {
            double* dst = ring + slot * PW * PH;
            const long kb = (long)(k - st.lo[2]) * st.sz;
            for (int e = tid; e < PW * PH; e += NT) {
                const int px = e % PW, py = e / PW;
                const int gi = x0 - 1 - LA + px - st.lo[0], gj = y0 - 1 - LA + py - st.lo[1];
                double v = 0.0;
                if (gi >= 0 && gi < st.n[0] && gj >= 0 && gj < st.n[1])
                    v = sp_[kb + (long)gj * st.sy + gi];
                dst[e] = v;
            }
        };

        // ---- prologue ---------------------------------------------------------------
        __syncthreads();  // previous component is done with shared memory
        for (int q = 0; q < LA; ++q) load_ring(a0 + q, q);
        double zw[2 * LA + 1];
#pragma unroll
        for (int m = 0; m < 2 * LA; ++m)
            zw[m + 1] = inG ? sp_[s_col + (long)(a0 - LA + m) * st.sz] : 0.0;
        zw[0] = 0.0;
        double w_m1 = 0.0, w_0 = 0.0;
        double w_p1 = inG ? A.umac[2].p[w_col + (long)a0 * A.umac[2].sz] : 0.0;
        double ipz_prev = 0.0, shz_prev = 0.0, t6x_prev = 0.0, t6y_prev = 0.0;
        double slx_p = 0.0, srx_p = 0.0, sly_p = 0.0, sry_p = 0.0, uf_p = 0.0, vf_p = 0.0;
        double c4x_p = 0.0, c4y_p = 0.0, hzx_prev = 0.0, hzy_prev = 0.0;
        double t4xy_prev = 0.0, t4yx_prev = 0.0, fs_prev = 0.0;

        for (int a = a0; a <= k1; ++a) {
            const int it = a - a0;
            const int slot_a = it % R;
            const double* pa = ring + slot_a * PW * PH;
            // ---- phase 0: loads -----------------------------------------------------
            load_ring(a + LA, (it + LA) % R);
            {
                const DFab& U = A.umac[0];
                for (int e = tid; e < GY * (GX + 1); e += NT) {
                    const int fxx = e % (GX + 1), r = e / (GX + 1);
                    const int gi = x0 - 1 + fxx, gj = y0 - 1 + r;
                    double v = 0.0;
                    if (gi <= thx + 2 && gj <= thy + 1) v = U.p[U.off(gi, gj, a)];
                    UP[e] = v;
                }
                const DFab& V = A.umac[1];
                for (int e = tid; e < (GY + 1) * GX; e += NT) {
                    const int c = e % GX, fyy = e / GX;
                    const int gi = x0 - 1 + c, gj = y0 - 1 + fyy;
                    double v = 0.0;
                    if (gi <= thx + 1 && gj <= thy + 2) v = V.p[V.off(gi, gj, a)];
                    VP[e] = v;
                }
            }
            w_m1 = w_0;
            w_0 = w_p1;
            w_p1 = inG ? A.umac[2].p[w_col + (long)(a + 1) * A.umac[2].sz] : 0.0;
            const double f0 = inG ? fp_[f_col + (long)a * A.force.sz] : 0.0;
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 2 * LA; ++m) zw[m] = zw[m + 1];
            zw[2 * LA] = ring[((it + LA) % R) * PW * PH + rc];
            const double s0 = zw[LA];

            // ---- phase 1: reconstruction in x, y, z for plane a ------------------------
            const double um = UP[ty * (GX + 1) + tx], up = UP[ty * (GX + 1) + tx + 1];
            const double vm = VP[ty * GX + tx], vp = VP[(ty + 1) * GX + tx];
            double ipx = 0.0, imx = 0.0, ipy = 0.0, imy = 0.0, ipz = 0.0, imz = 0.0;
            if (inG) {
                double S[7];
                S[0] = S[6] = 0.0;
#pragma unroll
                for (int q = -LA; q <= LA; ++q) S[q + 3] = pa[rc + q];
                reconstruct<PPM, BC, 0>(A, S, i, bclx, bchx, um, up, ipx, imx);
#pragma unroll
                for (int q = -LA; q <= LA; ++q) S[q + 3] = pa[rc + q * PW];
                reconstruct<PPM, BC, 1>(A, S, j, bcly, bchy, vm, vp, ipy, imy);
#pragma unroll
                for (int q = -LA; q <= LA; ++q) S[q + 3] = zw[q + LA];
                reconstruct<PPM, BC, 2>(A, S, a, bclz, bchz, w_0, w_p1, ipz, imz);
            }
            IPX[tid] = ipx;
            IPY[tid] = ipy;
            __syncthreads();

            // ---- phase 2: 1-D states + first Riemann solve -----------------------------
            double slx = 0.0, srx = 0.0, sly = 0.0, sry = 0.0, slz = 0.0, srz = 0.0, shz = 0.0;
            {
                double shx = 0.0, shy = 0.0;
                if (fx) {
                    face_states<PPM, 0>(A, um, pa[rc - 1], s0, IPX[tid - 1], ipx, imx, slx, srx);
                    if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, slx, srx);
                    shx = upwind(um, 0.0, slx, srx);  // QUIRK: |u| > 0.0 here (:746)
                }
                if (fy) {
                    face_states<PPM, 1>(A, vm, pa[rc - PW], s0, IPY[tid - GX], ipy, imy, sly, sry);
                    if (BC) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, sly, sry);
                    shy = upwind(vm, 0.0, sly, sry);
                }
                if (inG && a >= k0) {
                    face_states<PPM, 2>(A, w_0, zw[LA - 1], s0, ipz_prev, ipz, imz, slz, srz);
                    if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, slz, srz);
                    shz = upwind(w_0, 0.0, slz, srz);
                }
                SX[tid] = shx;
                SY[tid] = shy;
            }
            __syncthreads();

            // ---- phase 3: transverse differences (dt/6 form) ---------------------------
            const double usx = up + um, usy = vp + vm;
            double t6x = 0.0, t6y = 0.0;
            if (vxi) t6x = (A.c6h[0] * usx) * (SX[tid + 1] - SX[tid]);
            if (vyj) t6y = (A.c6h[1] * usy) * (SY[tid + GX] - SY[tid]);
            T6X[tid] = t6x;
            T6Y[tid] = t6y;
            const double wsum_p = w_0 + w_m1;  // faces a and a-1: cell a-1
            if (a >= k0 + 1) TZ[tid] = inG ? (A.c6h[2] * wsum_p) * (shz - shz_prev) : 0.0;
            __syncthreads();

            // ---- phase 4: corner-coupled states ----------------------------------------
            double hzx = 0.0, hzy = 0.0;
            {
                double hxy = 0.0, hyx = 0.0;
                if (fx && vyj) {
                    double l = slx - T6Y[tid - 1], r = srx - T6Y[tid];
                    if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, l, r);
                    hxy = upwind(um, A.eps, l, r);
                }
                if (fy && vxi) {
                    double l = sly - T6X[tid - GX], r = sry - T6X[tid];
                    if (BC) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, l, r);
                    hyx = upwind(vm, A.eps, l, r);
                }
                HXY[tid] = hxy;
                HYX[tid] = hyx;
                if (a >= k0) {
                    if (vxi) {
                        double l = slz - t6x_prev, r = srz - t6x;
                        if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, l, r);
                        hzx = upwind(w_0, A.eps, l, r);
                    }
                    if (vyj) {
                        double l = slz - t6y_prev, r = srz - t6y;
                        if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, l, r);
                        hzy = upwind(w_0, A.eps, l, r);
                    }
                }
                if (a >= k0 + 1) {
                    double hxz = 0.0, hyz = 0.0;
                    if (fx) {
                        double l = slx_p - TZ[tid - 1], r = srx_p - TZ[tid];
                        if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, l, r);
                        hxz = upwind(uf_p, A.eps, l, r);
                    }
                    if (fy) {
                        double l = sly_p - TZ[tid - GX], r = sry_p - TZ[tid];
                        if (BC) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, l, r);
                        hyz = upwind(vf_p, A.eps, l, r);
                    }
                    HXZ[tid] = hxz;
                    HYZ[tid] = hyz;
                }
            }
            __syncthreads();

            // ---- phase 5: transverse differences (dt/4 form) ---------------------------
            double t4xy = 0.0, t4yx = 0.0, t4yz = 0.0, t4zy = 0.0, t4xz = 0.0, t4zx = 0.0;
            if (vxi && vyj) {
                t4xy = (A.c4h[0] * usx) * (HXY[tid + 1] - HXY[tid]);
                t4yx = (A.c4h[1] * usy) * (HYX[tid + GX] - HYX[tid]);
            }
            const double fs_cur = A.dt2 * f0;
            if (a >= k0 + 1) {
                const double c4z = A.c4h[2] * wsum_p;
                if (vyj) {
                    t4yz = c4y_p * (HYZ[tid + GX] - HYZ[tid]);
                    t4zy = c4z * (hzy - hzy_prev);
                }
                if (vxi) {
                    t4xz = c4x_p * (HXZ[tid + 1] - HXZ[tid]);
                    t4zx = c4z * (hzx - hzx_prev);
                }
                QX1[tid] = t4yz;
                QX2[tid] = t4zy;
                QY1[tid] = t4xz;
                QY2[tid] = t4zx;
                FS[tid] = fs_prev;
            }
            __syncthreads();

            // ---- phase 6: final edge states (+ fluxes) ---------------------------------
            if (a >= k0 + 1) {
                const int p = a - 1;
                if (ownx) {
                    const double el = ((slx_p - QX1[tid - 1]) - QX2[tid - 1]) + FS[tid - 1];
                    const double er = ((srx_p - t4yz) - t4zy) + fs_prev;
                    double val = upwind(uf_p, A.eps, el, er);
                    if (BC)
                        val = bc_final(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, el, er,
                                       val);
                    const DFab& E = A.sedge[0];
                    const long eo = E.off(i, j, p);
                    E.comp(comp)[eo] = val;
                    if (flux1) {
                        const DFab& F = A.sflux[0];
                        const long fo = F.off(i, j, p);
                        const double ser = E.comp(A.f.rho)[eo];
                        const double r0e =
                            A.f.spherical ? 0.5 * A.r0o[0].p[A.r0o[0].off(i, j, p)] +
                                                0.5 * A.r0n[0].p[A.r0n[0].off(i, j, p)]
                                          : rho0_edge_planar(A.f, false, p);
                        const double fl = rhoX_flux(A.f.spt, uf_p, r0e, ser, val);
                        F.comp(comp)[fo] = fl;
                        F.comp(A.f.rho)[fo] = (first_spec ? 0.0 : F.comp(A.f.rho)[fo]) + fl;
                    } else if (flux2) {
                        const DFab& F = A.sflux[0];
                        const double ser =
                            needs_sedge_rho_for_rhoh(A.f) ? E.comp(A.f.rho)[eo] : 0.0;
                        double fl;
                        if (rhoh_flux_planar(A.f, false, p, uf_p, ser, val, &fl))
                            F.comp(comp)[F.off(i, j, p)] = fl;
                    }
                }
                if (owny) {
                    const double el = ((sly_p - QY1[tid - GX]) - QY2[tid - GX]) + FS[tid - GX];
                    const double er = ((sry_p - t4xz) - t4zx) + fs_prev;
                    double val = upwind(vf_p, A.eps, el, er);
                    if (BC)
                        val = bc_final(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, el, er,
                                       val);
                    const DFab& E = A.sedge[1];
                    const long eo = E.off(i, j, p);
                    E.comp(comp)[eo] = val;
                    if (flux1) {
                        const DFab& F = A.sflux[1];
                        const long fo = F.off(i, j, p);
                        const double ser = E.comp(A.f.rho)[eo];
                        const double r0e =
                            A.f.spherical ? 0.5 * A.r0o[1].p[A.r0o[1].off(i, j, p)] +
                                                0.5 * A.r0n[1].p[A.r0n[1].off(i, j, p)]
                                          : rho0_edge_planar(A.f, false, p);
                        const double fl = rhoX_flux(A.f.spt, vf_p, r0e, ser, val);
                        F.comp(comp)[fo] = fl;
                        F.comp(A.f.rho)[fo] = (first_spec ? 0.0 : F.comp(A.f.rho)[fo]) + fl;
                    } else if (flux2) {
                        const DFab& F = A.sflux[1];
                        const double ser =
                            needs_sedge_rho_for_rhoh(A.f) ? E.comp(A.f.rho)[eo] : 0.0;
                        double fl;
                        if (rhoh_flux_planar(A.f, false, p, vf_p, ser, val, &fl))
                            F.comp(comp)[F.off(i, j, p)] = fl;
                    }
                }
            }
            if (a >= k0 && (a <= k1 - 1 || last) && vxi && vyj) {
                const double el = ((slz - t4xy_prev) - t4yx_prev) + fs_prev;
                const double er = ((srz - t4xy) - t4yx) + fs_cur;
                double val = upwind(w_0, A.eps, el, er);
                if (BC)
                    val = bc_final(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, el, er, val);
                const DFab& E = A.sedge[2];
                const long eo = E.off(i, j, a);
                E.comp(comp)[eo] = val;
                if (flux1) {
                    const DFab& F = A.sflux[2];
                    const long fo = F.off(i, j, a);
                    const double ser = E.comp(A.f.rho)[eo];
                    const double r0e = A.f.spherical
                                           ? 0.5 * A.r0o[2].p[A.r0o[2].off(i, j, a)] +
                                                 0.5 * A.r0n[2].p[A.r0n[2].off(i, j, a)]
                                           : rho0_edge_planar(A.f, true, a);
                    const double fl = rhoX_flux(A.f.spt, w_0, r0e, ser, val);
                    F.comp(comp)[fo] = fl;
                    F.comp(A.f.rho)[fo] = (first_spec ? 0.0 : F.comp(A.f.rho)[fo]) + fl;
                    if (A.do_eta) {
                        // MaestroMakeFlux.cpp:288-301
                        double* ep = A.eta.p + A.eta.off(i, j, a);
                        double ev = *ep;
                        if (comp >= A.f.first_spec && comp <= A.f.first_spec + A.f.num_spec - 1)
                            ev += fl;
                        if (comp == A.f.first_spec + A.f.num_spec - 1)
                            ev -= A.f.w0[a * A.f.st_w0] *
                                  A.f.rho0_pred_edge[a * A.f.st_rho0_pred_edge];
                        *ep = ev;
                    }
                } else if (flux2) {
                    const DFab& F = A.sflux[2];
                    const double ser = needs_sedge_rho_for_rhoh(A.f) ? E.comp(A.f.rho)[eo] : 0.0;
                    double fl;
                    if (rhoh_flux_planar(A.f, true, a, w_0, ser, val, &fl))
                        F.comp(comp)[F.off(i, j, a)] = fl;
                }
            }

            // ---- rotate the per-thread z history --------------------------------------
            ipz_prev = ipz;
            shz_prev = shz;
            t6x_prev = t6x;
            t6y_prev = t6y;
            slx_p = slx;
            srx_p = srx;
            sly_p = sly;
            sry_p = sry;
            uf_p = um;
            vf_p = vm;
            c4x_p = A.c4h[0] * usx;
            c4y_p = A.c4h[1] * usy;
            hzx_prev = hzx;
            hzy_prev = hzy;
            t4xy_prev = t4xy;
            t4yx_prev = t4yx;
            fs_prev = fs_cur;
        }
    }
}


#include "mx_stream.cuh"
#include "mx_stream4.cuh"
#include "mx_stream5.cuh"

template <int PPM, bool BC, int GX, int GY>
int launch(const FusedArgs& A, cudaStream_t stream) {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    constexpr int R = LA + 1;
    constexpr int PW = GX + 2 * LA, PH = GY + 2 * LA;
    const size_t smem =
        sizeof(double) * (size_t)(R * PW * PH + GY * (GX + 1) + (GY + 1) * GX + 16 * GX * GY);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_fused<PPM, BC, GX, GY>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    const int nz = A.g.vhi[2] - A.g.vlo[2] + 1;
    dim3 grid(A.ntx * A.nty, (nz + A.zc - 1) / A.zc);
    k_fused<PPM, BC, GX, GY><<<grid, GX * GY, smem, stream>>>(A);
    mx_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

constexpr int FGX = 32, FGY = 16;

// v2 (k_stream) needs every per-component plane/column offset to fit a 32-bit int
bool stream_ok(const FusedArgs& A) {
    bool ok = A.state.sn < 0x7fffffffL && A.force.sn < 0x7fffffffL && A.eta.sn < 0x7fffffffL;
    for (int d = 0; d < 3; ++d)
        ok = ok && A.sedge[d].sn < 0x7fffffffL && A.umac[d].sn < 0x7fffffffL &&
             A.sflux[d].sn < 0x7fffffffL;
    return ok;
}

// chunking of the z range (see chunk_range): short boundary chunks only for k_stream4/5.
// Unless MXADV_ZCHUNK fixes it, the chunk length is chosen by simulating the block schedule
// (one resident block per SM): long chunks amortise the 2.5 planes of per-chunk overhead
// (512^3: 118.7 ms at 63 planes vs 123.0 at 32), short ones balance the last wave (256^3: 15.5 ms
// at 43 vs 16.8 at 32); profiles/r1c_experiments.md.
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// makespan (in plane-iterations) of `tiles` x chunks blocks handed out in launch order
// (blockIdx.x fastest) to `sms` SMs, one resident block each; a chunk of len planes costs
// len + 2.5 (two extra planes + prologue, fitted to the 512^3 z-chunk scan)
double chunk_makespan(long tiles, int sms, const std::vector<int>& lens) {
    std::vector<double> free_at(sms, 0.0);
    std::make_heap(free_at.begin(), free_at.end(), std::greater<double>());
    double end = 0.0;
    for (int len : lens) {
        const double d = len + 2.5;
        for (long t = 0; t < tiles; ++t) {
            std::pop_heap(free_at.begin(), free_at.end(), std::greater<double>());
            const double f = free_at.back() + d;
            free_at.back() = f;
            end = f > end ? f : end;
            std::push_heap(free_at.begin(), free_at.end(), std::greater<double>());
        }
    }
    return end;
}

void set_zchunks(FusedArgs& A, bool allow_short) {
    const int nz = A.g.vhi[2] - A.g.vlo[2] + 1;
    A.zs_lo = A.zs_hi = 0;
    if (allow_short && A.bcdir[2] && nz > 8 + A.zc / 2 && !getenv("MXADV_NO_SHORT_CHUNKS")) {
        if (A.g.vlo[2] <= A.g.dlo[2] + 3) A.zs_lo = 4;
        if (A.g.vhi[2] >= A.g.dhi[2] - 3) A.zs_hi = 4;
    }
    const int mid = nz - A.zs_lo - A.zs_hi;
    const int nshort = (A.zs_lo ? 1 : 0) + (A.zs_hi ? 1 : 0);
    if (allow_short && !getenv("MXADV_ZCHUNK") && mid > 24) {
        // pick the chunk length with the smallest simulated makespan; cached per shape
        const long tiles = (long)A.ntx * A.nty;
        static std::map<std::array<long, 4>, int> cache;
        static std::mutex mtx;
        const std::array<long, 4> key = {tiles, mid, A.zs_lo, A.zs_hi};
        std::lock_guard<std::mutex> lock(mtx);
        auto it = cache.find(key);
        if (it == cache.end()) {
            const int sms = sm_count();
            double best = -1.0;
            int best_zc = A.zc;
            for (int n = (mid + ST_MAXZC - 3) / (ST_MAXZC - 2); n <= (mid + 23) / 24; ++n) {
                const int zc = (mid + n - 1) / n;
                if (zc > ST_MAXZC - 2) continue;
                std::vector<int> lens;
                if (A.zs_lo) lens.push_back(A.zs_lo);
                for (int c = 0, left = mid; c < n && left > 0; ++c, left -= zc)
                    lens.push_back(left < zc ? left : zc);
                if (A.zs_hi) lens.push_back(A.zs_hi);
                const double m = chunk_makespan(tiles, sms, lens);
                if (best < 0.0 || m < best) {
                    best = m;
                    best_zc = zc;
                }
            }
            it = cache.emplace(key, best_zc).first;
        }
        A.zc = it->second;
    }
    A.nzc = (mid + A.zc - 1) / A.zc + nshort;
}

int dispatch(FusedArgs& A, int ppm_type, bool bc, cudaStream_t stream) {
    const int nx = A.g.vhi[0] - A.g.vlo[0] + 1, ny = A.g.vhi[1] - A.g.vlo[1] + 1;
    const char* kv = getenv("MXADV_KERNEL");
    const bool v1 = (kv && kv[0] == 'v' && kv[1] == '1') || !stream_ok(A);
    if (!v1) {
        A.ntx = (nx + 29) / 30;
        A.nty = (ny + 13) / 14;
        if (A.zc > ST_MAXZC) A.zc = ST_MAXZC;
        set_zchunks(A, !(kv && kv[0] == 'v' && kv[1] == '3'));
        const bool v3 = kv && kv[0] == 'v' && kv[1] == '3';
        // k_stream5 (two cell rows per thread) is bit-identical but measured 1.2x slower than
        // k_stream4 on B200: opt-in for A/B runs only
        const bool v5 = kv && kv[0] == 'v' && kv[1] == '5';
        if (v5) {
            if (ppm_type == 0) return launch_stream5_mode<0>(A, stream);
            if (ppm_type == 1) return launch_stream5_mode<1>(A, stream);
            return launch_stream5_mode<2>(A, stream);
        }
        if (!v3) {
            A.nty = (ny + (S4_GY - 2) - 1) / (S4_GY - 2);
            if (ppm_type == 0) return launch_stream4_mode<0>(A, stream);
            if (ppm_type == 1) return launch_stream4_mode<1>(A, stream);
            return launch_stream4_mode<2>(A, stream);
        }
        if (ppm_type == 0) return launch_stream_mode<0>(A, stream);
        if (ppm_type == 1) return launch_stream_mode<1>(A, stream);
        return launch_stream_mode<2>(A, stream);
    }
    A.ntx = (nx + (FGX - 2) - 1) / (FGX - 2);
    A.nty = (ny + (FGY - 2) - 1) / (FGY - 2);
    if (ppm_type == 0) return bc ? launch<0, true, FGX, FGY>(A, stream) : launch<0, false, FGX, FGY>(A, stream);
    if (ppm_type == 1) return bc ? launch<1, true, FGX, FGY>(A, stream) : launch<1, false, FGX, FGY>(A, stream);
    return bc ? launch<2, true, FGX, FGY>(A, stream) : launch<2, false, FGX, FGY>(A, stream);
}

bool is_pow2(double h) {
    int e;
    return std::frexp(h, &e) == 0.5;
}

void fill_common(FusedArgs& A, const KGeom& g, const KEdge& e, const DFab& state,
                 const DFab* sedge, const DFab* umac, const DFab& force) {
    memset(&A, 0, sizeof(A));
    A.g = g;
    A.is_vel = e.is_vel;
    A.dt = e.dt;
    A.eps = e.eps;
    A.dt2 = 0.5 * e.dt;
    const double dt4 = 0.25 * e.dt, dt6 = e.dt / 6.0;
    for (int d = 0; d < 3; ++d) {
        A.h[d] = g.dx[d];
        A.inv_h[d] = 1.0 / g.dx[d];
        A.inv_exact[d] = is_pow2(g.dx[d]) ? 1 : 0;
        A.c6h[d] = dt6 / g.dx[d];
        A.c4h[d] = dt4 / g.dx[d];
        A.sedge[d] = sedge[d];
        A.umac[d] = umac[d];
    }
    A.slope_order = e.slope_order;
    A.state = state;
    A.force = force;
    // z-chunk: enough blocks to fill the machine a few times over, small pipeline overhead
    const int nz = g.vhi[2] - g.vlo[2] + 1;
    int zc = 32;
    const char* env = getenv("MXADV_ZCHUNK");
    if (env) zc = atoi(env);
    if (zc < 1) zc = 1;
    A.zc = zc < nz ? zc : nz;
}

void set_bcdir(FusedArgs& A) {
    for (int d = 0; d < 3; ++d) {
        A.bcdir[d] = 0;
        for (int n = 0; n < A.ncomp; ++n)
            if (A.bclo[n][d] != 0 || A.bchi[n][d] != 0) A.bcdir[d] = 1;
    }
}

bool comps_ok(const MxEdgeParams* par, int first_bccomp, int ncomp, bool* any_bc) {
    for (int n = 0; n < ncomp; ++n) {
        const int* bc = par->bc + 6 * (first_bccomp + n);
        for (int q = 0; q < 6; ++q) {
            if (bc[q] == BC_EXT_DIR) return false;
            if (bc[q] != 0) *any_bc = true;
        }
    }
    return true;
}

}  // namespace

bool fused_edge_supported(const KGeom& g, const KEdge& e) {
    if (getenv("MXADV_FORCE_STAGED")) return false;
    // velocity sweeps (is_vel): k_stream4 only (the A/B kernels have no normal-velocity clamp)
    return g.dim == 3 && (!e.is_vel || !getenv("MXADV_KERNEL")) && !e.is_cons && !e.trace;
}

int fused_edge_scal(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, const DFab& state,
                    const DFab* sedge, const DFab* umac, const DFab& force, cudaStream_t stream) {
    bool any_bc = false;
    if (!comps_ok(par, par->start_bccomp, par->num_comp, &any_bc)) return 1;  // ext_dir: staged
    for (int c0 = 0; c0 < par->num_comp; c0 += MAXC) {
        FusedArgs A;
        fill_common(A, g, e0, state, sedge, umac, force);
        A.mode = 0;
        A.ncomp = std::min(MAXC, par->num_comp - c0);
        for (int n = 0; n < A.ncomp; ++n) {
            A.comp[n] = par->start_scomp + c0 + n;
            const int* bc = par->bc + 6 * (par->start_bccomp + c0 + n);
            for (int d = 0; d < 3; ++d) {
                A.bclo[n][d] = bc[d];
                A.bchi[n][d] = bc[3 + d];
            }
        }
        set_bcdir(A);
        if (dispatch(A, e0.ppm_type, any_bc, stream) != 0) return -1;
    }
    return 0;
}

bool fused_density_supported(const KGeom& g, const KEdge& e, const KFlux& f) {
    if (getenv("MXADV_FORCE_STAGED")) return false;
    return g.dim == 3 && !e.is_vel && !e.is_cons && !e.trace && f.num_comp + 1 <= MAXC &&
           (f.spt == MX_PRED_RHOPRIME_AND_X || f.spt == MX_PRED_RHO_AND_X);
}

int fused_density(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, int rho_bccomp,
                  const KFlux& f, const DFab& state, const DFab* sedge, const DFab* sflux,
                  const DFab& etarho, const DFab* umac, const DFab& force, const KMac& mac,
                  double* const* r0e, cudaStream_t stream) {
    bool any_bc = false;
    if (!comps_ok(par, par->start_bccomp, par->num_comp, &any_bc)) return 1;
    if (!comps_ok(par, rho_bccomp, 1, &any_bc)) return 1;
    FusedArgs A;
    fill_common(A, g, e0, state, sedge, umac, force);
    A.mode = 1;
    A.f = f;
    const bool sph = f.spherical != 0;
    A.do_eta = (!sph && f.evolve && !f.exact) ? 1 : 0;
    A.eta = etarho;
    for (int d = 0; d < 3; ++d) {
        A.sflux[d] = sflux[d];
        A.r0o[d] = mac.r0_old[d];
        A.r0n[d] = mac.r0_new[d];
        A.r0e[d] = r0e ? r0e[d] : nullptr;
    }
    A.ncomp = par->num_comp + 1;
    for (int n = 0; n < A.ncomp; ++n) {
        const int bcc = n == 0 ? rho_bccomp : par->start_bccomp + n - 1;
        A.comp[n] = n == 0 ? f.rho : par->start_scomp + n - 1;
        const int* bc = par->bc + 6 * bcc;
        for (int d = 0; d < 3; ++d) {
            A.bclo[n][d] = bc[d];
            A.bchi[n][d] = bc[3 + d];
        }
    }
    set_bcdir(A);
    return dispatch(A, e0.ppm_type, any_bc, stream);
}

bool fused_enthalpy_supported(const KGeom& g, const KEdge& e, const KFlux& f) {
    if (getenv("MXADV_FORCE_STAGED")) return false;
    return g.dim == 3 && !e.is_vel && !e.is_cons && !e.trace && !f.spherical;
}

int fused_enthalpy(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, const KFlux& f,
                   const DFab& state, const DFab* sedge, const DFab* sflux, const DFab* umac,
                   const DFab& force, cudaStream_t stream) {
    bool any_bc = false;
    if (!comps_ok(par, par->start_bccomp, 1, &any_bc)) return 1;
    FusedArgs A;
    fill_common(A, g, e0, state, sedge, umac, force);
    A.mode = 2;
    A.f = f;
    for (int d = 0; d < 3; ++d) A.sflux[d] = sflux[d];
    A.ncomp = 1;
    A.comp[0] = par->start_scomp;
    const int* bc = par->bc + 6 * par->start_bccomp;
    for (int d = 0; d < 3; ++d) {
        A.bclo[0][d] = bc[d];
        A.bchi[0][d] = bc[3 + d];
    }
    set_bcdir(A);
    return dispatch(A, e0.ppm_type, any_bc, stream);
}
