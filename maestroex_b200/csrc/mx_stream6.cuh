// mx_stream6.cuh — k_stream6: the role-specialised streaming edge-state (+flux) kernel.
//
// Same tile decomposition and the same arithmetic as k_stream4 (mx_stream4.cuh): a block owns a
// 32 x 16 tile of cells (30 x 14 valid + one-cell halo) and a chunk of z planes and marches
// upward in z, components looped inside.  What changed is who does what.  The round-1 profiles
// showed two halves with different limits that never overlapped because every warp ran every
// phase in lock step: the reconstruction (fp64-throughput bound, ~75 % fp64 pipe) and the
// Riemann / transverse / flux phases (latency bound, ~20 % fp64 pipe).  Here they are separate
// warp roles with separate register budgets (setmaxnreg):
//
//   * 4 RECONSTRUCTION warps (128 threads, 128 registers).  A thread owns a lane (x) and four
//     consecutive tile rows, one plane AHEAD of the Riemann warps.  It reads the scalar planes
//     from a shared-memory ring filled by TMA (cp.async.bulk.tensor, one elected thread, one
//     mbarrier per slot; fab-edge zero fill by the hardware), builds the limited parabolas
//     (ppm_type 1/2) or slopes (0) in x, y and z and writes, per cell, the SIX one-dimensional
//     states a face can take from it: left state of its high face and right state of its low
//     face per direction (Ip / Im, or s -+ (1 -+ u dt/h) slope / 2).  The y direction shares
//     van Leer slopes / edge values between the four rows of a thread (1.5 + 1.25 instead of
//     3 + 2 per cell), z is incremental as before.
//   * 16 RIEMANN warps (512 threads, 96 registers), one thread per cell as in k_stream4: upwind,
//     dt/6 transverse terms, corner states, dt/4 terms, final edge states, fluxes.  They no
//     longer touch the scalar at all, only the I-buffer, so the kernel body is the same for
//     every ppm_type.
//
// Hand-off: a 3-slot I-buffer in shared memory with FULL / EMPTY named barriers per slot
// (bar.arrive on one side, bar.sync on the other): the reconstruction warps may run up to two
// planes ahead.  Slot s of plane p holds Lx,Rx,Ly,Ry,Rz of plane p and Lz of face p (written
// while plane p-1 was reconstructed); Lx is stored shifted by one lane and Ly by one row so
// that every Riemann thread reads "its face" at its own index, without shuffles or a barrier.
#ifndef MX_STREAM6_CUH
#define MX_STREAM6_CUH

#include <type_traits>

constexpr int S6_RS = 8;          // scalar ring slots
constexpr int S6_IS = 3;          // I-buffer slots
constexpr int S6_IBS = 17 * 32;   // doubles per I-buffer array: 16 tile rows + one slack row
constexpr int S6_US = 16 * 33 + 16;  // x MAC velocity of a plane: 16 rows x (32 lanes + the face beyond)
constexpr int S6_VS = 17 * 32;       // y MAC velocity: 17 face rows
constexpr int S6_SLOT = 6 * S6_IBS + S6_US + S6_VS;  // one hand-off slot
// rows of the tile one reconstruction thread owns (4 -> 4 warps, 2 -> 8 warps) and the
// register budgets of the two roles after setmaxnreg (the launch budget is the largest
// multiple of 8 with (512 + S6_NR) * budget <= 65536; the budgets must add up to at most that)
#ifndef S6_RROWS
#define S6_RROWS 4
#endif
#ifndef S6_QREG
#define S6_QREG 88
#endif
#ifndef S6_RREG
#define S6_RREG 128
#endif
constexpr int S6_NQ = 512;        // Riemann threads
constexpr int S6_NR = 32 * (16 / S6_RROWS);  // reconstruction threads
constexpr int S6_NT = S6_NQ + S6_NR;
constexpr int S6_LREG = (65536 / S6_NT) / 8 * 8;  // registers per thread at launch
static_assert(S6_NQ * S6_QREG + S6_NR * S6_RREG <= S6_NT * S6_LREG, "register budgets exceed the launch allocation");
constexpr int S6_XS = 512 + 64;   // exchange array incl. one slack row either side
// named barriers (0 is __syncthreads)
constexpr int B6_Q1 = 1, B6_Q2 = 2, B6_R = 3, B6_FULL = 4, B6_EMPTY = 7;

template <int PPM>
__host__ __device__ constexpr int s6_pls() {  // doubles per ring slot: a 128-byte multiple (TMA destination)
    constexpr int LA = (PPM == 2) ? 3 : 2;
    return ((32 + 2 * LA) * (16 + 2 * LA) + 15) & ~15;
}
template <int PPM>
constexpr int stream6_smem_doubles() {
    return S6_RS * s6_pls<PPM>() + S6_IS * S6_SLOT + 7 * S6_XS + 8 * S6_NQ +
           5 * (ST_MAXZC + 2) + 2 + S6_RS;
}

MXD void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
MXD void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
MXD unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
MXD void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
MXD void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
                 : "memory");
}
MXD void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MX_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MX_DONE_%=;\n\t"
        "bra MX_WAIT_%=;\n\t"
        "MX_DONE_%=:\n\t}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// one box {PW, PH, 1, 1} of the state fab (x, y, z, component) -> shared memory
MXD void tma_load_plane(double* dst, const void* tmap, unsigned long long* bar, int c0, int c1,
                        int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct S6Tile {
    int x0, y0, thx, thy, k0, k1;
    bool last;
};
MXD S6Tile s6_tile(const FusedArgs& A) {
    const KGeom& g = A.g;
    S6Tile t;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    t.x0 = g.vlo[0] + bx * 30;
    t.y0 = g.vlo[1] + by * 14;
    t.thx = min(t.x0 + 29, g.vhi[0]);
    t.thy = min(t.y0 + 13, g.vhi[1]);
    chunk_range(A, blockIdx.y, t.k0, t.k1);
    t.last = (t.k1 == g.vhi[2] + 1) && !g.skip_top_z;
    return t;
}

// =======================================================================================
// Reconstruction role
// =======================================================================================
template <int PPM, bool BC>
__device__ __forceinline__ void stream6_recon(const FusedArgs& A, double* smem) {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    constexpr int PW = 32 + 2 * LA, PLS = s6_pls<PPM>();
    constexpr int PF = 7 - 2 * LA;  // planes in flight beyond the resident window
    constexpr int NRW = S6_RROWS;        // tile rows per thread
    constexpr int NY = NRW + 2 * LA;  // column values a thread needs for its rows
    double* const ring = smem;
    double* const ibuf = ring + S6_RS * PLS;
    unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(
        ibuf + S6_IS * S6_SLOT + 7 * S6_XS + 8 * S6_NQ + 5 * (ST_MAXZC + 2) + 2);

    const int rt = threadIdx.x - S6_NQ;
    const int tx = rt & 31, rg = rt >> 5;
    const KGeom& g = A.g;
    const S6Tile T = s6_tile(A);
    const int i = T.x0 - 1 + tx, j0 = T.y0 - 1 + NRW * rg;
    const int a0 = T.k0 - 1, nit = T.k1 - a0;  // steps 0 .. nit
    const int il = min(i, T.thx + 1);
    const int jb = min(j0, T.thy + 1), jvb = min(j0, T.thy + 2);
    const int dm = max(T.thy + 1 - j0, 0), dmv = max(T.thy + 2 - j0, 0);  // row clamps
    const long u_sz = A.umac[0].sz, v_sz = A.umac[1].sz, w_sz = A.umac[2].sz;
    const int u_sy = (int)A.umac[0].sy, v_sy = (int)A.umac[1].sy, w_sy = (int)A.umac[2].sy;
    const double eps = A.eps, dt = A.dt;
    const bool elected = (rt == 0);
    const int cx0 = T.x0 - 1 - LA - A.state.lo[0], cy0 = T.y0 - 1 - LA - A.state.lo[1];
    const int cz0 = a0 - LA - A.state.lo[2];
    const int rc0 = (NRW * rg + LA) * PW + tx + LA;  // my row 0, my lane, in a ring plane
    unsigned phase = 0;                            // per-slot parity of the next wait

    const bool nearx = BC && __any_sync(0xffffffffu, (i <= g.dlo[0] + 4) || (i >= g.dhi[0] - 4));

    for (int n = 0; n < A.ncomp; ++n) {
        const int comp = A.comp[n];
        const int bclx = A.bclo[n][0], bchx = A.bchi[n][0];
        const int bcly = A.bclo[n][1], bchy = A.bchi[n][1];
        const int bclz = A.bclo[n][2], bchz = A.bchi[n][2];
        const bool lobcx = nearx && (bclx == BC_EXT_DIR || bclx == BC_HOEXTRAP);
        const bool hibcx = nearx && (bchx == BC_EXT_DIR || bchx == BC_HOEXTRAP);
        const bool lobcy = BC && (bcly == BC_EXT_DIR || bcly == BC_HOEXTRAP);
        const bool hibcy = BC && (bchy == BC_EXT_DIR || bchy == BC_HOEXTRAP);
        const bool lobcz_c = BC && (bclz == BC_EXT_DIR || bclz == BC_HOEXTRAP);
        const bool hibcz_c = BC && (bchz == BC_EXT_DIR || bchz == BC_HOEXTRAP);

        const double* __restrict__ pU = A.umac[0].p + A.umac[0].off(il, jb, a0);
        const double* __restrict__ pV = A.umac[1].p + A.umac[1].off(il, jvb, a0);
        const double* __restrict__ pW = A.umac[2].p + A.umac[2].off(il, jb, a0);  // face a0

        __syncthreads();  // every role is done with the previous component
        // ---- prologue: planes a0-LA .. a0+LA+PF-1 -------------------------------------------
        if (elected) {
            for (int q = 0; q < 2 * LA + PF && q <= nit + 2 * LA; ++q) {
                mbar_expect_tx(mbar + (q & (S6_RS - 1)), PW * (16 + 2 * LA) * 8);
                tma_load_plane(ring + (q & (S6_RS - 1)) * PLS, &A.tmap_state, mbar + (q & (S6_RS - 1)),
                               cx0, cy0, cz0 + q, comp);
            }
        }
        double w_lo[NRW], sg_lo[NRW];
        double zA[NRW], zB[NRW], zC[NRW];  // PPM 1: dz_c, ez_lo;  PPM 2: ez_l2, ez_l1, ez_h1
#pragma unroll
        for (int q = 0; q < NRW; ++q) {
            w_lo[q] = pW[min(q, dm) * w_sy];
            sg_lo[q] = (PPM != 0) ? div_h<2>(A, fabs(w_lo[q]) * dt) : 0.0;
            zA[q] = zB[q] = zC[q] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < 2 * LA; ++q) {  // planes a0-LA .. a0+LA-1 (a0+LA is waited in step 0)
            mbar_wait(mbar + q, (phase >> q) & 1u);
            phase ^= 1u << q;
        }
        if (PPM != 0) {
#pragma unroll
            for (int q = 0; q < NRW; ++q) {
                const double* c = ring + rc0 + q * PW;
                if (PPM == 1) {
                    const double zm2 = c[(LA - 2) * PLS], zm1 = c[(LA - 1) * PLS], z0 = c[LA * PLS],
                                 zp1 = c[(LA + 1) * PLS];
                    const double dm1 = vanleer(zm2, zm1, z0);
                    zA[q] = vanleer(zm1, z0, zp1);
                    zB[q] = vl_edge(zm1, z0, dm1, zA[q]);
                } else {
                    const double z[6] = {c[0],       c[PLS],     c[2 * PLS],
                                         c[3 * PLS], c[4 * PLS], c[5 * PLS]};
                    zA[q] = cs_edge(z[0], z[1], z[2], z[3]);
                    zB[q] = cs_edge(z[1], z[2], z[3], z[4]);
                    zC[q] = cs_edge(z[2], z[3], z[4], z[5]);
                }
            }
        }

        for (int it = 0; it <= nit; ++it) {
            const int a = a0 + it;
            const bool bcz = BC && (bclz != 0 || bchz != 0) &&
                             ((a <= g.dlo[2] + 4) || (a >= g.dhi[2] - 4));
            const bool lobcz = lobcz_c && bcz, hibcz = hibcz_c && bcz;
            // ---- ring bookkeeping ----------------------------------------------------------
            bar_sync(B6_R, S6_NR);  // every reconstruction thread has finished step it-1
            if (elected) {
                const int q = it + 2 * LA + PF;  // plane a + LA + PF
                if (q <= nit + 2 * LA) {
                    mbar_expect_tx(mbar + (q & (S6_RS - 1)), PW * (16 + 2 * LA) * 8);
                    tma_load_plane(ring + (q & (S6_RS - 1)) * PLS, &A.tmap_state,
                                   mbar + (q & (S6_RS - 1)), cx0, cy0, cz0 + q, comp);
                }
            }
            // velocities of this plane (issued before the waits: they are not needed until the
            // parabolas are limited)
            double um[NRW], vm[NRW + 1], w_hi[NRW], up31[NRW];
            pW += w_sz;
            const bool more = it < nit;  // next step's lines start towards L2 while this one computes
#pragma unroll
            for (int q = 0; q < NRW; ++q) {
                const double* au = pU + min(q, dm) * u_sy;
                const double* aw = pW + min(q, dm) * w_sy;
                um[q] = au[0];
                w_hi[q] = aw[0];
                up31[q] = (tx == 31) ? au[1] : 0.0;
                if (more) {
                    prefetch_l2(au + u_sz);
                    prefetch_l2(aw + w_sz);
                }
            }
#pragma unroll
            for (int q = 0; q < NRW + 1; ++q) {
                const double* av = pV + min(q, dmv) * v_sy;
                vm[q] = av[0];
                if (more && q < NRW) prefetch_l2(av + v_sz);
            }
            pU += u_sz;
            pV += v_sz;
            {
                const int s = (it + 2 * LA) & (S6_RS - 1);  // plane a + LA
                mbar_wait(mbar + s, (phase >> s) & 1u);
                phase ^= 1u << s;
            }
            if (it >= 2) bar_sync(B6_EMPTY + it % 3, S6_NT);  // Riemann warps released this slot
            double* const I = ibuf + (it % 3) * S6_SLOT + (NRW * rg) * 32 + tx;
            double* const In = ibuf + ((it + 1) % 3) * S6_SLOT + (NRW * rg) * 32 + tx;
            const double* const P = ring + ((it + LA) & (S6_RS - 1)) * PLS + rc0;  // plane a, row 0

            // ---- x --------------------------------------------------------------------------
            // ppm_type 1: two rows at a time (D2), so that this lone warp of the sub-partition
            // always has an independent instruction next to the one it waits for
            constexpr bool PAIR = (PPM == 1) && (NRW % 2 == 0);
            if (PAIR) {
#pragma unroll
                for (int q = 0; q < NRW; q += 2) {
                    D2 S[7];
                    S[0] = S[6] = d2(0.0);
#pragma unroll
                    for (int o = -LA; o <= LA; ++o) S[o + 3] = d2(P[q * PW + o], P[(q + 1) * PW + o]);
                    D2 sm, sp;
                    ppm1_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, sm, sp);
                    const D2 umv = d2(um[q], um[q + 1]);
                    const D2 sg_m = div_h<0>(A, abs_(umv) * dt);
                    const D2 sg_p = shfl_dn1(sg_m);
                    const D2 up = shfl_dn1(umv);
                    const D2 s6 = 6.0 * S[3] - 3.0 * (sm + sp);
                    const D2 lx = ppm_Ip(sm, sp, S[3], s6, up, sg_p, eps);
                    const D2 rx = ppm_Im(sm, sp, S[3], s6, umv, sg_m, eps);
                    if (tx < 31) {
                        I[0 * S6_IBS + q * 32 + 1] = lx.a;
                        I[0 * S6_IBS + (q + 1) * 32 + 1] = lx.b;
                    }
                    I[1 * S6_IBS + q * 32] = rx.a;
                    I[1 * S6_IBS + (q + 1) * 32] = rx.b;
                }
            } else {
#pragma unroll
            for (int q = 0; q < NRW; ++q) {
                double S[7];
                S[0] = S[6] = 0.0;
#pragma unroll
                for (int o = -LA; o <= LA; ++o) S[o + 3] = P[q * PW + o];
                const double s0 = S[3];
                double lx, rx;
                if (PPM == 0) {
                    const double sl = slope_any(A.slope_order, S + 1, i, g.dlo[0], g.dhi[0], lobcx, hibcx, true);
                    const double qm = div_h<0>(A, dt * um[q]);
                    const double qp = shfl_dn1(qm);
                    lx = s0 + 0.5 * (1.0 - qp) * sl;
                    rx = s0 - 0.5 * (1.0 + qm) * sl;
                } else {
                    double sm, sp;
                    if (PPM == 1)
                        ppm1_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, sm, sp);
                    else
                        ppm2_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, true, sm, sp);
                    const double sg_m = div_h<0>(A, fabs(um[q]) * dt);
                    const double sg_p = shfl_dn1(sg_m);
                    const double up = shfl_dn1(um[q]);
                    const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
                    lx = ppm_Ip(sm, sp, s0, s6, up, sg_p, eps);
                    rx = ppm_Im(sm, sp, s0, s6, um[q], sg_m, eps);
                }
                if (tx < 31) I[0 * S6_IBS + q * 32 + 1] = lx;
                I[1 * S6_IBS + q * 32] = rx;
            }
            }
            // ---- y (slopes / edge values shared between my four rows) ------------------------
            {
                double Y[NY];
#pragma unroll
                for (int t = 0; t < NY; ++t) Y[t] = P[(t - LA) * PW];
                double sgy[NRW + 1];
                if (PPM != 0) {
                    if (PAIR) {
#pragma unroll
                        for (int q = 0; q + 1 < NRW + 1; q += 2) {
                            const D2 s = div_h<1>(A, abs_(d2(vm[q], vm[q + 1])) * dt);
                            sgy[q] = s.a;
                            sgy[q + 1] = s.b;
                        }
                        sgy[NRW] = div_h<1>(A, fabs(vm[NRW]) * dt);
                    } else {
#pragma unroll
                        for (int q = 0; q < NRW + 1; ++q) sgy[q] = div_h<1>(A, fabs(vm[q]) * dt);
                    }
                }
                double G[NRW + 3];  // PPM 1: G[0..NRW] edges below rows 0..NRW;  PPM 2: edges -1 .. NRW+1
                if (PPM == 1) {
                    double d[NRW + 2];
                    if (PAIR) {
#pragma unroll
                        for (int t = 0; t < NRW + 2; t += 2) {
                            const D2 v = vanleer(d2(Y[t], Y[t + 1]), d2(Y[t + 1], Y[t + 2]), d2(Y[t + 2], Y[t + 3]));
                            d[t] = v.a;
                            d[t + 1] = v.b;
                        }
#pragma unroll
                        for (int t = 0; t + 1 < NRW + 1; t += 2) {
                            const D2 v = vl_edge(d2(Y[t + 1], Y[t + 2]), d2(Y[t + 2], Y[t + 3]), d2(d[t], d[t + 1]),
                                                 d2(d[t + 1], d[t + 2]));
                            G[t] = v.a;
                            G[t + 1] = v.b;
                        }
                        G[NRW] = vl_edge(Y[NRW + 1], Y[NRW + 2], d[NRW], d[NRW + 1]);
                    } else {
#pragma unroll
                        for (int t = 0; t < NRW + 2; ++t) d[t] = vanleer(Y[t], Y[t + 1], Y[t + 2]);
#pragma unroll
                        for (int t = 0; t < NRW + 1; ++t) G[t] = vl_edge(Y[t + 1], Y[t + 2], d[t], d[t + 1]);
                    }
                } else if (PPM == 2) {
#pragma unroll
                    for (int t = 0; t < NRW + 3; ++t) G[t] = cs_edge(Y[t], Y[t + 1], Y[t + 2], Y[t + 3]);
                }
                if (PAIR && !BC) {
                    // away from physical boundaries the limiter does not depend on the row index
#pragma unroll
                    for (int q = 0; q < NRW; q += 2) {
                        D2 S[7];
                        S[0] = S[6] = d2(0.0);
#pragma unroll
                        for (int o = -LA; o <= LA; ++o) S[o + 3] = d2(Y[q + LA + o], Y[q + 1 + LA + o]);
                        D2 sm, sp;
                        ppm1_finish<false>(S, 0, 0, 0, false, false, d2(G[q], G[q + 1]), d2(G[q + 1], G[q + 2]), sm, sp);
                        const D2 s6 = 6.0 * S[3] - 3.0 * (sm + sp);
                        const D2 ly = ppm_Ip(sm, sp, S[3], s6, d2(vm[q + 1], vm[q + 2]), d2(sgy[q + 1], sgy[q + 2]), eps);
                        const D2 ry = ppm_Im(sm, sp, S[3], s6, d2(vm[q], vm[q + 1]), d2(sgy[q], sgy[q + 1]), eps);
                        I[2 * S6_IBS + (q + 1) * 32] = ly.a;
                        I[2 * S6_IBS + (q + 2) * 32] = ly.b;
                        I[3 * S6_IBS + q * 32] = ry.a;
                        I[3 * S6_IBS + (q + 1) * 32] = ry.b;
                    }
                } else {
#pragma unroll
                for (int q = 0; q < NRW; ++q) {
                    const int j = j0 + q;
                    const bool ny = BC && ((j <= g.dlo[1] + 4) || (j >= g.dhi[1] - 4));
                    const bool lo_ = lobcy && ny, hi_ = hibcy && ny;
                    double S[7];
                    S[0] = S[6] = 0.0;
#pragma unroll
                    for (int o = -LA; o <= LA; ++o) S[o + 3] = Y[q + LA + o];
                    const double s0 = S[3];
                    double ly, ry;
                    if (PPM == 0) {
                        const double sl = slope_any(A.slope_order, S + 1, j, g.dlo[1], g.dhi[1], lo_, hi_, false);
                        const double qm = div_h<1>(A, dt * vm[q]), qp = div_h<1>(A, dt * vm[q + 1]);
                        ly = s0 + 0.5 * (1.0 - qp) * sl;
                        ry = s0 - 0.5 * (1.0 + qm) * sl;
                    } else {
                        double sm, sp;
                        if (PPM == 1)
                            ppm1_finish<BC>(S, j, g.dlo[1], g.dhi[1], lo_, hi_, G[q], G[q + 1], sm, sp);
                        else
                            ppm2_finish<BC>(S, j, g.dlo[1], g.dhi[1], lo_, hi_, false, G[q], G[q + 1],
                                            G[q + 2], G[q + 3], sm, sp);
                        const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
                        ly = ppm_Ip(sm, sp, s0, s6, vm[q + 1], sgy[q + 1], eps);
                        ry = ppm_Im(sm, sp, s0, s6, vm[q], sgy[q], eps);
                    }
                    I[2 * S6_IBS + (q + 1) * 32] = ly;  // row q+1 reads it as its low-face left state
                    I[3 * S6_IBS + q * 32] = ry;
                }
                }
            }
            // ---- z (incremental) -------------------------------------------------------------
            if (PAIR) {
#pragma unroll
                for (int q = 0; q < NRW; q += 2) {
                    D2 S[7];
                    S[0] = S[6] = d2(0.0);
#pragma unroll
                    for (int o = -LA; o <= LA; ++o) {
                        const bool need = (BC && bcz) || o >= 0;
                        const double* c = ring + ((it + LA + o) & (S6_RS - 1)) * PLS + rc0 + q * PW;
                        S[o + 3] = need ? d2(c[0], c[PW]) : d2(0.0);
                    }
                    const D2 dz_n = vanleer(S[3], S[4], S[5]);
                    const D2 ez_hi = vl_edge(S[3], S[4], d2(zA[q], zA[q + 1]), dz_n);
                    D2 sm, sp;
                    ppm1_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, d2(zB[q], zB[q + 1]), ez_hi, sm, sp);
                    zA[q] = dz_n.a;
                    zA[q + 1] = dz_n.b;
                    zB[q] = ez_hi.a;
                    zB[q + 1] = ez_hi.b;
                    const D2 s6 = 6.0 * S[3] - 3.0 * (sm + sp);
                    const D2 whi = d2(w_hi[q], w_hi[q + 1]);
                    const D2 sg_hi = div_h<2>(A, abs_(whi) * dt);
                    const D2 lz = ppm_Ip(sm, sp, S[3], s6, whi, sg_hi, eps);
                    const D2 rz = ppm_Im(sm, sp, S[3], s6, d2(w_lo[q], w_lo[q + 1]), d2(sg_lo[q], sg_lo[q + 1]), eps);
                    sg_lo[q] = sg_hi.a;
                    sg_lo[q + 1] = sg_hi.b;
                    w_lo[q] = w_hi[q];
                    w_lo[q + 1] = w_hi[q + 1];
                    In[4 * S6_IBS + q * 32] = lz.a;  // left state of face a+1
                    In[4 * S6_IBS + (q + 1) * 32] = lz.b;
                    I[5 * S6_IBS + q * 32] = rz.a;
                    I[5 * S6_IBS + (q + 1) * 32] = rz.b;
                }
            } else {
#pragma unroll
            for (int q = 0; q < NRW; ++q) {
                double S[7];
                S[0] = S[6] = 0.0;
#pragma unroll
                for (int o = -LA; o <= LA; ++o) {
                    const bool need = (BC && bcz) || PPM != 1 || o >= 0;
                    S[o + 3] = need ? ring[((it + LA + o) & (S6_RS - 1)) * PLS + rc0 + q * PW] : 0.0;
                }
                const double s0 = S[3];
                double lz, rz;
                if (PPM == 0) {
                    const double sl = slope_any(A.slope_order, S + 1, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false);
                    const double qm = div_h<2>(A, dt * w_lo[q]), qp = div_h<2>(A, dt * w_hi[q]);
                    lz = s0 + 0.5 * (1.0 - qp) * sl;
                    rz = s0 - 0.5 * (1.0 + qm) * sl;
                } else {
                    double sm, sp;
                    if (PPM == 1) {
                        const double dz_n = vanleer(S[3], S[4], S[5]);
                        const double ez_hi = vl_edge(S[3], S[4], zA[q], dz_n);
                        ppm1_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, zB[q], ez_hi, sm, sp);
                        zA[q] = dz_n;
                        zB[q] = ez_hi;
                    } else {
                        const double ez_h2 = cs_edge(S[3], S[4], S[5], S[6]);
                        ppm2_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false, zA[q], zB[q], zC[q],
                                        ez_h2, sm, sp);
                        zA[q] = zB[q];
                        zB[q] = zC[q];
                        zC[q] = ez_h2;
                    }
                    const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
                    const double sg_hi = div_h<2>(A, fabs(w_hi[q]) * dt);
                    lz = ppm_Ip(sm, sp, s0, s6, w_hi[q], sg_hi, eps);
                    rz = ppm_Im(sm, sp, s0, s6, w_lo[q], sg_lo[q], eps);
                    sg_lo[q] = sg_hi;
                }
                w_lo[q] = w_hi[q];
                In[4 * S6_IBS + q * 32] = lz;  // left state of face a+1
                I[5 * S6_IBS + q * 32] = rz;
            }
            }
            {   // the MAC velocities of this plane for the Riemann warps (they keep no pointers or
                // prefetch registers for them): U[row][lane] plus the face beyond lane 31, V[row].
                // Stored at the END of the step: the loads were issued at its top, so nothing
                // here waits on global memory.
                double* const U = ibuf + (it % 3) * S6_SLOT + 6 * S6_IBS + (NRW * rg) * 33 + tx;
                double* const V = ibuf + (it % 3) * S6_SLOT + 6 * S6_IBS + S6_US + (NRW * rg) * 32 + tx;
#pragma unroll
                for (int q = 0; q < NRW; ++q) {
                    U[q * 33] = um[q];
                    V[q * 32] = vm[q];
                    if (tx == 31) U[q * 33 + 1] = up31[q];
                }
                if (rg == 16 / NRW - 1) V[NRW * 32] = vm[NRW];
            }
            bar_arrive(B6_FULL + it % 3, S6_NT);
        }
    }
}

// =======================================================================================
// Riemann / transverse / flux role (one thread per tile cell).  Phases P2-P6 of k_stream4.
// =======================================================================================
template <bool BC, int MODE, bool SPH>
__device__ __forceinline__ void stream6_riemann(const FusedArgs& A, double* smem, int ring_doubles) {
    constexpr int GX = 32, NT = S6_NQ;
    double* const ibuf = smem + ring_doubles;
    double* const ex = ibuf + S6_IS * S6_SLOT + GX;
    double* X_T6X = ex + 0 * S6_XS;
    double* X_TZ = ex + 1 * S6_XS;
    double* X_HYX = ex + 2 * S6_XS;
    double* X_HYZ = ex + 3 * S6_XS;
    double* X_QY1 = ex + 4 * S6_XS;
    double* X_QY2 = ex + 5 * S6_XS;
    double* X_FS = ex + 6 * S6_XS;
    double* hist = ex + 7 * S6_XS - GX;
    double* tab = hist + 8 * NT;
    double* T_RXY = tab;
    double* T_RZ = tab + (ST_MAXZC + 2);
    double* T_W0R = tab + 2 * (ST_MAXZC + 2);
    double* T_HXY = tab + 3 * (ST_MAXZC + 2);
    double* T_HZ = tab + 4 * (ST_MAXZC + 2);

    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    double* const hs = hist + tid;
#define HS(v) hs[(v) * NT]
    enum { H_HZX, H_HZY, H_T4XY, H_T4YX, H_FS, H_T6X, H_T6Y, H_SHZ };

    const KGeom& g = A.g;
    const S6Tile T = s6_tile(A);
    const int x0 = T.x0, y0 = T.y0, thx = T.thx, thy = T.thy, k0 = T.k0, k1 = T.k1;
    const bool last = T.last;
    const int i = x0 - 1 + tx, j = y0 - 1 + ty;
    const int a0 = k0 - 1;

    bool ownx, owny, ownz;
    {
        const bool inG = (i <= thx + 1) && (j <= thy + 1);
        const bool vxi = inG && tx >= 1 && i <= thx;
        const bool vyj = inG && ty >= 1 && j <= thy;
        ownx = inG && tx >= 1 && vyj && (i <= thx || thx == g.vhi[0]);
        owny = inG && ty >= 1 && vxi && (j <= thy || thy == g.vhi[1]);
        ownz = vxi && vyj;
    }
    const int il = min(i, thx + 1), jl = min(j, thy + 1);
    const int icx = min(max(i, x0), thx + 1), jcx = min(max(j, y0), thy);
    const int icy = min(max(i, x0), thx), jcy = min(max(j, y0), thy + 1);
    const int icz = min(max(i, x0), thx), jcz = min(max(j, y0), thy);
    const bool nearx = BC && __any_sync(0xffffffffu, (i <= g.dlo[0] + 4) || (i >= g.dhi[0] - 4));
    const bool neary = BC && ((j <= g.dlo[1] + 4) || (j >= g.dhi[1] - 4));

    const long f_sz = A.force.sz;
    const long w_sz = A.umac[2].sz;
    const long ex_sz = A.sedge[0].sz, ey_sz = A.sedge[1].sz, ez_sz = A.sedge[2].sz;
    const long fx_sz = A.sflux[0].sz, fy_sz = A.sflux[1].sz, fz_sz = A.sflux[2].sz;
    const long eta_sz = A.eta.sz;
    const KFlux& F_ = A.f;
    constexpr bool sph = SPH;

    const double c6x = A.c6h[0], c6y = A.c6h[1], c6z = A.c6h[2];
    const double c4x = A.c4h[0], c4y = A.c4h[1], c4z = A.c4h[2];
    const double eps = A.eps;
    const bool have_h = (MODE == 2) && (F_.ept == MX_PRED_H || F_.ept == MX_PRED_T_THEN_H ||
                                        F_.ept == MX_PRED_TPRIME_THEN_H);
    // sedge(Rho) is an input of the flux: species forms except (rho X); h forms; spherical h'
    const bool need_ser = (MODE == 1) ? (F_.spt != MX_PRED_RHOX)
                                      : (have_h || (sph && F_.ept == MX_PRED_HPRIME));

    const long w_c = A.umac[2].off(il, jl, a0);
    const long f_c = A.force.off(il, jl, a0);
    const long ex_c = A.sedge[0].off(icx, jcx, a0 - 1), ey_c = A.sedge[1].off(icy, jcy, a0 - 1);
    const long ez_c = A.sedge[2].off(icz, jcz, a0);
    long fx_c = 0, fy_c = 0, fz_c = 0, eta_c = 0;
    if (MODE != 0) {
        fx_c = A.sflux[0].off(icx, jcx, a0 - 1);
        fy_c = A.sflux[1].off(icy, jcy, a0 - 1);
        fz_c = A.sflux[2].off(icz, jcz, a0);
        if (MODE == 1 && A.do_eta) eta_c = A.eta.off(icz, jcz, a0);
    }

    for (int n = 0; n < A.ncomp; ++n) {
        const int comp = A.comp[n];
        const int bclx = A.bclo[n][0], bchx = A.bchi[n][0];
        const int bcly = A.bclo[n][1], bchy = A.bchi[n][1];
        const int bclz = A.bclo[n][2], bchz = A.bchi[n][2];
        const bool bcx = nearx && (bclx != 0 || bchx != 0);
        const bool bcy = neary && (bcly != 0 || bchy != 0);
        const bool bcz_c = BC && (bclz != 0 || bchz != 0);
        const bool vcx = BC && A.is_vel && comp == 0, vcy = BC && A.is_vel && comp == 1,
                   vcz = BC && A.is_vel && comp == 2;
        const bool flux1 = (MODE == 1 && n > 0);
        const bool first_spec = (n == 1);
        const bool is_spec = comp >= F_.first_spec && comp <= F_.first_spec + F_.num_spec - 1;
        const bool last_spec = comp == F_.first_spec + F_.num_spec - 1;

        const double* __restrict__ pW = A.umac[2].p + w_c + w_sz;
        const double* __restrict__ pF = A.force.comp(comp) + f_c;
        double* __restrict__ pEx = A.sedge[0].comp(comp) + ex_c;
        double* __restrict__ pEy = A.sedge[1].comp(comp) + ey_c;
        double* __restrict__ pEz = A.sedge[2].comp(comp) + ez_c;
        double* __restrict__ pFx = nullptr;
        double* __restrict__ pFy = nullptr;
        double* __restrict__ pFz = nullptr;
        double* __restrict__ pEta = nullptr;
        long dRx = 0, dRy = 0, dRz = 0, dGx = 0, dGy = 0, dGz = 0;
        long dQx = 0, dQy = 0, dQz = 0;
        long dHx = 0, dHy = 0, dHz = 0;  // spherical enthalpy pass: h0e - this comp of sedge
        if (MODE != 0) {
            pFx = A.sflux[0].comp(comp) + fx_c;
            pFy = A.sflux[1].comp(comp) + fy_c;
            pFz = A.sflux[2].comp(comp) + fz_c;
            dRx = (long)(F_.rho - comp) * A.sedge[0].sn;
            dRy = (long)(F_.rho - comp) * A.sedge[1].sn;
            dRz = (long)(F_.rho - comp) * A.sedge[2].sn;
            if (MODE == 1) {
                if (sph) {
                    dQx = A.r0e[0] - A.sedge[0].comp(comp);
                    dQy = A.r0e[1] - A.sedge[1].comp(comp);
                    dQz = A.r0e[2] - A.sedge[2].comp(comp);
                }
                dGx = (long)(F_.rho - comp) * A.sflux[0].sn;
                dGy = (long)(F_.rho - comp) * A.sflux[1].sn;
                dGz = (long)(F_.rho - comp) * A.sflux[2].sn;
                if (A.do_eta) pEta = A.eta.p + eta_c;
            }
            if (MODE == 2 && sph) {
                dQx = A.r0e[0] - A.sedge[0].comp(comp);
                dQy = A.r0e[1] - A.sedge[1].comp(comp);
                dQz = A.r0e[2] - A.sedge[2].comp(comp);
                dHx = A.h0e[0] - A.sedge[0].comp(comp);
                dHy = A.h0e[1] - A.sedge[1].comp(comp);
                dHz = A.h0e[2] - A.sedge[2].comp(comp);
            }
        }

        __syncthreads();  // every role is done with the previous component; tables written
        double w_m1 = 0.0, w_0 = 0.0;
        double w_p1 = pW[-w_sz];  // face a0
        double w_n = pW[0];  // face a0+1
        double f_n = pF[0];
#pragma unroll
        for (int v = 0; v < 8; ++v) HS(v) = 0.0;
        bool slow_p = true;

        for (int a = a0; a <= k1; ++a) {
            const int it = a - a0;
            const bool bcz = bcz_c && ((a <= g.dlo[2] + 4) || (a >= g.dhi[2] - 4));
            // ---- P0: current-plane registers, prefetch of the next plane ----------------------
            const double f0 = f_n;
            w_m1 = w_0;
            w_0 = w_p1;
            w_p1 = w_n;
            if (a < k1) {
                pW += w_sz;
                pF += f_sz;
                w_n = pW[0];
                f_n = pF[0];
            }
            bar_sync(B6_FULL + it % 3, S6_NT);  // plane a has been reconstructed, its velocities staged
            const double* const I = ibuf + (it % 3) * S6_SLOT + tid;
            const double* const Ip_ = ibuf + ((it + 2) % 3) * S6_SLOT + tid;  // plane a-1
            const double* const Uc = ibuf + (it % 3) * S6_SLOT + 6 * S6_IBS + ty * 33 + tx;
            const double* const Vc = ibuf + (it % 3) * S6_SLOT + 6 * S6_IBS + S6_US + tid;
            const double um = Uc[0], up = Uc[1], vm = Vc[0], vp = Vc[GX];
            const bool px_pos = um > 0.0, px_nz0 = fabs(um) > 0.0, px_nz = fabs(um) > eps;
            const bool py_pos = vm > 0.0, py_nz0 = fabs(vm) > 0.0, py_nz = fabs(vm) > eps;
            const bool pz_pos = w_0 > 0.0, pz_nz0 = fabs(w_0) > 0.0, pz_nz = fabs(w_0) > eps;
            const bool py2_pos = vp > 0.0, py2_nz0 = fabs(vp) > 0.0;
            const bool slow_c = __any_sync(
                0xffffffffu, !(px_nz0 && px_nz && py_nz0 && py_nz && pz_nz0 && pz_nz && py2_nz0));
            const bool slow = slow_c || slow_p;

            // Two copies of the plane body, chosen per warp and plane: the common one has no
            // stagnation arm at all (upwinding = one select); the other also averages where |u|
            // is below the threshold (MaestroMakeEdgeScalars.cpp:745-748).  Both contain the same
            // barrier sequence, so warps of one block may take different copies.
            auto plane = [&](auto slow_tag) {
                constexpr bool SLOW = decltype(slow_tag)::value;
            // ---- P2: 1-D states from the I-buffer, first Riemann solve -----------------------
                double slx = I[0 * S6_IBS], srx = I[1 * S6_IBS];
                double sly = I[2 * S6_IBS], sry = I[3 * S6_IBS];
                double slz = I[4 * S6_IBS], srz = I[5 * S6_IBS];
                if (BC) {
                    if (bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, slx, srx);
                    if (bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, sly, sry);
                    if (bcz) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, slz, srz);
                }
                const double shx = upw(px_pos, px_nz0, SLOW, slx, srx);  // QUIRK: |u| > 0.0 (:746)
                const double shy = upw(py_pos, py_nz0, SLOW, sly, sry);
                const double shz = upw(pz_pos, pz_nz0, SLOW, slz, srz);
                double shy_up;  // first Riemann solve of my HIGH y face (the row above's low face)
                {
                    double l = I[2 * S6_IBS + GX], r = I[3 * S6_IBS + GX];
                    if (BC && bcy) bc_pair(j + 1, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, l, r);
                    shy_up = upw(py2_pos, py2_nz0, SLOW, l, r);
                }

                // ---- P3: transverse differences (dt/6 form) --------------------------------------
                const double usx = up + um, usy = vp + vm;
                const double t6x = (c6x * usx) * (shfl_dn1(shx) - shx);
                const double t6y = (c6y * usy) * (shy_up - shy);
                const double wsum_p = w_0 + w_m1;
                const double tz = (c6z * wsum_p) * (shz - HS(H_SHZ));
                HS(H_SHZ) = shz;
                X_T6X[tid] = t6x;
                X_TZ[tid] = tz;
                const bool do_xy = a > k0;
                const bool do_z = a >= k0 && (a < k1 || last);
                double r_x = 0.0, r_y = 0.0, r_z = 0.0, g_x = 0.0, g_y = 0.0, g_z = 0.0, e_z = 0.0;
                double q_x = 0.0, q_y = 0.0, q_z = 0.0;
                if ((MODE == 1 && flux1) || MODE == 2) {
                    if (need_ser) {
                        if (do_xy) {
                            r_x = pEx[dRx];
                            r_y = pEy[dRy];
                        }
                        if (do_z) r_z = pEz[dRz];
                    }
                    if (MODE == 1 && !first_spec) {
                        if (do_xy) {
                            g_x = pFx[dGx];
                            g_y = pFy[dGy];
                        }
                        if (do_z) g_z = pFz[dGz];
                    }
                    if (MODE == 1 && A.do_eta && do_z) e_z = pEta[0];
                    if (sph) {  // time-centred base state on this face (mac_average output)
                        if (do_xy) {
                            q_x = pEx[dQx];
                            q_y = pEy[dQy];
                        }
                        if (do_z) q_z = pEz[dQz];
                    }
                }
                bar_sync(B6_Q1, S6_NQ);  // S4

                // ---- P4: corner-coupled states ---------------------------------------------------
                // velocities of plane a-1 from its slot (instead of a thread-private history)
                const double* const Up = ibuf + ((it + 2) % 3) * S6_SLOT + 6 * S6_IBS + ty * 33 + tx;
                const double* const Vp = ibuf + ((it + 2) % 3) * S6_SLOT + 6 * S6_IBS + S6_US + tid;
                const double uf_p = Up[0], vf_p = Vp[0];
                const bool qx_pos = uf_p > 0.0, qx_nz = fabs(uf_p) > eps;
                const bool qy_pos = vf_p > 0.0, qy_nz = fabs(vf_p) > eps;
                double slx_p = Ip_[0 * S6_IBS], srx_p = Ip_[1 * S6_IBS];
                double sly_p = Ip_[2 * S6_IBS], sry_p = Ip_[3 * S6_IBS];
                if (BC) {  // the plane below went through the same boundary overrides
                    if (bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, slx_p, srx_p);
                    if (bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, sly_p, sry_p);
                }
                double hxy, hyx, hzx, hzy, hxz, hyz;
                {
                    double l = slx - shfl_up1(t6y), r = srx - t6y;
                    if (BC && bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, l, r);
                    hxy = upw(px_pos, px_nz, SLOW, l, r);
                    l = sly - X_T6X[tid - GX];
                    r = sry - t6x;
                    if (BC && bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, l, r);
                    hyx = upw(py_pos, py_nz, SLOW, l, r);
                    l = slz - HS(H_T6X);
                    r = srz - t6x;
                    if (BC && bcz) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, l, r);
                    hzx = upw(pz_pos, pz_nz, SLOW, l, r);
                    l = slz - HS(H_T6Y);
                    r = srz - t6y;
                    if (BC && bcz) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, l, r);
                    hzy = upw(pz_pos, pz_nz, SLOW, l, r);
                    l = slx_p - shfl_up1(tz);
                    r = srx_p - tz;
                    if (BC && bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, l, r);
                    hxz = upw(qx_pos, qx_nz, SLOW, l, r);
                    l = sly_p - X_TZ[tid - GX];
                    r = sry_p - tz;
                    if (BC && bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, l, r);
                    hyz = upw(qy_pos, qy_nz, SLOW, l, r);
                }
                HS(H_T6X) = t6x;
                HS(H_T6Y) = t6y;
                X_HYX[tid] = hyx;
                X_HYZ[tid] = hyz;
                const double c4zw = c4z * wsum_p;
                const double t4xz = (c4x * (Up[1] + uf_p)) * (shfl_dn1(hxz) - hxz);
                const double t4zx = c4zw * (hzx - HS(H_HZX));
                const double fs_prev = HS(H_FS);
                X_QY1[tid] = t4xz;
                X_QY2[tid] = t4zx;
                X_FS[tid] = fs_prev;
                bar_sync(B6_Q2, S6_NQ);  // S5

                // ---- P5: transverse differences (dt/4 form) --------------------------------------
                const double c4xu = c4x * usx, c4yu = c4y * usy;
                const double t4xy = c4xu * (shfl_dn1(hxy) - hxy);
                const double t4yx = c4yu * (X_HYX[tid + GX] - hyx);
                const double t4yz = (c4y * (Vp[GX] + vf_p)) * (X_HYZ[tid + GX] - hyz);
                const double t4zy = c4zw * (hzy - HS(H_HZY));
                const double fs_cur = A.dt2 * f0;

                // ---- P6: final edge states (+ fluxes) --------------------------------------------
                if (do_xy) {
                    double ex_v, ey_v;
                    {
                        const double el = ((slx_p - shfl_up1(t4yz)) - shfl_up1(t4zy)) + shfl_up1(fs_prev);
                        const double er = ((srx_p - t4yz) - t4zy) + fs_prev;
                        ex_v = upw(qx_pos, qx_nz, SLOW, el, er);
                        if (BC && bcx) ex_v = bc_final(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, el, er, ex_v);
                    }
                    {
                        const double el = ((sly_p - X_QY1[tid - GX]) - X_QY2[tid - GX]) + X_FS[tid - GX];
                        const double er = ((sry_p - t4xz) - t4zx) + fs_prev;
                        ey_v = upw(qy_pos, qy_nz, SLOW, el, er);
                        if (BC && bcy) ey_v = bc_final(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, el, er, ey_v);
                    }
                    if (ownx) pEx[0] = ex_v;
                    if (owny) pEy[0] = ey_v;
                    if (MODE == 1 && flux1) {
                        double r0x, r0y;
                        if (sph) {
                            r0x = q_x;
                            r0y = q_y;
                        } else {
                            r0x = r0y = T_RXY[it - 1];
                        }
                        const double flx = rhoX_flux(F_.spt, uf_p, r0x, r_x, ex_v);
                        const double fly = rhoX_flux(F_.spt, vf_p, r0y, r_y, ey_v);
                        if (ownx) {
                            pFx[0] = flx;
                            pFx[dGx] = g_x + flx;
                        }
                        if (owny) {
                            pFy[0] = fly;
                            pFy[dGy] = g_y + fly;
                        }
                    } else if (MODE == 2 && sph) {
                        // MakeRhoHFlux (spherical), MaestroMakeFlux.cpp:871-1021
                        double fl;
                        if (rhoh_flux_sph(F_, uf_p, r_x, ex_v, q_x, pEx[dHx], &fl))
                            if (ownx) pFx[0] = fl;
                        if (rhoh_flux_sph(F_, vf_p, r_y, ey_v, q_y, pEy[dHy], &fl))
                            if (owny) pFy[0] = fl;
                    } else if (MODE == 2) {
                        double fl;
                        if (rhoh_flux_tab(F_, have_h, uf_p, r_x, ex_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                            if (ownx) pFx[0] = fl;
                        if (rhoh_flux_tab(F_, have_h, vf_p, r_y, ey_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                            if (owny) pFy[0] = fl;
                    }
                }
                const double el_z = ((slz - HS(H_T4XY)) - HS(H_T4YX)) + fs_prev;
                const double er_z = ((srz - t4xy) - t4yx) + fs_cur;
                if (do_z) {
                    double ez_v = upw(pz_pos, pz_nz, SLOW, el_z, er_z);
                    if (BC && bcz) ez_v = bc_final(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, el_z, er_z, ez_v);
                    if (ownz) pEz[0] = ez_v;
                    if (MODE == 1 && flux1) {
                        double r0z;
                        if (sph) {
                            r0z = q_z;
                        } else {
                            r0z = T_RZ[it];
                        }
                        const double flz = rhoX_flux(F_.spt, w_0, r0z, r_z, ez_v);
                        if (ownz) {
                            pFz[0] = flz;
                            pFz[dGz] = g_z + flz;
                        }
                        if (A.do_eta) {
                            double ev = e_z;
                            if (is_spec) ev += flz;
                            if (last_spec) ev -= T_W0R[it];
                            if (ownz) pEta[0] = ev;
                        }
                    } else if (MODE == 2 && sph) {
                        double fl;
                        if (rhoh_flux_sph(F_, w_0, r_z, ez_v, q_z, pEz[dHz], &fl))
                            if (ownz) pFz[0] = fl;
                    } else if (MODE == 2) {
                        double fl;
                        if (rhoh_flux_tab(F_, have_h, w_0, r_z, ez_v, T_RZ[it], T_HZ[it], true, &fl))
                            if (ownz) pFz[0] = fl;
                    }
                }
                pEx += ex_sz;
                pEy += ey_sz;
                pEz += ez_sz;
                if (MODE != 0) {
                    pFx += fx_sz;
                    pFy += fy_sz;
                    pFz += fz_sz;
                    if (MODE == 1) pEta += eta_sz;
                }
                // ---- z history for the next plane; release the slot of plane a-1 --------------------
                HS(H_HZX) = hzx;
                HS(H_HZY) = hzy;
                HS(H_T4XY) = t4xy;
                HS(H_T4YX) = t4yx;
                HS(H_FS) = fs_cur;
            };
#ifdef MX_COUNT_FAST_PATH  // static instruction counting of the common copy only (tools/loopstat.py)
            plane(std::false_type{});
#else
            if (slow)
                plane(std::true_type{});
            else
                plane(std::false_type{});
#endif
            slow_p = slow_c;
            if (a + 2 <= k1) bar_arrive(B6_EMPTY + (it + 2) % 3, S6_NT);
        }
    }
#undef HS
}

template <int PPM, int MODE, bool SPH>
__global__ void __launch_bounds__(S6_NT, 1) k_stream6(const __grid_constant__ FusedArgs A) {
    extern __shared__ __align__(128) double sm6[];
    constexpr int RD = S6_RS * s6_pls<PPM>();
    const KGeom& g = A.g;
    const S6Tile T = s6_tile(A);
    bool bc = false;
    if (A.bcdir[0]) bc |= (T.x0 - 1 <= g.dlo[0] + 2) || (T.x0 + 30 >= g.dhi[0] - 2);
    if (A.bcdir[1]) bc |= (T.y0 - 1 <= g.dlo[1] + 2) || (T.y0 + 14 >= g.dhi[1] - 2);
    if (A.bcdir[2]) bc |= (T.k0 - 1 <= g.dlo[2] + 2) || (T.k1 >= g.dhi[2] - 2);

    // ---- block-wide set-up: mbarriers, I-buffer cleared, base-state tables ----------------------
    double* const ibuf = sm6 + RD;
    double* const tab = ibuf + S6_IS * S6_SLOT + 7 * S6_XS + 8 * S6_NQ;
    unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(tab + 5 * (ST_MAXZC + 2) + 2);
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < S6_RS; ++q) mbar_init(mbar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int q = tid; q < S6_IS * S6_SLOT; q += S6_NT) ibuf[q] = 0.0;
    if (MODE != 0 && !SPH) {
        const KFlux& F_ = A.f;
        const int a0 = T.k0 - 1;
        double* T_RXY = tab;
        double* T_RZ = tab + (ST_MAXZC + 2);
        double* T_W0R = tab + 2 * (ST_MAXZC + 2);
        double* T_HXY = tab + 3 * (ST_MAXZC + 2);
        double* T_HZ = tab + 4 * (ST_MAXZC + 2);
        for (int q = tid; q <= T.k1 - a0; q += S6_NT) {
            const int r = a0 + q;
            double rxy = 0.0, rz = 0.0, w0r = 0.0, hxy = 0.0, hz = 0.0;
            if (F_.rho0_old && F_.rho0_new && r >= 0 && r <= g.dhi[2])
                rxy = 0.5 * (F_.rho0_old[r * F_.st_rho0_old] + F_.rho0_new[r * F_.st_rho0_new]);
            if (F_.rho0_edge_old && F_.rho0_edge_new && r >= 0 && r <= g.dhi[2] + 1)
                rz = 0.5 * (F_.rho0_edge_old[r * F_.st_rho0_edge_old] +
                            F_.rho0_edge_new[r * F_.st_rho0_edge_new]);
            if (MODE == 1 && A.do_eta && r >= 0 && r <= g.dhi[2] + 1)
                w0r = F_.w0[r * F_.st_w0] * F_.rho0_pred_edge[r * F_.st_rho0_pred_edge];
            if (MODE == 2) {
                if (F_.rhoh0_old && F_.rhoh0_new && r >= 0 && r <= g.dhi[2])
                    hxy = 0.5 * (F_.rhoh0_old[r * F_.st_rhoh0_old] + F_.rhoh0_new[r * F_.st_rhoh0_new]);
                if (F_.rhoh0_edge_old && F_.rhoh0_edge_new && r >= 0 && r <= g.dhi[2] + 1)
                    hz = 0.5 * (F_.rhoh0_edge_old[r * F_.st_rhoh0_edge_old] +
                                F_.rhoh0_edge_new[r * F_.st_rhoh0_edge_new]);
            }
            T_RXY[q] = rxy;
            T_RZ[q] = rz;
            T_W0R[q] = w0r;
            T_HXY[q] = hxy;
            T_HZ[q] = hz;
        }
    }
    // (the first __syncthreads of either role publishes all of the above)

    // Register budgets.  The kernel is launched at 96 registers per thread (640 threads); the
    // four Riemann warpgroups give 8 each back to the CTA pool, which is exactly what the one
    // reconstruction warpgroup needs to grow to 128 (setmaxnreg.inc only draws on registers
    // released by setmaxnreg.dec of the same CTA): 512 x 88 + 128 x 128 = 640 x 96.
    if (tid >= S6_NQ) {
        if (S6_RREG > S6_LREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(S6_RREG));
        if (bc)
            stream6_recon<PPM, true>(A, sm6);
        else
            stream6_recon<PPM, false>(A, sm6);
    } else {
        if (S6_QREG < S6_LREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(S6_QREG));
        if (bc)
            stream6_riemann<true, MODE, SPH>(A, sm6, RD);
        else
            stream6_riemann<false, MODE, SPH>(A, sm6, RD);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// The scalar ring is filled by TMA from a 4-D tensor map (x, y, z, component) over the caller's
// ghost-grown state fab.  TMA needs a 16-byte aligned base and global strides that are
// multiples of 16 bytes, i.e. an even number of doubles per fab row (every BASELINE
// configuration: 518, 514, 520); boxes with an odd row length take k_stream4 (cp.async).
typedef CUresult (*mx_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                       const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline mx_encode_tiled_fn encode_tiled() {
    static std::atomic<void*> fn{nullptr};
    void* f = fn.load(std::memory_order_acquire);
    if (f == nullptr) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn.store(f, std::memory_order_release);
    }
    return reinterpret_cast<mx_encode_tiled_fn>(f);
}

template <int PPM>
bool stream6_tmap(FusedArgs& A) {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    const DFab& st = A.state;
    if ((reinterpret_cast<uintptr_t>(st.p) & 15u) != 0 || (st.n[0] & 1) != 0) return false;
    mx_encode_tiled_fn enc = encode_tiled();
    if (enc == nullptr) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)st.n[0], (cuuint64_t)st.n[1], (cuuint64_t)st.n[2],
                                (cuuint64_t)st.ncomp};
    const cuuint64_t strides[3] = {(cuuint64_t)st.sy * 8, (cuuint64_t)st.sz * 8, (cuuint64_t)st.sn * 8};
    const cuuint32_t box[4] = {32 + 2 * LA, 16 + 2 * LA, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(&A.tmap_state, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, st.p, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int PPM, int MODE, bool SPH>
int launch_stream6(const FusedArgs& A, cudaStream_t stream) {
    const size_t smem = sizeof(double) * (size_t)stream6_smem_doubles<PPM>();
    static std::atomic<unsigned long long> attr_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & bit)) {
        if (cudaFuncSetAttribute(k_stream6<PPM, MODE, SPH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return -1;
        attr_set.fetch_or(bit, std::memory_order_relaxed);
    }
    dim3 grid(A.ntx * A.nty, A.nzc);
    k_stream6<PPM, MODE, SPH><<<grid, S6_NT, smem, stream>>>(A);
    mx_count_launch();
    mx_count_tma_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// > 0: this box cannot take the TMA kernel (odd fab rows / unaligned base): use k_stream4
template <int PPM>
int launch_stream6_mode(FusedArgs& A, cudaStream_t stream) {
    if (!stream6_tmap<PPM>(A)) return 1;
    if (A.mode == 0) return launch_stream6<PPM, 0, false>(A, stream);
    if (A.mode == 1)
        return A.f.spherical ? launch_stream6<PPM, 1, true>(A, stream)
                             : launch_stream6<PPM, 1, false>(A, stream);
    return A.f.spherical ? launch_stream6<PPM, 2, true>(A, stream)
                         : launch_stream6<PPM, 2, false>(A, stream);
}

#endif
