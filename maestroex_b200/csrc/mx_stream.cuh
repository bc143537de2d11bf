// mx_stream.cuh — the streaming edge-state (+flux) kernel k_stream: the performance path of
// libmxadv for 3-D, non-conservative scalar sweeps (included by mx_fused.cu, which owns
// FusedArgs and the host-side dispatch).
//
// Decomposition: a thread block of 32 x 16 threads owns a tile of
// 30 x 14 valid cells (one thread per cell of the tile grown by one; a warp is one x-row) and
// a chunk of z planes, and marches upward in z as a software pipeline: iteration `a`
//   P0  issues cp.async for scalar plane a+LA+1 into a shared-memory ring, prefetches the
//       MAC velocities / force of plane a+1 into registers and the rho'-edge / rho-flux
//       entries it will read-modify-write into L1;
//   P1  reconstructs plane a in x and y (slopes or PPM parabolas) from the ring;
//   P2  extends the z parabola incrementally (one new van Leer slope + one new edge value per
//       plane, kept in registers), forms the 1-D left/right states and the first upwind;
//   P3  transverse differences of the dt/6 form;  P4 the six corner-coupled states;
//   P5  transverse differences of the dt/4 form;  P6 final edge states of the x/y faces of
//       plane a-1 and of z face a, and (density / enthalpy mode) the fluxes.
// x-neighbour values travel by warp shuffle, y-neighbour values through shared memory (five
// barriers per plane), z-neighbour values stay in the thread: registers for what the
// reconstruction needs, a thread-private shared-memory slot for what only P4-P6 need (this
// keeps ~26 registers free for instruction-level parallelism in P1/P2).
// Every lane computes; only stores are masked.  Physical-boundary code is compiled out of
// blocks that cannot touch the domain boundary (BC = false instantiation).
#ifndef MX_STREAM_CUH
#define MX_STREAM_CUH

MXD double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
MXD double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

MXD void cp_async8(double* dst_smem, const double* src, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = pred ? 8 : 0;  // src-size 0: zero-fill, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(n)
                 : "memory");
}
MXD void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
MXD void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
MXD void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

// x / h_D for the loop-invariant mesh spacing, correctly rounded (== IEEE division): a power
// of two multiplies exactly; otherwise q0 = RN(x * RN(1/h)) is refined twice with exact FMA
// residuals.  After the first step q is a faithful rounding of x/h, so the second step is
// Markstein's theorem (RN(q + RN(x - h q) * RN(1/h)) = RN(x / h)).  Barring under/overflow
// of the intermediate products (|x| = |u| dt is never near either); tests/test_divconst.py
// checks the sequence against exact rational arithmetic.
template <int D>
MXD double div_h(const FusedArgs& A, double x) {
    const double y = A.inv_h[D];
    double q = x * y;
    if (!A.inv_exact[D]) {
        const double h = A.h[D];
        double r = __fma_rn(-h, q, x);
        q = __fma_rn(r, y, q);
        r = __fma_rn(-h, q, x);
        q = __fma_rn(r, y, q);
    }
    return q;
}

constexpr int ST_MAXZC = 64;  // largest z chunk (size of the base-state tables)

// shared-memory footprint in doubles
template <int PPM>
constexpr int stream_smem_doubles() {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    return (LA + 2) * (32 + 2 * LA) * (16 + 2 * LA) + 9 * 512 + 32 + 13 * 512 + 5 * (ST_MAXZC + 2);
}

template <int PPM, bool BC, int MODE>
__device__ __forceinline__ void stream_body(const FusedArgs& A, double* smem) {
    constexpr int GX = 32, GY = 16;
    constexpr int LA = (PPM == 2) ? 3 : 2;  // stencil radius
    constexpr int RS = LA + 2;              // ring: planes a .. a+LA resident, a+LA+1 in flight
    constexpr int PW = GX + 2 * LA, PH = GY + 2 * LA;
    constexpr int NT = GX * GY;
    constexpr int BX = GX - 2, BY = GY - 2;
    constexpr int NLD = (PW * PH + NT - 1) / NT;

    double* ring = smem;
    double* ex = ring + RS * PW * PH;  // y-neighbour exchange arrays
    double* X_IPY = ex + 0 * NT;
    double* X_SY = ex + 1 * NT;
    double* X_T6X = ex + 2 * NT;
    double* X_TZ = ex + 3 * NT;
    double* X_HYX = ex + 4 * NT;
    double* X_HYZ = ex + 5 * NT;
    double* X_QY1 = ex + 6 * NT;
    double* X_QY2 = ex + 7 * NT;
    double* X_FS = ex + 8 * NT;
    double* hist = ex + 9 * NT + GX;  // thread-private z history (slot v of thread tid)
    double* tab = hist + 13 * NT;     // per-plane base-state values of this chunk
    double* T_RXY = tab;              // 0.5 (rho0_old + rho0_new)(p)          x/y faces
    double* T_RZ = tab + (ST_MAXZC + 2);       // 0.5 (rho0_edge_old + _new)(a) z faces
    double* T_W0R = tab + 2 * (ST_MAXZC + 2);  // w0(a) * rho0_predicted_edge(a)
    double* T_HXY = tab + 3 * (ST_MAXZC + 2);  // 0.5 (rhoh0_old + rhoh0_new)(p)
    double* T_HZ = tab + 4 * (ST_MAXZC + 2);   // 0.5 (rhoh0_edge_old + _new)(a)

    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
#define HS(v) hist[(v) * NT + tid]
    enum { H_SLX, H_SRX, H_SLY, H_SRY, H_UF, H_VF, H_C4X, H_C4Y, H_HZX, H_HZY, H_T4XY, H_T4YX, H_FS };

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
    const bool vxi = inG && tx >= 1 && i <= thx;
    const bool vyj = inG && ty >= 1 && j <= thy;
    const bool fx = inG && tx >= 1;
    const bool fy = inG && ty >= 1;
    // faces this block stores: the tile's low faces, plus the high face only on the last tile
    const bool ownx = fx && vyj && (i <= thx || thx == g.vhi[0]);
    const bool owny = fy && vxi && (j <= thy || thy == g.vhi[1]);
    const bool ownz = vxi && vyj;
    const int il = min(i, thx + 1), jl = min(j, thy + 1);  // clamped: every load in bounds
    const int rc = (ty + LA) * PW + (tx + LA);

    const DFab& st = A.state;
    const int st_sz = (int)st.sz, f_sz = (int)A.force.sz;
    const int u_sz = (int)A.umac[0].sz, v_sz = (int)A.umac[1].sz, w_sz = (int)A.umac[2].sz;
    const int v_sy = (int)A.umac[1].sy;
    // per-thread column offsets (plane k adds k * sz; the -lo[2]*sz term is folded in)
    const int s_co = (int)st.off(il, jl, 0);
    const int f_co = (int)A.force.off(il, jl, 0);
    const int u_co = (int)A.umac[0].off(il, jl, 0);
    const int v_co = (int)A.umac[1].off(il, jl, 0);
    const int w_co = (int)A.umac[2].off(il, jl, 0);
    const int ex_co = (int)A.sedge[0].off(il, jl, 0), ey_co = (int)A.sedge[1].off(il, jl, 0),
              ez_co = (int)A.sedge[2].off(il, jl, 0);
    const int ex_sz = (int)A.sedge[0].sz, ey_sz = (int)A.sedge[1].sz, ez_sz = (int)A.sedge[2].sz;
    int fx_co = 0, fy_co = 0, fz_co = 0, fx_sz = 0, fy_sz = 0, fz_sz = 0, eta_co = 0, eta_sz = 0;
    if (MODE != 0) {
        fx_co = (int)A.sflux[0].off(il, jl, 0);
        fy_co = (int)A.sflux[1].off(il, jl, 0);
        fz_co = (int)A.sflux[2].off(il, jl, 0);
        fx_sz = (int)A.sflux[0].sz;
        fy_sz = (int)A.sflux[1].sz;
        fz_sz = (int)A.sflux[2].sz;
        if (MODE == 1 && A.do_eta) {
            eta_co = (int)A.eta.off(il, jl, 0);
            eta_sz = (int)A.eta.sz;
        }
    }
    const double* __restrict__ Up = A.umac[0].p;
    const double* __restrict__ Vp = A.umac[1].p;
    const double* __restrict__ Wp = A.umac[2].p;
    const KFlux& F_ = A.f;
    const bool sph = (MODE != 0) && F_.spherical;

    // ---- base-state tables of this chunk (planes a0 .. k1) -------------------------------
    if (MODE != 0 && !sph) {
        for (int q = tid; q <= k1 - a0; q += NT) {
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
                    hxy = 0.5 * (F_.rhoh0_old[r * F_.st_rhoh0_old] +
                                 F_.rhoh0_new[r * F_.st_rhoh0_new]);
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

    // cp.async staging plan of this thread: element e = tid + q*NT of a PW x PH plane
    int ld_off[NLD];
    bool ld_do[NLD], ld_in[NLD];
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
        const int e = tid + q * NT;
        const int px = e % PW, py = e / PW;
        const int gi = x0 - 1 - LA + px - st.lo[0], gj = y0 - 1 - LA + py - st.lo[1];
        ld_do[q] = e < PW * PH;
        ld_in[q] = ld_do[q] && gi >= 0 && gi < st.n[0] && gj >= 0 && gj < st.n[1];
        ld_off[q] = ld_in[q] ? gj * (int)st.sy + gi : 0;
    }

    const double c6x = A.c6h[0], c6y = A.c6h[1], c6z = A.c6h[2];
    const double c4x = A.c4h[0], c4y = A.c4h[1], c4z = A.c4h[2];
    const double eps = A.eps, dt = A.dt;
    // flux-form switches (block uniform)
    const bool have_h = (MODE == 2) && (F_.ept == MX_PRED_H || F_.ept == MX_PRED_T_THEN_H ||
                                        F_.ept == MX_PRED_TPRIME_THEN_H);
    const bool need_ser = (MODE == 1) ? (F_.spt != MX_PRED_RHOX) : have_h;

    for (int n = 0; n < A.ncomp; ++n) {
        const int comp = A.comp[n];
        const double* __restrict__ sp_ = st.comp(comp);
        const double* __restrict__ fp_ = A.force.comp(comp);
        double* __restrict__ Ex = A.sedge[0].comp(comp);
        double* __restrict__ Ey = A.sedge[1].comp(comp);
        double* __restrict__ Ez = A.sedge[2].comp(comp);
        const int bclx = A.bclo[n][0], bchx = A.bchi[n][0];
        const int bcly = A.bclo[n][1], bchy = A.bchi[n][1];
        const int bclz = A.bclo[n][2], bchz = A.bchi[n][2];
        const bool lobcz = BC && (bclz == BC_EXT_DIR || bclz == BC_HOEXTRAP);
        const bool hibcz = BC && (bchz == BC_EXT_DIR || bchz == BC_HOEXTRAP);
        const bool flux1 = (MODE == 1 && n > 0);
        const bool first_spec = (n == 1);
        const bool is_spec =
            comp >= F_.first_spec && comp <= F_.first_spec + F_.num_spec - 1;  // eta: :288-301
        const bool last_spec = comp == F_.first_spec + F_.num_spec - 1;
        // rho' edge state / rho flux component (density + h-type enthalpy forms)
        const double* __restrict__ Rx = A.sedge[0].comp(F_.rho);
        const double* __restrict__ Ry = A.sedge[1].comp(F_.rho);
        const double* __restrict__ Rz = A.sedge[2].comp(F_.rho);

        auto stage_plane = [&](int k, int slot) {
            double* dst = ring + slot * PW * PH;
            const double* src = sp_ + (long)(k - st.lo[2]) * st_sz;
#pragma unroll
            for (int q = 0; q < NLD; ++q)
                if (ld_do[q]) cp_async8(dst + tid + q * NT, src + ld_off[q], ld_in[q]);
        };

        // ---- prologue ---------------------------------------------------------------
        __syncthreads();  // previous component is done with shared memory; tables written
#pragma unroll
        for (int q = 0; q <= LA; ++q) stage_plane(a0 + q, q);
        cp_async_commit();
        double zw[2 * LA + 1];
#pragma unroll
        for (int m = 0; m < 2 * LA; ++m) zw[m + 1] = sp_[s_co + (a0 - LA + m) * st_sz];
        zw[0] = 0.0;
        // incremental z reconstruction state
        double dz_c = 0.0, ez_lo = 0.0;                // PPM 1: V(a), E(a-1/2)
        double ez_l2 = 0.0, ez_l1 = 0.0, ez_h1 = 0.0;  // PPM 2: E(a-3/2), E(a-1/2), E(a+1/2)
        if (PPM == 1) {
            const double dm1 = vanleer(zw[1], zw[2], zw[3]);
            dz_c = vanleer(zw[2], zw[3], zw[4]);
            ez_lo = vl_edge(zw[2], zw[3], dm1, dz_c);
        } else if (PPM == 2) {
            ez_l2 = cs_edge(zw[1], zw[2], zw[3], zw[4]);
            ez_l1 = cs_edge(zw[2], zw[3], zw[4], zw[5]);
            ez_h1 = cs_edge(zw[3], zw[4], zw[5], zw[6]);
        }
        double w_m1 = 0.0, w_0 = 0.0;
        double w_p1 = Wp[w_co + a0 * w_sz];
        double sgz_hi = (PPM != 0) ? div_h<2>(A, fabs(w_p1) * dt) : 0.0;  // sigma of face a0
        // register prefetch of plane a0
        double um_n = Up[u_co + a0 * u_sz];
        double up31_n = (tx == 31) ? Up[u_co + a0 * u_sz + 1] : 0.0;
        double vm_n = Vp[v_co + a0 * v_sz];
        double vp_n = Vp[v_co + a0 * v_sz + v_sy];
        double w_n = Wp[w_co + (a0 + 1) * w_sz];
        double f_n = fp_[f_co + a0 * f_sz];

        double ipz_prev = 0.0, shz_prev = 0.0, t6x_prev = 0.0, t6y_prev = 0.0;
        HS(H_SLX) = 0.0; HS(H_SRX) = 0.0; HS(H_SLY) = 0.0; HS(H_SRY) = 0.0;
        HS(H_UF) = 0.0; HS(H_VF) = 0.0; HS(H_C4X) = 0.0; HS(H_C4Y) = 0.0;
        HS(H_HZX) = 0.0; HS(H_HZY) = 0.0; HS(H_T4XY) = 0.0; HS(H_T4YX) = 0.0; HS(H_FS) = 0.0;
        bool slow_p = true;
        cp_async_wait<0>();
        __syncthreads();  // planes a0 .. a0+LA are in the ring

        for (int a = a0; a <= k1; ++a) {
            const int it = a - a0;
            const int p = a - 1;
            const double* pa = ring + (it % RS) * PW * PH;
            // ---- P0: current-plane registers, prefetch of the next plane ----------------
            const double um = um_n, vm = vm_n, vp = vp_n, f0 = f_n;
            double up = shfl_dn1(um);
            if (tx == 31) up = up31_n;
            w_m1 = w_0;
            w_0 = w_p1;
            w_p1 = w_n;
            if (a + 1 <= k1) stage_plane(a + LA + 1, (it + LA + 1) % RS);
            cp_async_commit();
            {
                const int an = min(a + 1, k1);
                um_n = Up[u_co + an * u_sz];
                if (tx == 31) up31_n = Up[u_co + an * u_sz + 1];
                vm_n = Vp[v_co + an * v_sz];
                vp_n = Vp[v_co + an * v_sz + v_sy];
                w_n = Wp[w_co + min(a + 2, k1 + 1) * w_sz];
                f_n = fp_[f_co + an * f_sz];
            }
            const bool wx = ownx && a >= k0 + 1, wy = owny && a >= k0 + 1;
            const bool wz = ownz && a >= k0 && (a <= k1 - 1 || last);
            if ((MODE == 1 && flux1) || (MODE == 2 && need_ser)) {
                // what P6 of this iteration reads (rho' edges) and read-modify-writes (rho
                // flux, etarhoflux): pull the lines into L1 now, a whole plane ahead
                if (need_ser) {
                    if (wx) prefetch_l1(Rx + (ex_co + p * ex_sz));
                    if (wy) prefetch_l1(Ry + (ey_co + p * ey_sz));
                    if (wz) prefetch_l1(Rz + (ez_co + a * ez_sz));
                }
                if (MODE == 1 && !first_spec) {
                    if (wx) prefetch_l1(A.sflux[0].comp(F_.rho) + (fx_co + p * fx_sz));
                    if (wy) prefetch_l1(A.sflux[1].comp(F_.rho) + (fy_co + p * fy_sz));
                    if (wz) prefetch_l1(A.sflux[2].comp(F_.rho) + (fz_co + a * fz_sz));
                }
                if (MODE == 1 && A.do_eta && wz) prefetch_l1(A.eta.p + (eta_co + a * eta_sz));
            }
            // velocity predicates of this plane's faces
            const bool px_pos = um > 0.0, px_nz0 = fabs(um) > 0.0, px_nz = fabs(um) > eps;
            const bool py_pos = vm > 0.0, py_nz0 = fabs(vm) > 0.0, py_nz = fabs(vm) > eps;
            const bool pz_pos = w_0 > 0.0, pz_nz0 = fabs(w_0) > 0.0, pz_nz = fabs(w_0) > eps;
            const bool slow_c = __any_sync(
                0xffffffffu, !(px_nz0 && px_nz && py_nz0 && py_nz && pz_nz0 && pz_nz));
            const bool slow = slow_c || slow_p;

            // ---- P1: x / y reconstruction of plane a ------------------------------------
            const double s0 = pa[rc];
            double ipx, imx, ipy, imy;
            {
                const bool lobcx = BC && (bclx == BC_EXT_DIR || bclx == BC_HOEXTRAP);
                const bool hibcx = BC && (bchx == BC_EXT_DIR || bchx == BC_HOEXTRAP);
                const bool lobcy = BC && (bcly == BC_EXT_DIR || bcly == BC_HOEXTRAP);
                const bool hibcy = BC && (bchy == BC_EXT_DIR || bchy == BC_HOEXTRAP);
                double S[7], T[7];
                S[0] = S[6] = T[0] = T[6] = 0.0;
#pragma unroll
                for (int q = -LA; q <= LA; ++q) {
                    S[q + 3] = pa[rc + q];
                    T[q + 3] = pa[rc + q * PW];
                }
                if (PPM == 0) {
                    ipx = slope_any(A.slope_order, S + 1, i, g.dlo[0], g.dhi[0], lobcx, hibcx, true);
                    ipy = slope_any(A.slope_order, T + 1, j, g.dlo[1], g.dhi[1], lobcy, hibcy, false);
                    imx = imy = 0.0;
                } else {
                    double smx, spx, smy, spy;
                    if (PPM == 1) {
                        ppm1_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, smx, spx);
                        ppm1_parabola<BC>(T, j, g.dlo[1], g.dhi[1], lobcy, hibcy, smy, spy);
                    } else {
                        ppm2_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, true, smx, spx);
                        ppm2_parabola<BC>(T, j, g.dlo[1], g.dhi[1], lobcy, hibcy, false, smy, spy);
                    }
                    // sigma = |u| dt / h once per face: the x high face is the right
                    // neighbour's low face (lane 31 never needs Ip)
                    const double sgx_m = div_h<0>(A, fabs(um) * dt);
                    const double sgx_p = shfl_dn1(sgx_m);
                    const double sgy_m = div_h<1>(A, fabs(vm) * dt);
                    const double sgy_p = div_h<1>(A, fabs(vp) * dt);
                    const double s6x = 6.0 * s0 - 3.0 * (smx + spx);
                    const double s6y = 6.0 * s0 - 3.0 * (smy + spy);
                    ipx = ppm_Ip(smx, spx, s0, s6x, up, sgx_p, eps);
                    imx = ppm_Im(smx, spx, s0, s6x, um, sgx_m, eps);
                    ipy = ppm_Ip(smy, spy, s0, s6y, vp, sgy_p, eps);
                    imy = ppm_Im(smy, spy, s0, s6y, vm, sgy_m, eps);
                }
            }
            X_IPY[tid] = ipy;
            cp_async_wait<1>();  // this thread's share of plane a+LA has landed
            __syncthreads();     // S2: publishes IPY and ring plane a+LA

            // ---- P2: z reconstruction, 1-D states, first Riemann solve -------------------
#pragma unroll
            for (int m = 0; m < 2 * LA; ++m) zw[m] = zw[m + 1];
            zw[2 * LA] = ring[((it + LA) % RS) * PW * PH + rc];
            double ipz, imz;
            {
                double S[7];
                S[0] = S[6] = 0.0;
#pragma unroll
                for (int q = -LA; q <= LA; ++q) S[q + 3] = zw[q + LA];
                if (PPM == 0) {
                    ipz = slope_any(A.slope_order, S + 1, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false);
                    imz = 0.0;
                } else {
                    double sm, sp;
                    if (PPM == 1) {
                        const double dz_n = vanleer(S[3], S[4], S[5]);
                        const double ez_hi = vl_edge(S[3], S[4], dz_c, dz_n);
                        ppm1_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, ez_lo, ez_hi, sm, sp);
                        dz_c = dz_n;
                        ez_lo = ez_hi;
                    } else {
                        const double ez_h2 = cs_edge(S[3], S[4], S[5], S[6]);
                        ppm2_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false, ez_l2, ez_l1,
                                        ez_h1, ez_h2, sm, sp);
                        ez_l2 = ez_l1;
                        ez_l1 = ez_h1;
                        ez_h1 = ez_h2;
                    }
                    const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
                    const double sgz_lo = sgz_hi;  // face a: last plane's high face
                    sgz_hi = div_h<2>(A, fabs(w_p1) * dt);
                    ipz = ppm_Ip(sm, sp, s0, s6, w_p1, sgz_hi, eps);
                    imz = ppm_Im(sm, sp, s0, s6, w_0, sgz_lo, eps);
                }
            }
            double slx, srx, sly, sry, slz, srz;
            if (PPM == 0) {
                // MaestroMakeEdgeScalars.cpp:697-703
                const double qx = div_h<0>(A, dt * um), qy = div_h<1>(A, dt * vm),
                             qz = div_h<2>(A, dt * w_0);
                slx = pa[rc - 1] + 0.5 * (1.0 - qx) * shfl_up1(ipx);
                srx = s0 - 0.5 * (1.0 + qx) * ipx;
                sly = pa[rc - PW] + 0.5 * (1.0 - qy) * X_IPY[tid - GX];
                sry = s0 - 0.5 * (1.0 + qy) * ipy;
                slz = zw[LA - 1] + 0.5 * (1.0 - qz) * ipz_prev;
                srz = s0 - 0.5 * (1.0 + qz) * ipz;
            } else {
                slx = shfl_up1(ipx);
                srx = imx;
                sly = X_IPY[tid - GX];
                sry = imy;
                slz = ipz_prev;
                srz = imz;
            }
            if (BC) {
                bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, slx, srx);
                bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, sly, sry);
                bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, slz, srz);
            }
            const double shx = upw(px_pos, px_nz0, slow, slx, srx);  // QUIRK: |u| > 0.0 (:746)
            const double shy = upw(py_pos, py_nz0, slow, sly, sry);
            const double shz = upw(pz_pos, pz_nz0, slow, slz, srz);
            X_SY[tid] = shy;
            __syncthreads();  // S3

            // ---- P3: transverse differences (dt/6 form) ---------------------------------
            const double usx = up + um, usy = vp + vm;
            const double t6x = (c6x * usx) * (shfl_dn1(shx) - shx);
            const double t6y = (c6y * usy) * (X_SY[tid + GX] - shy);
            const double wsum_p = w_0 + w_m1;  // faces a and a-1: cell a-1
            const double tz = (c6z * wsum_p) * (shz - shz_prev);
            X_T6X[tid] = t6x;
            X_TZ[tid] = tz;
            __syncthreads();  // S4

            // ---- P4: corner-coupled states ----------------------------------------------
            const double uf_p = HS(H_UF), vf_p = HS(H_VF);
            const bool qx_pos = uf_p > 0.0, qx_nz = fabs(uf_p) > eps;
            const bool qy_pos = vf_p > 0.0, qy_nz = fabs(vf_p) > eps;
            double hxy, hyx, hzx, hzy, hxz, hyz;
            {
                double l = slx - shfl_up1(t6y), r = srx - t6y;
                if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, l, r);
                hxy = upw(px_pos, px_nz, slow, l, r);
                l = sly - X_T6X[tid - GX];
                r = sry - t6x;
                if (BC) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, l, r);
                hyx = upw(py_pos, py_nz, slow, l, r);
                l = slz - t6x_prev;
                r = srz - t6x;
                if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, l, r);
                hzx = upw(pz_pos, pz_nz, slow, l, r);
                l = slz - t6y_prev;
                r = srz - t6y;
                if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, l, r);
                hzy = upw(pz_pos, pz_nz, slow, l, r);
                l = HS(H_SLX) - shfl_up1(tz);
                r = HS(H_SRX) - tz;
                if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, l, r);
                hxz = upw(qx_pos, qx_nz, slow, l, r);
                l = HS(H_SLY) - X_TZ[tid - GX];
                r = HS(H_SRY) - tz;
                if (BC) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, l, r);
                hyz = upw(qy_pos, qy_nz, slow, l, r);
            }
            X_HYX[tid] = hyx;
            X_HYZ[tid] = hyz;
            __syncthreads();  // S5

            // ---- P5: transverse differences (dt/4 form) ---------------------------------
            const double c4xu = c4x * usx, c4yu = c4y * usy, c4zw = c4z * wsum_p;
            const double t4xy = c4xu * (shfl_dn1(hxy) - hxy);
            const double t4yx = c4yu * (X_HYX[tid + GX] - hyx);
            const double t4yz = HS(H_C4Y) * (X_HYZ[tid + GX] - hyz);
            const double t4zy = c4zw * (hzy - HS(H_HZY));
            const double t4xz = HS(H_C4X) * (shfl_dn1(hxz) - hxz);
            const double t4zx = c4zw * (hzx - HS(H_HZX));
            const double fs_cur = A.dt2 * f0;
            const double fs_prev = HS(H_FS);
            X_QY1[tid] = t4xz;
            X_QY2[tid] = t4zx;
            X_FS[tid] = fs_prev;
            __syncthreads();  // S6

            // ---- P6: final edge states (+ fluxes) ---------------------------------------
            double ex_v, ey_v, ez_v;
            {
                const double slx_p = HS(H_SLX), srx_p = HS(H_SRX);
                const double el = ((slx_p - shfl_up1(t4yz)) - shfl_up1(t4zy)) + shfl_up1(fs_prev);
                const double er = ((srx_p - t4yz) - t4zy) + fs_prev;
                ex_v = upw(qx_pos, qx_nz, slow, el, er);
                if (BC) ex_v = bc_final(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, el, er, ex_v);
            }
            {
                const double sly_p = HS(H_SLY), sry_p = HS(H_SRY);
                const double el = ((sly_p - X_QY1[tid - GX]) - X_QY2[tid - GX]) + X_FS[tid - GX];
                const double er = ((sry_p - t4xz) - t4zx) + fs_prev;
                ey_v = upw(qy_pos, qy_nz, slow, el, er);
                if (BC) ey_v = bc_final(j, g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, el, er, ey_v);
            }
            {
                const double el = ((slz - HS(H_T4XY)) - HS(H_T4YX)) + fs_prev;
                const double er = ((srz - t4xy) - t4yx) + fs_cur;
                ez_v = upw(pz_pos, pz_nz, slow, el, er);
                if (BC) ez_v = bc_final(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, el, er, ez_v);
            }
            const int eox = ex_co + p * ex_sz, eoy = ey_co + p * ey_sz, eoz = ez_co + a * ez_sz;
            if (wx) Ex[eox] = ex_v;
            if (wy) Ey[eoy] = ey_v;
            if (wz) Ez[eoz] = ez_v;
            if (MODE == 1 && flux1) {
                // MakeRhoXFlux, MaestroMakeFlux.cpp:191-306 (planar) / :316-407 (spherical)
                const int fox = fx_co + p * fx_sz, foy = fy_co + p * fy_sz, foz = fz_co + a * fz_sz;
                double* __restrict__ Fx = A.sflux[0].comp(comp);
                double* __restrict__ Fy = A.sflux[1].comp(comp);
                double* __restrict__ Fz = A.sflux[2].comp(comp);
                double* __restrict__ Gx = A.sflux[0].comp(F_.rho);
                double* __restrict__ Gy = A.sflux[1].comp(F_.rho);
                double* __restrict__ Gz = A.sflux[2].comp(F_.rho);
                const int spt = F_.spt;
                if (wx) {
                    const double ser = need_ser ? Rx[eox] : 0.0;
                    const double gold = first_spec ? 0.0 : Gx[fox];
                    const double r0e = sph ? 0.5 * A.r0o[0].p[A.r0o[0].off(i, j, p)] +
                                                 0.5 * A.r0n[0].p[A.r0n[0].off(i, j, p)]
                                           : T_RXY[it - 1];
                    const double fl = rhoX_flux(spt, uf_p, r0e, ser, ex_v);
                    Fx[fox] = fl;
                    Gx[fox] = gold + fl;
                }
                if (wy) {
                    const double ser = need_ser ? Ry[eoy] : 0.0;
                    const double gold = first_spec ? 0.0 : Gy[foy];
                    const double r0e = sph ? 0.5 * A.r0o[1].p[A.r0o[1].off(i, j, p)] +
                                                 0.5 * A.r0n[1].p[A.r0n[1].off(i, j, p)]
                                           : T_RXY[it - 1];
                    const double fl = rhoX_flux(spt, vf_p, r0e, ser, ey_v);
                    Fy[foy] = fl;
                    Gy[foy] = gold + fl;
                }
                if (wz) {
                    const double ser = need_ser ? Rz[eoz] : 0.0;
                    const double gold = first_spec ? 0.0 : Gz[foz];
                    const double r0e = sph ? 0.5 * A.r0o[2].p[A.r0o[2].off(i, j, a)] +
                                                 0.5 * A.r0n[2].p[A.r0n[2].off(i, j, a)]
                                           : T_RZ[it];
                    const double fl = rhoX_flux(spt, w_0, r0e, ser, ez_v);
                    Fz[foz] = fl;
                    Gz[foz] = gold + fl;
                    if (A.do_eta) {
                        // MaestroMakeFlux.cpp:288-301
                        double* ep = A.eta.p + (eta_co + a * eta_sz);
                        double ev = *ep;
                        if (is_spec) ev += fl;
                        if (last_spec) ev -= T_W0R[it];
                        *ep = ev;
                    }
                }
            } else if (MODE == 2) {
                // MakeRhoHFlux (planar), MaestroMakeFlux.cpp:674-786
                const int fox = fx_co + p * fx_sz, foy = fy_co + p * fy_sz, foz = fz_co + a * fz_sz;
                double* __restrict__ Fx = A.sflux[0].comp(comp);
                double* __restrict__ Fy = A.sflux[1].comp(comp);
                double* __restrict__ Fz = A.sflux[2].comp(comp);
                double fl;
                if (wx) {
                    const double ser = need_ser ? Rx[eox] : 0.0;
                    if (rhoh_flux_tab(F_, have_h, uf_p, ser, ex_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                        Fx[fox] = fl;
                }
                if (wy) {
                    const double ser = need_ser ? Ry[eoy] : 0.0;
                    if (rhoh_flux_tab(F_, have_h, vf_p, ser, ey_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                        Fy[foy] = fl;
                }
                if (wz) {
                    const double ser = need_ser ? Rz[eoz] : 0.0;
                    if (rhoh_flux_tab(F_, have_h, w_0, ser, ez_v, T_RZ[it], T_HZ[it], true, &fl))
                        Fz[foz] = fl;
                }
            }

            // ---- z history for the next plane -----------------------------------------------
            ipz_prev = ipz;
            shz_prev = shz;
            t6x_prev = t6x;
            t6y_prev = t6y;
            HS(H_SLX) = slx;
            HS(H_SRX) = srx;
            HS(H_SLY) = sly;
            HS(H_SRY) = sry;
            HS(H_UF) = um;
            HS(H_VF) = vm;
            HS(H_C4X) = c4xu;
            HS(H_C4Y) = c4yu;
            HS(H_HZX) = hzx;
            HS(H_HZY) = hzy;
            HS(H_T4XY) = t4xy;
            HS(H_T4YX) = t4yx;
            HS(H_FS) = fs_cur;
            slow_p = slow_c;
        }
        cp_async_wait<0>();
    }
#undef HS
}

template <int PPM, int MODE>
__global__ void __launch_bounds__(512, 1) k_stream(const __grid_constant__ FusedArgs A) {
    extern __shared__ double sm[];
    // physical-boundary code only in blocks whose (grown) tile reaches a domain face that has
    // a non-interior BC; everything else runs the BC-free instantiation
    constexpr int BX = 30, BY = 14;
    const KGeom& g = A.g;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    const int x0 = g.vlo[0] + bx * BX, y0 = g.vlo[1] + by * BY;
    const int k0 = g.vlo[2] + blockIdx.y * A.zc;
    bool bc = false;
    if (A.bcdir[0]) bc |= (x0 - 1 <= g.dlo[0] + 2) || (x0 + BX >= g.dhi[0] - 2);
    if (A.bcdir[1]) bc |= (y0 - 1 <= g.dlo[1] + 2) || (y0 + BY >= g.dhi[1] - 2);
    if (A.bcdir[2]) bc |= (k0 - 1 <= g.dlo[2] + 2) || (k0 + A.zc >= g.dhi[2] - 2);
    if (bc)
        stream_body<PPM, true, MODE>(A, sm);
    else
        stream_body<PPM, false, MODE>(A, sm);
}

template <int PPM, int MODE>
int launch_stream(const FusedArgs& A, cudaStream_t stream) {
    const size_t smem = sizeof(double) * (size_t)stream_smem_doubles<PPM>();
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_stream<PPM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    const int nz = A.g.vhi[2] - A.g.vlo[2] + 1;
    dim3 grid(A.ntx * A.nty, (nz + A.zc - 1) / A.zc);
    k_stream<PPM, MODE><<<grid, 512, smem, stream>>>(A);
    mx_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int PPM>
int launch_stream_mode(const FusedArgs& A, cudaStream_t stream) {
    if (A.mode == 0) return launch_stream<PPM, 0>(A, stream);
    if (A.mode == 1) return launch_stream<PPM, 1>(A, stream);
    return launch_stream<PPM, 2>(A, stream);
}

#endif
