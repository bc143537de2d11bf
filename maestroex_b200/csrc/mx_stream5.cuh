// mx_stream5.cuh — k_stream5: the streaming edge-state (+flux) kernel, fifth revision.
// Same pipeline, phases and arithmetic as k_stream4 (mx_stream4.cuh); what changed is the
// thread -> cell map: a block of 32 x 8 threads owns the same 32 x 16 (30 x 14 valid) tile,
// each thread carrying TWO vertically adjacent cells (rows 2*ty and 2*ty+1) through the
// plane pipeline.  The round-1b capture (profiles/r1b_k_stream4_ncu_full_n256.txt) showed
// k_stream4 issue-bound with ~1200 SASS instructions per cell-plane of which only ~390 are
// fp64 arithmetic; the rest is addressing, loop control, constant reloads, register moves and
// shared-memory exchange.  Two cells per thread
//   * amortise every running pointer, uniform-datapath op, branch and barrier over two cells,
//   * share the y reconstruction between the two rows in registers (PPM 1: four van Leer
//     slopes and three edge values per pair instead of six and four; PPM 2: five limited
//     edges instead of eight),
//   * halve the shared-memory y-exchange (only the value the neighbouring *thread* needs is
//     published) and move the thread-private z history as 16-byte double2 slots,
//   * give the scheduler two independent dependency chains per warp (255 registers/thread).
#ifndef MX_STREAM5_CUH
#define MX_STREAM5_CUH

constexpr int S5_NS = 8;    // ring slots (power of two)
constexpr int S5_NT = 256;  // threads per block

template <int PPM>
constexpr int stream5_smem_doubles() {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    return S5_NS * (32 + 2 * LA) * (16 + 2 * LA) + 32 + 9 * S5_NT + 32 + 2 * 16 * S5_NT +
           5 * (ST_MAXZC + 2);
}

// streaming (evict-first) global accesses for data touched once per pass, so that the rho'
// edge / rho flux lines a block re-reads for every species stay L2 resident
MXD double ldg_stream(const double* p) { return __ldcs(p); }
MXD void stg_stream(double* p, double v) { __stcs(p, v); }

template <int PPM, bool BC, int MODE>
__device__ __forceinline__ void stream5_body(const FusedArgs& A, double* smem) {
    constexpr int GX = 32, GY = 16, NT = S5_NT;
    constexpr int LA = (PPM == 2) ? 3 : 2;  // stencil radius
    constexpr int PW = GX + 2 * LA, PH = GY + 2 * LA;
    constexpr int PL = PW * PH;  // doubles per ring plane
    constexpr int BX = GX - 2, BY = GY - 2;
    constexpr int NLD = (PL + NT - 1) / NT;

    double* ring = smem;
    double* ex = ring + S5_NS * PL + 32;  // y-neighbour exchange arrays (32 doubles of slack
                                          // on either side for the ty-1 / ty+1 reads)
    double* X_IPY = ex + 0 * NT;
    double* X_SY = ex + 1 * NT;
    double* X_T6X = ex + 2 * NT;
    double* X_TZ = ex + 3 * NT;
    double* X_HYX = ex + 4 * NT;
    double* X_HYZ = ex + 5 * NT;
    double* X_QY1 = ex + 6 * NT;
    double* X_QY2 = ex + 7 * NT;
    double* X_FS = ex + 8 * NT;
    double2* hist = reinterpret_cast<double2*>(ex + 9 * NT + 32);  // thread-private z history
    double* tab = reinterpret_cast<double*>(hist + 16 * NT);
    double* T_RXY = tab;                       // 0.5 (rho0_old + rho0_new)(p)          x/y faces
    double* T_RZ = tab + (ST_MAXZC + 2);       // 0.5 (rho0_edge_old + _new)(a) z faces
    double* T_W0R = tab + 2 * (ST_MAXZC + 2);  // w0(a) * rho0_predicted_edge(a)
    double* T_HXY = tab + 3 * (ST_MAXZC + 2);  // 0.5 (rhoh0_old + rhoh0_new)(p)
    double* T_HZ = tab + 4 * (ST_MAXZC + 2);   // 0.5 (rhoh0_edge_old + _new)(a)

    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    double2* const hs = hist + tid;
    enum { H_SLX, H_SRX, H_SLY, H_SRY, H_UF, H_VF, H_C4X, H_C4Y, H_HZX, H_HZY, H_T4XY, H_T4YX,
           H_FS, H_T6X, H_T6Y, H_SHZ };
#define HLD(v, dst)                        \
    {                                      \
        const double2 t__ = hs[(v) * NT];  \
        dst[0] = t__.x;                    \
        dst[1] = t__.y;                    \
    }
#define HST(v, src) hs[(v) * NT] = make_double2(src[0], src[1])

    const KGeom& g = A.g;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    const int x0 = g.vlo[0] + bx * BX, y0 = g.vlo[1] + by * BY;
    const int thx = min(x0 + BX - 1, g.vhi[0]), thy = min(y0 + BY - 1, g.vhi[1]);
    const int i = x0 - 1 + tx;
    const int jj[2] = {y0 - 1 + 2 * ty, y0 + 2 * ty};
    int k0, k1;  // this block's z chunk, k1 exclusive
    chunk_range(A, blockIdx.y, k0, k1);
    const bool last = (k1 == g.vhi[2] + 1);
    const int a0 = k0 - 1;

    // faces this thread stores: the tile's low faces, plus the high face only on the last tile
    bool ownx[2], owny[2], ownz[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int j = jj[c];
        const bool inG = (i <= thx + 1) && (j <= thy + 1);
        const bool vxi = inG && tx >= 1 && i <= thx;
        const bool vyj = inG && (2 * ty + c) >= 1 && j <= thy;
        ownx[c] = inG && tx >= 1 && vyj && (i <= thx || thx == g.vhi[0]);
        owny[c] = inG && (2 * ty + c) >= 1 && vxi && (j <= thy || thy == g.vhi[1]);
        ownz[c] = vxi && vyj;
    }
    // clamped columns (see k_stream4): every lane addresses an element that exists
    const int il = min(i, thx + 1);
    const int icx = min(max(i, x0), thx + 1), icy = min(max(i, x0), thx), icz = icy;
    int jl[2], jcx[2], jcy[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        jl[c] = min(jj[c], thy + 1);
        jcx[c] = min(max(jj[c], y0), thy);      // x and z faces
        jcy[c] = min(max(jj[c], y0), thy + 1);  // y faces
    }
    const int rc = (2 * ty + LA) * PW + (tx + LA);  // ring offset of cell 0; cell 1 is + PW

    const DFab& st = A.state;
    const long st_sz = st.sz, f_sz = A.force.sz;
    const long u_sz = A.umac[0].sz, v_sz = A.umac[1].sz, w_sz = A.umac[2].sz;
    const int v_sy = (int)A.umac[1].sy;
    const long ex_sz = A.sedge[0].sz, ey_sz = A.sedge[1].sz, ez_sz = A.sedge[2].sz;
    const long fx_sz = A.sflux[0].sz, fy_sz = A.sflux[1].sz, fz_sz = A.sflux[2].sz;
    const long eta_sz = A.eta.sz;
    const KFlux& F_ = A.f;
    const bool sph = (MODE != 0) && F_.spherical;
    const bool lane31 = (tx == 31);

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
        ld_do[q] = e < PL;
        ld_in[q] = ld_do[q] && gi >= 0 && gi < st.n[0] && gj >= 0 && gj < st.n[1];
        ld_off[q] = ld_in[q] ? (int)(gj * st.sy + gi) : 0;
    }

    const double c6x = A.c6h[0], c6y = A.c6h[1], c6z = A.c6h[2];
    const double c4x = A.c4h[0], c4y = A.c4h[1], c4z = A.c4h[2];
    const double eps = A.eps, dt = A.dt;
    const bool eps_neg = !(eps >= 0.0);
    const bool have_h = (MODE == 2) && (F_.ept == MX_PRED_H || F_.ept == MX_PRED_T_THEN_H ||
                                        F_.ept == MX_PRED_TPRIME_THEN_H);
    const bool need_ser = (MODE == 1) ? (F_.spt != MX_PRED_RHOX) : have_h;

    // column offsets of cell 0 (component independent) and the cell-1 deltas (0 or one row)
    const long u_c = A.umac[0].off(il, jl[0], a0), v_c = A.umac[1].off(il, jl[0], a0);
    const long w_c = A.umac[2].off(il, jl[0], a0);
    const long f_c = A.force.off(il, jl[0], a0);
    const long ex_c = A.sedge[0].off(icx, jcx[0], a0 - 1), ey_c = A.sedge[1].off(icy, jcy[0], a0 - 1);
    const long ez_c = A.sedge[2].off(icz, jcx[0], a0);
    const int dU1 = (jl[1] - jl[0]) * (int)A.umac[0].sy, dV1 = (jl[1] - jl[0]) * v_sy;
    const int dW1 = (jl[1] - jl[0]) * (int)A.umac[2].sy, dF1 = (jl[1] - jl[0]) * (int)A.force.sy;
    const int dEx1 = (jcx[1] - jcx[0]) * (int)A.sedge[0].sy;
    const int dEy1 = (jcy[1] - jcy[0]) * (int)A.sedge[1].sy;
    const int dEz1 = (jcx[1] - jcx[0]) * (int)A.sedge[2].sy;
    long fx_c = 0, fy_c = 0, fz_c = 0, eta_c = 0;
    int dFx1 = 0, dFy1 = 0, dFz1 = 0, dEta1 = 0;
    if (MODE != 0) {
        fx_c = A.sflux[0].off(icx, jcx[0], a0 - 1);
        fy_c = A.sflux[1].off(icy, jcy[0], a0 - 1);
        fz_c = A.sflux[2].off(icz, jcx[0], a0);
        dFx1 = (jcx[1] - jcx[0]) * (int)A.sflux[0].sy;
        dFy1 = (jcy[1] - jcy[0]) * (int)A.sflux[1].sy;
        dFz1 = (jcx[1] - jcx[0]) * (int)A.sflux[2].sy;
        if (MODE == 1 && A.do_eta) {
            eta_c = A.eta.off(icz, jcx[0], a0);
            dEta1 = (jcx[1] - jcx[0]) * (int)A.eta.sy;
        }
    }

    for (int n = 0; n < A.ncomp; ++n) {
        const int comp = A.comp[n];
        const int bclx = A.bclo[n][0], bchx = A.bchi[n][0];
        const int bcly = A.bclo[n][1], bchy = A.bchi[n][1];
        const int bclz = A.bclo[n][2], bchz = A.bchi[n][2];
        const bool lobcx = BC && (bclx == BC_EXT_DIR || bclx == BC_HOEXTRAP);
        const bool hibcx = BC && (bchx == BC_EXT_DIR || bchx == BC_HOEXTRAP);
        const bool lobcy = BC && (bcly == BC_EXT_DIR || bcly == BC_HOEXTRAP);
        const bool hibcy = BC && (bchy == BC_EXT_DIR || bchy == BC_HOEXTRAP);
        const bool lobcz = BC && (bclz == BC_EXT_DIR || bclz == BC_HOEXTRAP);
        const bool hibcz = BC && (bchz == BC_EXT_DIR || bchz == BC_HOEXTRAP);
        const bool flux1 = (MODE == 1 && n > 0);
        const bool first_spec = (n == 1);
        const bool is_spec =
            comp >= F_.first_spec && comp <= F_.first_spec + F_.num_spec - 1;  // eta: :288-301
        const bool last_spec = comp == F_.first_spec + F_.num_spec - 1;

        // ---- running pointers (cell 0): they address the plane this iteration touches ----
        const double* __restrict__ pU = A.umac[0].p + u_c;         // plane a
        const double* __restrict__ pV = A.umac[1].p + v_c;         // plane a
        const double* __restrict__ pW = A.umac[2].p + w_c + w_sz;  // face a+1
        const double* __restrict__ pF = A.force.comp(comp) + f_c;  // plane a
        double* __restrict__ pEx = A.sedge[0].comp(comp) + ex_c;   // plane p = a-1
        double* __restrict__ pEy = A.sedge[1].comp(comp) + ey_c;
        double* __restrict__ pEz = A.sedge[2].comp(comp) + ez_c;   // face a
        double* __restrict__ pFx = nullptr;
        double* __restrict__ pFy = nullptr;
        double* __restrict__ pFz = nullptr;
        double* __restrict__ pEta = nullptr;
        long dRx = 0, dRy = 0, dRz = 0, dGx = 0, dGy = 0, dGz = 0;  // rho comp - this comp
        long dQx = 0, dQy = 0, dQz = 0;  // spherical: r0e - this comp of sedge
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
        }
        const double* __restrict__ pS[NLD];  // source of ring element q, plane being staged
        {
            const double* sp_ = st.comp(comp) + (long)(a0 - LA - st.lo[2]) * st_sz;
#pragma unroll
            for (int q = 0; q < NLD; ++q) pS[q] = sp_ + ld_off[q];
        }
        auto stage_plane = [&](int slot) {
            double* dst = ring + slot * PL + tid;
#pragma unroll
            for (int q = 0; q < NLD; ++q) {
                if (ld_do[q]) cp_async8(dst + q * NT, pS[q], ld_in[q]);
                pS[q] += st_sz;
            }
        };

        // ---- prologue ---------------------------------------------------------------
        __syncthreads();  // previous component is done with shared memory; tables written
#pragma unroll
        for (int q = 0; q <= 2 * LA; ++q) stage_plane(q);  // planes a0-LA .. a0+LA
        cp_async_commit();
        // register prefetch of plane a0
        double w_m1[2] = {0.0, 0.0}, w_0[2] = {0.0, 0.0}, w_p1[2], w_n[2];
        double um_n[2], up31_n[2], f_n[2];
        double vf_n[3];  // y faces jl0, jl1, jl1+1
        w_p1[0] = pW[-w_sz];
        w_p1[1] = pW[dW1 - w_sz];  // face a0
        w_n[0] = pW[0];
        w_n[1] = pW[dW1];  // face a0+1
        um_n[0] = pU[0];
        um_n[1] = pU[dU1];
        up31_n[0] = lane31 ? pU[1] : 0.0;
        up31_n[1] = lane31 ? pU[dU1 + 1] : 0.0;
        vf_n[0] = pV[0];
        vf_n[1] = pV[dV1];
        vf_n[2] = pV[dV1 + v_sy];
        f_n[0] = ldg_stream(pF);
        f_n[1] = ldg_stream(pF + dF1);
        double sgz_hi[2], ipz_prev[2] = {0.0, 0.0};
        double dz_c[2] = {0.0, 0.0}, ez_lo[2] = {0.0, 0.0};                              // PPM 1
        double ez_l2[2] = {0.0, 0.0}, ez_l1[2] = {0.0, 0.0}, ez_h1[2] = {0.0, 0.0};      // PPM 2
#pragma unroll
        for (int c = 0; c < 2; ++c)
            sgz_hi[c] = (PPM != 0) ? div_h<2>(A, fabs(w_p1[c]) * dt) : 0.0;  // sigma of face a0
        {
            const double z2[2] = {0.0, 0.0};
#pragma unroll
            for (int v = 0; v < 16; ++v) HST(v, z2);
        }
        bool slow_p = true;
        cp_async_wait<0>();
        __syncthreads();  // planes a0-LA .. a0+LA are in the ring
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            // incremental z reconstruction state from planes a0-2 .. a0+1
            const double* cc = ring + rc + c * PW;
            if (PPM == 1) {
                const double zm2 = cc[(LA - 2) * PL], zm1 = cc[(LA - 1) * PL], z0 = cc[LA * PL],
                             zp1 = cc[(LA + 1) * PL];
                const double dm1 = vanleer(zm2, zm1, z0);
                dz_c[c] = vanleer(zm1, z0, zp1);
                ez_lo[c] = vl_edge(zm1, z0, dm1, dz_c[c]);
            } else if (PPM == 2) {
                const double z[6] = {cc[0], cc[PL], cc[2 * PL], cc[3 * PL], cc[4 * PL], cc[5 * PL]};
                ez_l2[c] = cs_edge(z[0], z[1], z[2], z[3]);
                ez_l1[c] = cs_edge(z[1], z[2], z[3], z[4]);
                ez_h1[c] = cs_edge(z[2], z[3], z[4], z[5]);
            }
        }

        for (int a = a0; a <= k1; ++a) {
            const int it = a - a0;
            const double* pa = ring + ((it + LA) & (S5_NS - 1)) * PL + rc;  // plane a, cell 0
            // ---- P0: current-plane registers, prefetch of the next plane ----------------
            double um[2], up[2], f0[2];
            const double vf[3] = {vf_n[0], vf_n[1], vf_n[2]};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                um[c] = um_n[c];
                f0[c] = f_n[c];
                up[c] = shfl_dn1(um[c]);
                if (lane31) up[c] = up31_n[c];
                w_m1[c] = w_0[c];
                w_0[c] = w_p1[c];
                w_p1[c] = w_n[c];
            }
            if (a < k1) {  // block uniform
                stage_plane((it + 2 * LA + 1) & (S5_NS - 1));  // plane a+LA+1
                pU += u_sz;
                pV += v_sz;
                pW += w_sz;
                pF += f_sz;
                um_n[0] = pU[0];
                um_n[1] = pU[dU1];
                if (lane31) {
                    up31_n[0] = pU[1];
                    up31_n[1] = pU[dU1 + 1];
                }
                vf_n[0] = pV[0];
                vf_n[1] = pV[dV1];
                vf_n[2] = pV[dV1 + v_sy];
                w_n[0] = pW[0];
                w_n[1] = pW[dW1];
                f_n[0] = ldg_stream(pF);
                f_n[1] = ldg_stream(pF + dF1);
            }
            cp_async_commit();
            // velocity predicates of this plane's low faces
            bool px_pos[2], px_nz[2], py_pos[2], py_nz[2], pz_pos[2], pz_nz[2];
            bool all_nz = true;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                px_pos[c] = um[c] > 0.0;
                px_nz[c] = fabs(um[c]) > eps;
                py_pos[c] = vf[c] > 0.0;
                py_nz[c] = fabs(vf[c]) > eps;
                pz_pos[c] = w_0[c] > 0.0;
                pz_nz[c] = fabs(w_0[c]) > eps;
                all_nz = all_nz && px_nz[c] && py_nz[c] && pz_nz[c];
            }
            // QUIRK: the first Riemann solve tests |u| > 0.0 (:746) where the later ones test
            // rel_eps; for eps >= 0, |u| > eps implies |u| > 0, so the fast path (no averaging
            // arm anywhere in the warp) only needs the eps tests; eps < 0 always goes slow
            const bool slow_c = __any_sync(0xffffffffu, !all_nz) || eps_neg;
            const bool slow = slow_c || slow_p;

            // ---- P1: x / y reconstruction of plane a ------------------------------------
            double s0[2], ipx[2], imx[2], ipy[2], imy[2];
            s0[0] = pa[0];
            s0[1] = pa[PW];
            {
                // x: each cell its own five/seven-point row
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    double S[7];
                    S[0] = S[6] = 0.0;
#pragma unroll
                    for (int q = -LA; q <= LA; ++q) S[q + 3] = (q == 0) ? s0[c] : pa[c * PW + q];
                    if (PPM == 0) {
                        ipx[c] = slope_any(A.slope_order, S + 1, i, g.dlo[0], g.dhi[0], lobcx, hibcx, true);
                        imx[c] = 0.0;
                    } else {
                        double smx, spx;
                        if (PPM == 1)
                            ppm1_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, smx, spx);
                        else
                            ppm2_parabola<BC>(S, i, g.dlo[0], g.dhi[0], lobcx, hibcx, true, smx, spx);
                        const double sgx_m = div_h<0>(A, fabs(um[c]) * dt);
                        const double sgx_p = shfl_dn1(sgx_m);  // lane 31 never needs Ip
                        const double s6x = 6.0 * s0[c] - 3.0 * (smx + spx);
                        ipx[c] = ppm_Ip(smx, spx, s0[c], s6x, up[c], sgx_p, eps);
                        imx[c] = ppm_Im(smx, spx, s0[c], s6x, um[c], sgx_m, eps);
                    }
                }
                // y: the column rows -LA .. LA+1 relative to cell 0 serve both cells
                double Y[2 * LA + 2];
#pragma unroll
                for (int q = -LA; q <= LA + 1; ++q)
                    Y[q + LA] = (q == 0) ? s0[0] : (q == 1) ? s0[1] : pa[q * PW];
                if (PPM == 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        double T[5];
#pragma unroll
                        for (int q = 0; q < 5; ++q) T[q] = Y[q + c];
                        ipy[c] = slope_any(A.slope_order, T, jj[c], g.dlo[1], g.dhi[1], lobcy, hibcy, false);
                        imy[c] = 0.0;
                    }
                } else {
                    double smy[2], spy[2];
                    if (PPM == 1) {
                        // van Leer slopes of rows -1 .. 2, edges -1/2, 1/2, 3/2
                        double d[4], E[3];
#pragma unroll
                        for (int q = 0; q < 4; ++q) d[q] = vanleer(Y[q], Y[q + 1], Y[q + 2]);
#pragma unroll
                        for (int q = 0; q < 3; ++q) E[q] = vl_edge(Y[q + 1], Y[q + 2], d[q], d[q + 1]);
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            double T[7];
                            T[0] = T[6] = 0.0;
#pragma unroll
                            for (int q = 0; q < 5; ++q) T[q + 1] = Y[q + c];
                            ppm1_finish<BC>(T, jj[c], g.dlo[1], g.dhi[1], lobcy, hibcy, E[c], E[c + 1],
                                            smy[c], spy[c]);
                        }
                    } else {
                        // limited edges -3/2 .. 5/2 from rows -3 .. 4
                        double E[5];
#pragma unroll
                        for (int q = 0; q < 5; ++q) E[q] = cs_edge(Y[q], Y[q + 1], Y[q + 2], Y[q + 3]);
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            double T[7];
#pragma unroll
                            for (int q = 0; q < 7; ++q) T[q] = Y[q + c];
                            ppm2_finish<BC>(T, jj[c], g.dlo[1], g.dhi[1], lobcy, hibcy, false, E[c],
                                            E[c + 1], E[c + 2], E[c + 3], smy[c], spy[c]);
                        }
                    }
                    double sgy[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) sgy[q] = div_h<1>(A, fabs(vf[q]) * dt);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const double s6y = 6.0 * s0[c] - 3.0 * (smy[c] + spy[c]);
                        ipy[c] = ppm_Ip(smy[c], spy[c], s0[c], s6y, vf[c + 1], sgy[c + 1], eps);
                        imy[c] = ppm_Im(smy[c], spy[c], s0[c], s6y, vf[c], sgy[c], eps);
                    }
                }
            }
            X_IPY[tid] = ipy[1];
            cp_async_wait<1>();  // this thread's share of plane a+LA has landed
            __syncthreads();     // S2: publishes IPY and ring plane a+LA

            // ---- P2: z reconstruction, 1-D states, first Riemann solve -------------------
            double ipz[2], imz[2], s_zm1[2] = {0.0, 0.0};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double S[7];
                S[0] = S[6] = 0.0;
                // planes a-LA .. a+LA; without BC rows PPM 1 only reads a .. a+2
#pragma unroll
                for (int q = -LA; q <= LA; ++q) {
                    const bool need = BC || PPM != 1 || q > 0;
                    S[q + 3] = (q == 0) ? s0[c]
                               : need   ? ring[((it + LA + q) & (S5_NS - 1)) * PL + rc + c * PW]
                                        : 0.0;
                }
                if (PPM == 0) {
                    ipz[c] = slope_any(A.slope_order, S + 1, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false);
                    imz[c] = 0.0;
                    s_zm1[c] = S[2];
                } else {
                    double sm, sp;
                    if (PPM == 1) {
                        const double dz_n = vanleer(S[3], S[4], S[5]);
                        const double ez_hi = vl_edge(S[3], S[4], dz_c[c], dz_n);
                        ppm1_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, ez_lo[c], ez_hi, sm, sp);
                        dz_c[c] = dz_n;
                        ez_lo[c] = ez_hi;
                    } else {
                        const double ez_h2 = cs_edge(S[3], S[4], S[5], S[6]);
                        ppm2_finish<BC>(S, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false, ez_l2[c],
                                        ez_l1[c], ez_h1[c], ez_h2, sm, sp);
                        ez_l2[c] = ez_l1[c];
                        ez_l1[c] = ez_h1[c];
                        ez_h1[c] = ez_h2;
                    }
                    const double s6 = 6.0 * s0[c] - 3.0 * (sm + sp);
                    const double sgz_lo = sgz_hi[c];  // face a: last plane's high face
                    sgz_hi[c] = div_h<2>(A, fabs(w_p1[c]) * dt);
                    ipz[c] = ppm_Ip(sm, sp, s0[c], s6, w_p1[c], sgz_hi[c], eps);
                    imz[c] = ppm_Im(sm, sp, s0[c], s6, w_0[c], sgz_lo, eps);
                }
            }
            double slx[2], srx[2], sly[2], sry[2], slz[2], srz[2];
            {
                const double ipy_below = X_IPY[tid - GX];  // cell 1 of the thread below
                if (PPM == 0) {
                    // MaestroMakeEdgeScalars.cpp:697-703
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const double qx = div_h<0>(A, dt * um[c]), qy = div_h<1>(A, dt * vf[c]),
                                     qz = div_h<2>(A, dt * w_0[c]);
                        slx[c] = pa[c * PW - 1] + 0.5 * (1.0 - qx) * shfl_up1(ipx[c]);
                        srx[c] = s0[c] - 0.5 * (1.0 + qx) * ipx[c];
                        sly[c] = (c == 0 ? pa[-PW] : s0[0]) +
                                 0.5 * (1.0 - qy) * (c == 0 ? ipy_below : ipy[0]);
                        sry[c] = s0[c] - 0.5 * (1.0 + qy) * ipy[c];
                        slz[c] = s_zm1[c] + 0.5 * (1.0 - qz) * ipz_prev[c];
                        srz[c] = s0[c] - 0.5 * (1.0 + qz) * ipz[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        slx[c] = shfl_up1(ipx[c]);
                        srx[c] = imx[c];
                        sly[c] = (c == 0) ? ipy_below : ipy[0];
                        sry[c] = imy[c];
                        slz[c] = ipz_prev[c];
                        srz[c] = imz[c];
                    }
                }
            }
            double shx[2], shy[2], shz[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                ipz_prev[c] = ipz[c];
                if (BC) {
                    bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, slx[c], srx[c]);
                    bc_pair(jj[c], g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, sly[c], sry[c]);
                    bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, slz[c], srz[c]);
                }
                // QUIRK: |u| > 0.0 here (:746), rel_eps in the later solves
                shx[c] = upw(px_pos[c], fabs(um[c]) > 0.0, slow, slx[c], srx[c]);
                shy[c] = upw(py_pos[c], fabs(vf[c]) > 0.0, slow, sly[c], sry[c]);
                shz[c] = upw(pz_pos[c], fabs(w_0[c]) > 0.0, slow, slz[c], srz[c]);
            }
            X_SY[tid] = shy[0];
            __syncthreads();  // S3

            // ---- P3: transverse differences (dt/6 form) ---------------------------------
            double usx[2], usy[2], wsum_p[2], t6x[2], t6y[2], tz[2];
            {
                double shz_prev[2];
                HLD(H_SHZ, shz_prev);
                const double shy_above = X_SY[tid + GX];  // cell 0 of the thread above
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    usx[c] = up[c] + um[c];
                    usy[c] = vf[c + 1] + vf[c];
                    wsum_p[c] = w_0[c] + w_m1[c];  // faces a and a-1: cell a-1
                    t6x[c] = (c6x * usx[c]) * (shfl_dn1(shx[c]) - shx[c]);
                    t6y[c] = (c6y * usy[c]) * ((c == 0 ? shy[1] : shy_above) - shy[c]);
                    tz[c] = (c6z * wsum_p[c]) * (shz[c] - shz_prev[c]);
                }
                HST(H_SHZ, shz);
            }
            X_T6X[tid] = t6x[1];
            X_TZ[tid] = tz[1];
            // what P6 of this iteration reads (rho' edges) and read-modify-writes (rho flux,
            // etarhoflux): issue the loads now, three phases ahead of their use
            const bool do_xy = a > k0;                      // x/y faces of plane a-1 exist
            const bool do_z = a >= k0 && (a < k1 || last);  // z face a belongs to this chunk
            double r_x[2] = {0.0, 0.0}, r_y[2] = {0.0, 0.0}, r_z[2] = {0.0, 0.0};
            double g_x[2] = {0.0, 0.0}, g_y[2] = {0.0, 0.0}, g_z[2] = {0.0, 0.0}, e_z[2] = {0.0, 0.0};
            if ((MODE == 1 && flux1) || MODE == 2) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int oEx = c ? dEx1 : 0, oEy = c ? dEy1 : 0, oEz = c ? dEz1 : 0;
                    const int oFx = c ? dFx1 : 0, oFy = c ? dFy1 : 0, oFz = c ? dFz1 : 0;
                    if (need_ser) {
                        if (do_xy) {
                            r_x[c] = pEx[dRx + oEx];
                            r_y[c] = pEy[dRy + oEy];
                        }
                        if (do_z) r_z[c] = pEz[dRz + oEz];
                    }
                    if (MODE == 1 && !first_spec) {
                        if (do_xy) {
                            g_x[c] = pFx[dGx + oFx];
                            g_y[c] = pFy[dGy + oFy];
                        }
                        if (do_z) g_z[c] = pFz[dGz + oFz];
                    }
                    if (MODE == 1 && A.do_eta && do_z) e_z[c] = pEta[c ? dEta1 : 0];
                }
            }
            __syncthreads();  // S4

            // ---- P4: corner-coupled states ----------------------------------------------
            double uf_p[2], vf_p[2], slx_p[2], srx_p[2], sly_p[2], sry_p[2];
            HLD(H_UF, uf_p);
            HLD(H_VF, vf_p);
            HLD(H_SLX, slx_p);
            HLD(H_SRX, srx_p);
            HLD(H_SLY, sly_p);
            HLD(H_SRY, sry_p);
            bool qx_pos[2], qx_nz[2], qy_pos[2], qy_nz[2];
            double hxy[2], hyx[2], hzx[2], hzy[2], hxz[2], hyz[2];
            {
                double t6x_prev[2], t6y_prev[2];
                HLD(H_T6X, t6x_prev);
                HLD(H_T6Y, t6y_prev);
                const double t6x_below = X_T6X[tid - GX], tz_below = X_TZ[tid - GX];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    qx_pos[c] = uf_p[c] > 0.0;
                    qx_nz[c] = fabs(uf_p[c]) > eps;
                    qy_pos[c] = vf_p[c] > 0.0;
                    qy_nz[c] = fabs(vf_p[c]) > eps;
                    double l = slx[c] - shfl_up1(t6y[c]), r = srx[c] - t6y[c];
                    if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, l, r);
                    hxy[c] = upw(px_pos[c], px_nz[c], slow, l, r);
                    l = sly[c] - (c == 0 ? t6x_below : t6x[0]);
                    r = sry[c] - t6x[c];
                    if (BC) bc_pair(jj[c], g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, l, r);
                    hyx[c] = upw(py_pos[c], py_nz[c], slow, l, r);
                    l = slz[c] - t6x_prev[c];
                    r = srz[c] - t6x[c];
                    if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, l, r);
                    hzx[c] = upw(pz_pos[c], pz_nz[c], slow, l, r);
                    l = slz[c] - t6y_prev[c];
                    r = srz[c] - t6y[c];
                    if (BC) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, l, r);
                    hzy[c] = upw(pz_pos[c], pz_nz[c], slow, l, r);
                    l = slx_p[c] - shfl_up1(tz[c]);
                    r = srx_p[c] - tz[c];
                    if (BC) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, l, r);
                    hxz[c] = upw(qx_pos[c], qx_nz[c], slow, l, r);
                    l = sly_p[c] - (c == 0 ? tz_below : tz[0]);
                    r = sry_p[c] - tz[c];
                    if (BC) bc_pair(jj[c], g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, l, r);
                    hyz[c] = upw(qy_pos[c], qy_nz[c], slow, l, r);
                }
            }
            HST(H_T6X, t6x);
            HST(H_T6Y, t6y);
            X_HYX[tid] = hyx[0];
            X_HYZ[tid] = hyz[0];
            __syncthreads();  // S5

            // ---- P5: transverse differences (dt/4 form) ---------------------------------
            double c4xu[2], c4yu[2], t4xy[2], t4yx[2], t4yz[2], t4zy[2], t4xz[2], t4zx[2];
            double fs_cur[2], fs_prev[2];
            {
                double c4x_prev[2], c4y_prev[2], hzx_prev[2], hzy_prev[2];
                HLD(H_C4X, c4x_prev);
                HLD(H_C4Y, c4y_prev);
                HLD(H_HZX, hzx_prev);
                HLD(H_HZY, hzy_prev);
                HLD(H_FS, fs_prev);
                const double hyx_above = X_HYX[tid + GX], hyz_above = X_HYZ[tid + GX];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    c4xu[c] = c4x * usx[c];
                    c4yu[c] = c4y * usy[c];
                    const double c4zw = c4z * wsum_p[c];
                    t4xy[c] = c4xu[c] * (shfl_dn1(hxy[c]) - hxy[c]);
                    t4yx[c] = c4yu[c] * ((c == 0 ? hyx[1] : hyx_above) - hyx[c]);
                    t4yz[c] = c4y_prev[c] * ((c == 0 ? hyz[1] : hyz_above) - hyz[c]);
                    t4zy[c] = c4zw * (hzy[c] - hzy_prev[c]);
                    t4xz[c] = c4x_prev[c] * (shfl_dn1(hxz[c]) - hxz[c]);
                    t4zx[c] = c4zw * (hzx[c] - hzx_prev[c]);
                    fs_cur[c] = A.dt2 * f0[c];
                }
            }
            X_QY1[tid] = t4xz[1];
            X_QY2[tid] = t4zx[1];
            X_FS[tid] = fs_prev[1];
            __syncthreads();  // S6

            // ---- P6: final edge states (+ fluxes) ---------------------------------------
            if (do_xy) {  // block uniform: x / y faces of plane p = a-1
                const double qy1_below = X_QY1[tid - GX], qy2_below = X_QY2[tid - GX],
                             fs_below = X_FS[tid - GX];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int oEx = c ? dEx1 : 0, oEy = c ? dEy1 : 0;
                    const int oFx = c ? dFx1 : 0, oFy = c ? dFy1 : 0;
                    double ex_v, ey_v;
                    {
                        const double el = ((slx_p[c] - shfl_up1(t4yz[c])) - shfl_up1(t4zy[c])) +
                                          shfl_up1(fs_prev[c]);
                        const double er = ((srx_p[c] - t4yz[c]) - t4zy[c]) + fs_prev[c];
                        ex_v = upw(qx_pos[c], qx_nz[c], slow, el, er);
                        if (BC) ex_v = bc_final(i, g.dlo[0], g.dhi[0], bclx, bchx, false, 0.0, 0.0, el, er, ex_v);
                    }
                    {
                        const double el = c == 0 ? ((sly_p[0] - qy1_below) - qy2_below) + fs_below
                                                 : ((sly_p[1] - t4xz[0]) - t4zx[0]) + fs_prev[0];
                        const double er = ((sry_p[c] - t4xz[c]) - t4zx[c]) + fs_prev[c];
                        ey_v = upw(qy_pos[c], qy_nz[c], slow, el, er);
                        if (BC) ey_v = bc_final(jj[c], g.dlo[1], g.dhi[1], bcly, bchy, false, 0.0, 0.0, el, er, ey_v);
                    }
                    if (MODE == 1 && n == 0) {  // rho' edges are re-read for every species
                        if (ownx[c]) pEx[oEx] = ex_v;
                        if (owny[c]) pEy[oEy] = ey_v;
                    } else {
                        if (ownx[c]) stg_stream(pEx + oEx, ex_v);
                        if (owny[c]) stg_stream(pEy + oEy, ey_v);
                    }
                    if (MODE == 1 && flux1) {
                        // MakeRhoXFlux, MaestroMakeFlux.cpp:191-306 (planar) / :316-407 (spherical)
                        double r0x, r0y;
                        if (sph) {
                            r0x = pEx[dQx + oEx];  // mac_average output, laid out like sedge(comp)
                            r0y = pEy[dQy + oEy];
                        } else {
                            r0x = r0y = T_RXY[it - 1];
                        }
                        const double flx = rhoX_flux(F_.spt, uf_p[c], r0x, r_x[c], ex_v);
                        const double fly = rhoX_flux(F_.spt, vf_p[c], r0y, r_y[c], ey_v);
                        if (ownx[c]) {
                            stg_stream(pFx + oFx, flx);
                            pFx[dGx + oFx] = g_x[c] + flx;
                        }
                        if (owny[c]) {
                            stg_stream(pFy + oFy, fly);
                            pFy[dGy + oFy] = g_y[c] + fly;
                        }
                    } else if (MODE == 2) {
                        // MakeRhoHFlux (planar), MaestroMakeFlux.cpp:674-786
                        double fl;
                        if (rhoh_flux_tab(F_, have_h, uf_p[c], r_x[c], ex_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                            if (ownx[c]) stg_stream(pFx + oFx, fl);
                        if (rhoh_flux_tab(F_, have_h, vf_p[c], r_y[c], ey_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                            if (owny[c]) stg_stream(pFy + oFy, fl);
                    }
                }
            }
            {
                double t4xy_prev[2], t4yx_prev[2];
                HLD(H_T4XY, t4xy_prev);
                HLD(H_T4YX, t4yx_prev);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double el_z = ((slz[c] - t4xy_prev[c]) - t4yx_prev[c]) + fs_prev[c];
                    const double er_z = ((srz[c] - t4xy[c]) - t4yx[c]) + fs_cur[c];
                    if (do_z) {  // block uniform: z face a
                        const int oEz = c ? dEz1 : 0, oFz = c ? dFz1 : 0;
                        double ez_v = upw(pz_pos[c], pz_nz[c], slow, el_z, er_z);
                        if (BC) ez_v = bc_final(a, g.dlo[2], g.dhi[2], bclz, bchz, false, 0.0, 0.0, el_z, er_z, ez_v);
                        if (MODE == 1 && n == 0) {
                            if (ownz[c]) pEz[oEz] = ez_v;
                        } else {
                            if (ownz[c]) stg_stream(pEz + oEz, ez_v);
                        }
                        if (MODE == 1 && flux1) {
                            double r0z;
                            if (sph) {
                                r0z = pEz[dQz + oEz];
                            } else {
                                r0z = T_RZ[it];
                            }
                            const double flz = rhoX_flux(F_.spt, w_0[c], r0z, r_z[c], ez_v);
                            if (ownz[c]) {
                                stg_stream(pFz + oFz, flz);
                                pFz[dGz + oFz] = g_z[c] + flz;
                            }
                            if (A.do_eta) {
                                // MaestroMakeFlux.cpp:288-301
                                double ev = e_z[c];
                                if (is_spec) ev += flz;
                                if (last_spec) ev -= T_W0R[it];
                                if (ownz[c]) pEta[c ? dEta1 : 0] = ev;
                            }
                        } else if (MODE == 2) {
                            double fl;
                            if (rhoh_flux_tab(F_, have_h, w_0[c], r_z[c], ez_v, T_RZ[it], T_HZ[it], true, &fl))
                                if (ownz[c]) stg_stream(pFz + oFz, fl);
                        }
                    }
                }
            }
            // advance the output pointers to the next plane
            pEx += ex_sz;
            pEy += ey_sz;
            pEz += ez_sz;
            if (MODE != 0) {
                pFx += fx_sz;
                pFy += fy_sz;
                pFz += fz_sz;
                if (MODE == 1) pEta += eta_sz;
            }

            // ---- z history for the next plane -----------------------------------------------
            HST(H_SLX, slx);
            HST(H_SRX, srx);
            HST(H_SLY, sly);
            HST(H_SRY, sry);
            HST(H_UF, um);
            {
                const double vlo[2] = {vf[0], vf[1]};
                HST(H_VF, vlo);
            }
            HST(H_C4X, c4xu);
            HST(H_C4Y, c4yu);
            HST(H_HZX, hzx);
            HST(H_HZY, hzy);
            HST(H_T4XY, t4xy);
            HST(H_T4YX, t4yx);
            HST(H_FS, fs_cur);
            slow_p = slow_c;
        }
        cp_async_wait<0>();
    }
#undef HLD
#undef HST
}

template <int PPM, int MODE>
__global__ void __launch_bounds__(S5_NT, 1) k_stream5(const __grid_constant__ FusedArgs A) {
    extern __shared__ double sm[];
    // physical-boundary code only in blocks whose (grown) tile reaches a domain face that has
    // a non-interior BC; everything else runs the BC-free instantiation
    constexpr int BX = 30, BY = 14;
    const KGeom& g = A.g;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    const int x0 = g.vlo[0] + bx * BX, y0 = g.vlo[1] + by * BY;
    int k0, k1;
    chunk_range(A, blockIdx.y, k0, k1);
    bool bc = false;
    if (A.bcdir[0]) bc |= (x0 - 1 <= g.dlo[0] + 2) || (x0 + BX >= g.dhi[0] - 2);
    if (A.bcdir[1]) bc |= (y0 - 1 <= g.dlo[1] + 2) || (y0 + BY >= g.dhi[1] - 2);
    if (A.bcdir[2]) bc |= (k0 - 1 <= g.dlo[2] + 2) || (k1 >= g.dhi[2] - 2);
    if (bc)
        stream5_body<PPM, true, MODE>(A, sm);
    else
        stream5_body<PPM, false, MODE>(A, sm);
}

template <int PPM, int MODE>
int launch_stream5(const FusedArgs& A, cudaStream_t stream) {
    const size_t smem = sizeof(double) * (size_t)stream5_smem_doubles<PPM>();
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_stream5<PPM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    dim3 grid(A.ntx * A.nty, A.nzc);
    k_stream5<PPM, MODE><<<grid, S5_NT, smem, stream>>>(A);
    mx_count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int PPM>
int launch_stream5_mode(const FusedArgs& A, cudaStream_t stream) {
    if (A.mode == 0) return launch_stream5<PPM, 0>(A, stream);
    if (A.mode == 1) return launch_stream5<PPM, 1>(A, stream);
    return launch_stream5<PPM, 2>(A, stream);
}

#endif
