// mx_stream4.cuh — k_stream4: the streaming edge-state (+flux) kernel, fourth revision.
// Same decomposition, phases and arithmetic as the first streaming kernel of round 1 (k_stream,
// since removed); what changed is everything *around* the arithmetic, which the
// round-1 ncu capture showed to be ~60 % of the issued instructions
// (profiles/r1a_k_stream_v3_ncu_full_n256.txt):
//   * every global array is addressed through a per-thread running pointer that advances by
//     one plane stride per iteration (one IMAD.WIDE per array per plane instead of
//     index -> offset -> address chains, constant-bank reloads and sign extensions);
//   * column offsets of the face arrays are clamped into the tile once, so every load is
//     unconditional and only stores are predicated (no divergent regions around the flux);
//   * the rho'-edge values and the rho-flux / etarhoflux entries a face read-modify-writes are
//     loaded into registers two phases ahead instead of being prefetched into L1;
//   * the z stencil is read from an 8-slot shared-memory ring (planes a-LA .. a+LA resident,
//     a+LA+1 in flight) instead of a register window that was shifted every plane.
// Round 2: the ring is filled by TMA (template flag TMA): one elected thread issues one
// cp.async.bulk.tensor per plane from a 4-D tensor map (x, y, z, component) over the caller's
// ghost-grown state fab, completion is tracked by one mbarrier per ring slot, and the hardware
// zero-fills whatever lies outside the fab.  The per-thread staging plan of the cp.async
// version (two source pointers, two offsets, four predicates per thread) disappears.  TMA needs
// a 16-byte aligned base and 16-byte multiples as global strides, i.e. an even number of
// doubles per fab row (every BASELINE configuration); other boxes keep the cp.async version.
#ifndef MX_STREAM4_CUH
#define MX_STREAM4_CUH

constexpr int S4_NS = 8;  // ring slots (power of two)
// tile rows per block (incl. the one-cell halo) and resident blocks per SM
#ifndef S4_GY
#define S4_GY 16
#endif
#ifndef S4_MINB
#define S4_MINB 1
#endif

// doubles per ring slot: a 128-byte multiple (TMA destination alignment)
template <int PPM>
__host__ __device__ constexpr int s4_pls() {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    return ((32 + 2 * LA) * (S4_GY + 2 * LA) + 15) & ~15;
}
template <int PPM>
__host__ __device__ constexpr int stream4_smem_doubles() {
    return S4_NS * s4_pls<PPM>() + 9 * (32 * S4_GY + 64) + 16 * (32 * S4_GY) + 5 * (ST_MAXZC + 2) + 2 +
           S4_NS;  // + the mbarriers of the ring slots (8-byte aligned: every term before is even)
}

template <int PPM, bool BC, int MODE, bool SPH, bool TMA>
__device__ __forceinline__ void stream4_body(const FusedArgs& A, double* smem) {
    constexpr int GX = 32, GY = S4_GY;
    constexpr int LA = (PPM == 2) ? 3 : 2;  // stencil radius
    constexpr int PW = GX + 2 * LA, PH = GY + 2 * LA;
    constexpr int PL = PW * PH;       // doubles per staged plane
    constexpr int PLS = s4_pls<PPM>();  // ring slot pitch
    constexpr int NT = GX * GY;
    constexpr int BX = GX - 2, BY = GY - 2;
    constexpr int NLD = (PL + NT - 1) / NT;

    double* ring = smem;
    // y-neighbour exchange arrays, each with one row of never-written slack on either side: the
    // outermost tile rows read "their neighbour" there (their y-direction results are never
    // stored), so no thread ever reads a slot another thread writes without a barrier between
    constexpr int XS = NT + 2 * GX;
    double* ex = ring + S4_NS * PLS + GX;
    double* X_IPY = ex + 0 * XS;
    double* X_SY = ex + 1 * XS;
    double* X_T6X = ex + 2 * XS;
    double* X_TZ = ex + 3 * XS;
    double* X_HYX = ex + 4 * XS;
    double* X_HYZ = ex + 5 * XS;
    double* X_QY1 = ex + 6 * XS;
    double* X_QY2 = ex + 7 * XS;
    double* X_FS = ex + 8 * XS;
    double* hist = ex + 9 * XS - GX;  // thread-private z history (slot v of thread tid)
    double* tab = hist + 16 * NT;     // per-plane base-state values of this chunk
    double* T_RXY = tab;              // 0.5 (rho0_old + rho0_new)(p)          x/y faces
    double* T_RZ = tab + (ST_MAXZC + 2);       // 0.5 (rho0_edge_old + _new)(a) z faces
    double* T_W0R = tab + 2 * (ST_MAXZC + 2);  // w0(a) * rho0_predicted_edge(a)
    double* T_HXY = tab + 3 * (ST_MAXZC + 2);  // 0.5 (rhoh0_old + rhoh0_new)(p)
    double* T_HZ = tab + 4 * (ST_MAXZC + 2);   // 0.5 (rhoh0_edge_old + _new)(a)
    unsigned long long* const mbar =
        reinterpret_cast<unsigned long long*>(tab + 5 * (ST_MAXZC + 2) + 2);  // one per ring slot (TMA)

    // p advanced by one plane (byte stride taken from the kernel parameters)
#define S4_ADV(p_, bytes_) \
    p_ = reinterpret_cast<decltype(p_)>(reinterpret_cast<uintptr_t>(p_) + (uintptr_t)(bytes_))
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    double* const hs = hist + tid;
#define HS(v) hs[(v) * NT]
    enum { H_SLX, H_SRX, H_SLY, H_SRY, H_UF, H_VF, H_C4X, H_C4Y, H_HZX, H_HZY, H_T4XY, H_T4YX,
           H_FS, H_T6X, H_T6Y, H_SHZ };

    const KGeom& g = A.g;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    const int x0 = g.vlo[0] + bx * BX, y0 = g.vlo[1] + by * BY;
    const int thx = min(x0 + BX - 1, g.vhi[0]), thy = min(y0 + BY - 1, g.vhi[1]);
    const int i = x0 - 1 + tx, j = y0 - 1 + ty;
    int k0, k1;  // this block's z chunk, k1 exclusive
    chunk_range(A, blockIdx.y, k0, k1);
    const bool last = (k1 == g.vhi[2] + 1) && !g.skip_top_z;
    const int a0 = k0 - 1;

    // faces this thread stores: the tile's low faces, plus the high face only on the last tile
    bool ownx, owny, ownz;
    {
        const bool inG = (i <= thx + 1) && (j <= thy + 1);
        const bool vxi = inG && tx >= 1 && i <= thx;
        const bool vyj = inG && ty >= 1 && j <= thy;
        ownx = inG && tx >= 1 && vyj && (i <= thx || thx == g.vhi[0]);
        owny = inG && ty >= 1 && vxi && (j <= thy || thy == g.vhi[1]);
        ownz = vxi && vyj;
    }
    // clamped columns: cell-like arrays (1 ghost available) and the three face classes; every
    // lane then addresses an element that exists, and lanes outside the tile just do not store
    const int il = min(i, thx + 1), jl = min(j, thy + 1);
    const int icx = min(max(i, x0), thx + 1), jcx = min(max(j, y0), thy);  // x faces
    const int icy = min(max(i, x0), thx), jcy = min(max(j, y0), thy + 1);  // y faces
    const int icz = min(max(i, x0), thx), jcz = min(max(j, y0), thy);      // z faces
    const int rc = (ty + LA) * PW + (tx + LA);
    // BC instantiation: physical-boundary code only acts within a few cells of a domain face.
    // Warp-uniform "near a face" flags per direction let every other warp / plane skip it with
    // uniform branches (a tile on an x face pays for the x code only): the BC body executed
    // 1.54x the instructions of the BC-free one before (profiles/r1d_sweep.md).  The margin
    // (4 cells) covers every index the BC rows test: lo-1 .. lo+2, hi-2 .. hi+1.
    const bool nearx = BC && __any_sync(0xffffffffu, (i <= g.dlo[0] + 4) || (i >= g.dhi[0] - 4));
    const bool neary = BC && ((j <= g.dlo[1] + 4) || (j >= g.dhi[1] - 4));  // j is per warp

    const DFab& st = A.state;
    const long st_sz = st.sz;
    const long w_sz = A.umac[2].sz;
    const long v_sy = A.umac[1].sy;
    const KFlux& F_ = A.f;
    constexpr bool sph = SPH;  // spherical flux forms (density pass only; compile-time so that
                               // the planar instantiation carries no dead loads / registers)
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
    long ld_off[NLD];
    bool ld_do[NLD], ld_in[NLD];
    if (!TMA) {
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int e = tid + q * NT;
            const int px = e % PW, py = e / PW;
            const int gi = x0 - 1 - LA + px - st.lo[0], gj = y0 - 1 - LA + py - st.lo[1];
            ld_do[q] = e < PL;
            ld_in[q] = ld_do[q] && gi >= 0 && gi < st.n[0] && gj >= 0 && gj < st.n[1];
            ld_off[q] = ld_in[q] ? (long)gj * st.sy + gi : 0;
        }
    }
    // TMA: tile origin in the fab's own coordinates (may lie outside: zero fill), per-slot
    // parity of the next wait (every thread waits for every plane exactly once)
    const int cx0 = x0 - 1 - LA - st.lo[0], cy0 = y0 - 1 - LA - st.lo[1], cz0 = a0 - LA - st.lo[2];
    unsigned phase = 0;

    const double c6x = A.c6h[0], c6y = A.c6h[1], c6z = A.c6h[2];
    const double c4x = A.c4h[0], c4y = A.c4h[1], c4z = A.c4h[2];
    const double eps = A.eps, dt = A.dt;
    // flux-form switches (block uniform)
    const bool have_h = (MODE == 2) && (F_.ept == MX_PRED_H || F_.ept == MX_PRED_T_THEN_H ||
                                        F_.ept == MX_PRED_TPRIME_THEN_H);
    // sedge(Rho) is an input of the flux: species forms except (rho X); h forms; spherical h'
    const bool need_ser = (MODE == 1) ? (F_.spt != MX_PRED_RHOX)
                                      : (have_h || (sph && F_.ept == MX_PRED_HPRIME));

    // velocity / force columns are component independent
    const long u_c = A.umac[0].off(il, jl, a0), v_c = A.umac[1].off(il, jl, a0);
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
        const bool lobcz_c = BC && (bclz == BC_EXT_DIR || bclz == BC_HOEXTRAP);
        const bool hibcz_c = BC && (bchz == BC_EXT_DIR || bchz == BC_HOEXTRAP);
        const bool bcx = nearx && (bclx != 0 || bchx != 0);  // warp uniform
        const bool bcy = neary && (bcly != 0 || bchy != 0);  // warp uniform
        const bool bcz_c = BC && (bclz != 0 || bchz != 0);   // block uniform; per plane below
        // is_vel: the normal velocity component is clamped at outflow faces
        // (MaestroMakeEdgeScalars.cpp:716-728 and the same blocks of the later stages)
        const bool vcx = BC && A.is_vel && comp == 0, vcy = BC && A.is_vel && comp == 1,
                   vcz = BC && A.is_vel && comp == 2;
        const bool flux1 = (MODE == 1 && n > 0);
        const bool first_spec = (n == 1);
        const bool is_spec =
            comp >= F_.first_spec && comp <= F_.first_spec + F_.num_spec - 1;  // eta: :288-301
        const bool last_spec = comp == F_.first_spec + F_.num_spec - 1;

        // ---- running pointers: they always address the plane this iteration touches ------
        const double* __restrict__ pU = A.umac[0].p + u_c;   // plane a
        const double* __restrict__ pV = A.umac[1].p + v_c;   // plane a
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
        const double* __restrict__ pS[NLD];  // source of ring element q, plane being staged
        if (!TMA) {
            const double* sp_ = st.comp(comp) + (long)(a0 - LA - st.lo[2]) * st_sz;
#pragma unroll
            for (int q = 0; q < NLD; ++q) pS[q] = sp_ + ld_off[q];
        }
        // stage plane number q of this component's sweep (plane a0 - LA + q) into its ring slot
        auto stage_plane = [&](int q_) {
            const int slot = q_ & (S4_NS - 1);
            if (TMA) {
                if (tid == 0) {
                    mbar_expect_tx(mbar + slot, PL * 8);
                    tma_load_plane(ring + slot * PLS, &A.tmap_state, mbar + slot, cx0, cy0, cz0 + q_, comp);
                }
            } else {
                double* dst = ring + slot * PLS + tid;
#pragma unroll
                for (int q = 0; q < NLD; ++q) {
                    if (ld_do[q]) cp_async8(dst + q * NT, pS[q], ld_in[q]);
                    pS[q] += st_sz;
                }
            }
        };
        auto wait_plane = [&](int q_) {  // TMA only: plane q_ has landed (acquire for this thread)
            const int slot = q_ & (S4_NS - 1);
            mbar_wait(mbar + slot, (phase >> slot) & 1u);
            phase ^= 1u << slot;
        };

        // ---- prologue ---------------------------------------------------------------
        __syncthreads();  // previous component is done with shared memory; tables written
#pragma unroll
        for (int q = 0; q <= 2 * LA; ++q) stage_plane(q);  // planes a0-LA .. a0+LA
        if (!TMA) cp_async_commit();
        // register prefetch of plane a0
        double w_m1 = 0.0, w_0 = 0.0;
        double w_p1 = pW[-w_sz];  // face a0
        double um_n = pU[0];
        double up31_n = lane31 ? pU[1] : 0.0;
        double vm_n = pV[0];
        double vp_n = pV[v_sy];
        double w_n = pW[0];  // face a0+1
        double f_n = pF[0];
        double sgz_hi = (PPM != 0) ? div_h<2>(A, fabs(w_p1) * dt) : 0.0;  // sigma of face a0
        double ipz_prev = 0.0;
        double dz_c = 0.0, ez_lo = 0.0;                // PPM 1: V(a), E(a-1/2)
        double ez_l2 = 0.0, ez_l1 = 0.0, ez_h1 = 0.0;  // PPM 2: E(a-3/2), E(a-1/2), E(a+1/2)
#pragma unroll
        for (int v = 0; v < 16; ++v) HS(v) = 0.0;
        bool slow_p = true;
        if (TMA) {
#pragma unroll
            for (int q = 0; q <= 2 * LA; ++q) wait_plane(q);
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();  // planes a0-LA .. a0+LA are in the ring
        {
            // incremental z reconstruction state from planes a0-2 .. a0+1
            const double* c = ring + rc;
            if (PPM == 1) {
                const double zm2 = c[(LA - 2) * PLS], zm1 = c[(LA - 1) * PLS], z0 = c[LA * PLS],
                             zp1 = c[(LA + 1) * PLS];
                const double dm1 = vanleer(zm2, zm1, z0);
                dz_c = vanleer(zm1, z0, zp1);
                ez_lo = vl_edge(zm1, z0, dm1, dz_c);
            } else if (PPM == 2) {
                const double z[6] = {c[0], c[PLS], c[2 * PLS], c[3 * PLS], c[4 * PLS], c[5 * PLS]};
                ez_l2 = cs_edge(z[0], z[1], z[2], z[3]);
                ez_l1 = cs_edge(z[1], z[2], z[3], z[4]);
                ez_h1 = cs_edge(z[2], z[3], z[4], z[5]);
            }
        }

        for (int a = a0; a <= k1; ++a) {
            const int it = a - a0;
            const bool bcz = bcz_c && ((a <= g.dlo[2] + 4) || (a >= g.dhi[2] - 4));  // block uniform
            const bool lobcz = lobcz_c && bcz, hibcz = hibcz_c && bcz;
            const double* pa = ring + ((it + LA) & (S4_NS - 1)) * PLS + rc;  // plane a, my cell
            // ---- P0: current-plane registers, prefetch of the next plane ----------------
#ifdef MX_EXP_FAKEIN
            // experiment (wrong results): velocities / force read from a shared-memory plane as a
            // loader warp would provide them -> no input pointers, prefetch or rotation registers
            const double* fk = tab + (it & 3);
            const double um = fk[tid & 63], up = fk[(tid & 63) + 1], vm = fk[(tid & 31) + 2],
                         vp = fk[(tid & 31) + 34], f0 = fk[(tid & 15) + 7];
            w_m1 = fk[(tid & 31) + 3];
            w_0 = fk[(tid & 31) + 4];
            w_p1 = fk[(tid & 31) + 5];
            if (a < k1) stage_plane(it + 2 * LA + 1);
            if (false) {
#else
            const double um = um_n, vm = vm_n, vp = vp_n, f0 = f_n;
            double up = shfl_dn1(um);
            if (lane31) up = up31_n;
            w_m1 = w_0;
            w_0 = w_p1;
            w_p1 = w_n;
            if (a < k1) {  // block uniform
                stage_plane(it + 2 * LA + 1);  // plane a+LA+1
#endif
                S4_ADV(pU, A.bs[0]);
                S4_ADV(pV, A.bs[1]);
                S4_ADV(pW, A.bs[2]);
                S4_ADV(pF, A.bs[3]);
                um_n = pU[0];
                if (lane31) up31_n = pU[1];
                vm_n = pV[0];
                vp_n = pV[v_sy];
                w_n = pW[0];
                f_n = pF[0];
            }
            if (!TMA) cp_async_commit();
            // velocity predicates of this plane's faces
            const bool px_pos = um > 0.0, px_nz0 = fabs(um) > 0.0, px_nz = fabs(um) > eps;
            const bool py_pos = vm > 0.0, py_nz0 = fabs(vm) > 0.0, py_nz = fabs(vm) > eps;
            const bool pz_pos = w_0 > 0.0, pz_nz0 = fabs(w_0) > 0.0, pz_nz = fabs(w_0) > eps;
            // high y face (its first Riemann solve is formed here too, see P2)
            const bool py2_pos = vp > 0.0, py2_nz0 = fabs(vp) > 0.0;
            const bool slow_c = __any_sync(
                0xffffffffu,
                !(px_nz0 && px_nz && py_nz0 && py_nz && pz_nz0 && pz_nz && py2_nz0));
            const bool slow = slow_c || slow_p;

            // ---- P1: x / y reconstruction of plane a ------------------------------------
            const double s0 = pa[0];
            double ipx, imx, ipy, imy;
            {
                const bool lobcx = bcx && (bclx == BC_EXT_DIR || bclx == BC_HOEXTRAP);
                const bool hibcx = bcx && (bchx == BC_EXT_DIR || bchx == BC_HOEXTRAP);
                const bool lobcy = bcy && (bcly == BC_EXT_DIR || bcly == BC_HOEXTRAP);
                const bool hibcy = bcy && (bchy == BC_EXT_DIR || bchy == BC_HOEXTRAP);
                double S[7], T[7];
                S[0] = S[6] = T[0] = T[6] = 0.0;
#pragma unroll
                for (int q = -LA; q <= LA; ++q) {
                    S[q + 3] = pa[q];
                    T[q + 3] = pa[q * PW];
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
            if (PPM != 0) X_SY[tid] = imy;  // right state of my low y face, for the row below
            if (TMA) {
                if (it > 0) wait_plane(it + 2 * LA);  // plane a+LA has landed (it == 0: prologue)
            } else {
                cp_async_wait<1>();  // this thread's share of plane a+LA has landed
            }
            __syncthreads();  // S2: publishes IPY (and, cp.async, ring plane a+LA)

            // ---- P2: z reconstruction, 1-D states, first Riemann solve -------------------
            double ipz, imz, s_zm1 = 0.0;
            {
                double S[7];
                S[0] = S[6] = 0.0;
                // planes a-LA .. a+LA; without BC rows PPM 1 only reads a .. a+2
#pragma unroll
                for (int q = -LA; q <= LA; ++q) {
                    const bool need = (BC && bcz) || PPM != 1 || q > 0;
                    S[q + 3] = (q == 0) ? s0
                               : need   ? ring[((it + LA + q) & (S4_NS - 1)) * PLS + rc]
                                        : 0.0;
                }
                if (PPM == 0) {
                    ipz = slope_any(A.slope_order, S + 1, a, g.dlo[2], g.dhi[2], lobcz, hibcz, false);
                    imz = 0.0;
                    s_zm1 = S[2];
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
                slx = pa[-1] + 0.5 * (1.0 - qx) * shfl_up1(ipx);
                srx = s0 - 0.5 * (1.0 + qx) * ipx;
                sly = pa[-PW] + 0.5 * (1.0 - qy) * X_IPY[tid - GX];
                sry = s0 - 0.5 * (1.0 + qy) * ipy;
                slz = s_zm1 + 0.5 * (1.0 - qz) * ipz_prev;
                srz = s0 - 0.5 * (1.0 + qz) * ipz;
            } else {
                slx = shfl_up1(ipx);
                srx = imx;
                sly = X_IPY[tid - GX];
                sry = imy;
                slz = ipz_prev;
                srz = imz;
            }
            ipz_prev = ipz;
            if (BC) {
                if (bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, slx, srx);
                if (bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, sly, sry);
                if (bcz) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, slz, srz);
            }
            const double shx = upw(px_pos, px_nz0, slow, slx, srx);  // QUIRK: |u| > 0.0 (:746)
            const double shy = upw(py_pos, py_nz0, slow, sly, sry);
            const double shz = upw(pz_pos, pz_nz0, slow, slz, srz);
            // first Riemann solve of my HIGH y face (the row above's low face), formed here from
            // my own left state and the neighbour's right state published at S2: saves the
            // exchange round (barrier S3) that passing its result down would cost
            double shy_up;
            {
                double l, r;
                if (PPM == 0) {
                    const double qy = div_h<1>(A, dt * vp);
                    l = s0 + 0.5 * (1.0 - qy) * ipy;
                    r = pa[PW] - 0.5 * (1.0 + qy) * X_IPY[tid + GX];
                } else {
                    l = ipy;
                    r = X_SY[tid + GX];
                }
                if (BC && bcy) bc_pair(j + 1, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, l, r);
                shy_up = upw(py2_pos, py2_nz0, slow, l, r);
            }

            // ---- P3: transverse differences (dt/6 form) ---------------------------------
            const double usx = up + um, usy = vp + vm;
            const double t6x = (c6x * usx) * (shfl_dn1(shx) - shx);
            const double t6y = (c6y * usy) * (shy_up - shy);
            const double wsum_p = w_0 + w_m1;  // faces a and a-1: cell a-1
            const double tz = (c6z * wsum_p) * (shz - HS(H_SHZ));
            HS(H_SHZ) = shz;
            X_T6X[tid] = t6x;
            X_TZ[tid] = tz;
            // what P6 of this iteration reads (rho' edges) and read-modify-writes (rho flux,
            // etarhoflux): issue the loads now, three phases ahead of their use
            const bool do_xy = a > k0;                         // x/y faces of plane a-1 exist
            const bool do_z = a >= k0 && (a < k1 || last);     // z face a belongs to this chunk
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
            __syncthreads();  // S4

            // ---- P4: corner-coupled states ----------------------------------------------
            const double uf_p = HS(H_UF), vf_p = HS(H_VF);
            const bool qx_pos = uf_p > 0.0, qx_nz = fabs(uf_p) > eps;
            const bool qy_pos = vf_p > 0.0, qy_nz = fabs(vf_p) > eps;
            const double slx_p = HS(H_SLX), srx_p = HS(H_SRX);
            const double sly_p = HS(H_SLY), sry_p = HS(H_SRY);
            double hxy, hyx, hzx, hzy, hxz, hyz;
            {
                double l = slx - shfl_up1(t6y), r = srx - t6y;
                if (BC && bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, l, r);
                hxy = upw(px_pos, px_nz, slow, l, r);
                l = sly - X_T6X[tid - GX];
                r = sry - t6x;
                if (BC && bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, l, r);
                hyx = upw(py_pos, py_nz, slow, l, r);
                l = slz - HS(H_T6X);
                r = srz - t6x;
                if (BC && bcz) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, l, r);
                hzx = upw(pz_pos, pz_nz, slow, l, r);
                l = slz - HS(H_T6Y);
                r = srz - t6y;
                if (BC && bcz) bc_pair(a, g.dlo[2], g.dhi[2], bclz, bchz, vcz, 0.0, 0.0, l, r);
                hzy = upw(pz_pos, pz_nz, slow, l, r);
                l = slx_p - shfl_up1(tz);
                r = srx_p - tz;
                if (BC && bcx) bc_pair(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, l, r);
                hxz = upw(qx_pos, qx_nz, slow, l, r);
                l = sly_p - X_TZ[tid - GX];
                r = sry_p - tz;
                if (BC && bcy) bc_pair(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, l, r);
                hyz = upw(qy_pos, qy_nz, slow, l, r);
            }
            HS(H_T6X) = t6x;
            HS(H_T6Y) = t6y;
            X_HYX[tid] = hyx;
            X_HYZ[tid] = hyz;
            // the dt/4 terms the row above needs for its low y face do not depend on this
            // exchange round (x shuffle / own z history only): publish them with it, which
            // makes a separate round (barrier S6) unnecessary
            const double c4zw = c4z * wsum_p;
            const double t4xz = HS(H_C4X) * (shfl_dn1(hxz) - hxz);
            const double t4zx = c4zw * (hzx - HS(H_HZX));
            const double fs_prev = HS(H_FS);
            X_QY1[tid] = t4xz;
            X_QY2[tid] = t4zx;
            X_FS[tid] = fs_prev;
            __syncthreads();  // S5

            // ---- P5: transverse differences (dt/4 form) ---------------------------------
            const double c4xu = c4x * usx, c4yu = c4y * usy;
            const double t4xy = c4xu * (shfl_dn1(hxy) - hxy);
            const double t4yx = c4yu * (X_HYX[tid + GX] - hyx);
            const double t4yz = HS(H_C4Y) * (X_HYZ[tid + GX] - hyz);
            const double t4zy = c4zw * (hzy - HS(H_HZY));
            const double fs_cur = A.dt2 * f0;

            // ---- P6: final edge states (+ fluxes) ---------------------------------------
            if (do_xy) {  // block uniform: x / y faces of plane p = a-1
                double ex_v, ey_v;
                {
                    const double el = ((slx_p - shfl_up1(t4yz)) - shfl_up1(t4zy)) + shfl_up1(fs_prev);
                    const double er = ((srx_p - t4yz) - t4zy) + fs_prev;
                    ex_v = upw(qx_pos, qx_nz, slow, el, er);
                    if (BC && bcx) ex_v = bc_final(i, g.dlo[0], g.dhi[0], bclx, bchx, vcx, 0.0, 0.0, el, er, ex_v);
                }
                {
                    const double el = ((sly_p - X_QY1[tid - GX]) - X_QY2[tid - GX]) + X_FS[tid - GX];
                    const double er = ((sry_p - t4xz) - t4zx) + fs_prev;
                    ey_v = upw(qy_pos, qy_nz, slow, el, er);
                    if (BC && bcy) ey_v = bc_final(j, g.dlo[1], g.dhi[1], bcly, bchy, vcy, 0.0, 0.0, el, er, ey_v);
                }
                if (ownx) pEx[0] = ex_v;
                if (owny) pEy[0] = ey_v;
                if (MODE == 1 && flux1) {
                    // MakeRhoXFlux, MaestroMakeFlux.cpp:191-306 (planar) / :316-407 (spherical)
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
                    // MakeRhoHFlux (planar), MaestroMakeFlux.cpp:674-786
                    double fl;
                    if (rhoh_flux_tab(F_, have_h, uf_p, r_x, ex_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                        if (ownx) pFx[0] = fl;
                    if (rhoh_flux_tab(F_, have_h, vf_p, r_y, ey_v, T_RXY[it - 1], T_HXY[it - 1], false, &fl))
                        if (owny) pFy[0] = fl;
                }
            }
            const double el_z = ((slz - HS(H_T4XY)) - HS(H_T4YX)) + fs_prev;
            const double er_z = ((srz - t4xy) - t4yx) + fs_cur;
            if (do_z) {  // block uniform: z face a
                double ez_v = upw(pz_pos, pz_nz, slow, el_z, er_z);
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
                        // MaestroMakeFlux.cpp:288-301
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
            // advance the output pointers to the next plane
            S4_ADV(pEx, A.bs[4]);
            S4_ADV(pEy, A.bs[5]);
            S4_ADV(pEz, A.bs[6]);
            if (MODE != 0) {
                S4_ADV(pFx, A.bs[7]);
                S4_ADV(pFy, A.bs[8]);
                S4_ADV(pFz, A.bs[9]);
                if (MODE == 1) S4_ADV(pEta, A.bs[10]);
            }

            // ---- z history for the next plane -----------------------------------------------
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
        if (!TMA) cp_async_wait<0>();
    }
#undef HS
#undef S4_ADV
}

template <int PPM, int MODE, bool SPH, bool TMA>
__global__ void __launch_bounds__(32 * S4_GY, S4_MINB) k_stream4(const __grid_constant__ FusedArgs A) {
    extern __shared__ __align__(128) double sm[];
    // physical-boundary code only in blocks whose (grown) tile reaches a domain face that has
    // a non-interior BC; everything else runs the BC-free instantiation
    constexpr int BX = 30, BY = S4_GY - 2;
    const KGeom& g = A.g;
    const int bx = blockIdx.x % A.ntx, by = blockIdx.x / A.ntx;
    const int x0 = g.vlo[0] + bx * BX, y0 = g.vlo[1] + by * BY;
    int k0, k1;
    chunk_range(A, blockIdx.y, k0, k1);
    bool bc = false;
    if (A.bcdir[0]) bc |= (x0 - 1 <= g.dlo[0] + 2) || (x0 + BX >= g.dhi[0] - 2);
    if (A.bcdir[1]) bc |= (y0 - 1 <= g.dlo[1] + 2) || (y0 + BY >= g.dhi[1] - 2);
    if (A.bcdir[2]) bc |= (k0 - 1 <= g.dlo[2] + 2) || (k1 >= g.dhi[2] - 2);
    if (TMA && threadIdx.x == 0) {
        // one mbarrier per ring slot; the body's first __syncthreads publishes them
        unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(
            sm + stream4_smem_doubles<PPM>() - S4_NS);
#pragma unroll
        for (int q = 0; q < S4_NS; ++q) mbar_init(mbar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (bc)
        stream4_body<PPM, true, MODE, SPH, TMA>(A, sm);
    else
        stream4_body<PPM, false, MODE, SPH, TMA>(A, sm);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// 4-D tensor map (x, y, z, component) over the caller's ghost-grown state fab, box = one staged
// plane {PW, PH, 1, 1}.  cuTensorMapEncodeTiled is fetched through the runtime's driver entry
// point query, so libmxadv.so does not link libcuda.
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

// false: this box cannot take TMA (odd fab rows / unaligned base / no driver entry point)
template <int PPM>
bool stream4_tmap(FusedArgs& A) {
    constexpr int LA = (PPM == 2) ? 3 : 2;
    const DFab& st = A.state;
    if ((reinterpret_cast<uintptr_t>(st.p) & 15u) != 0 || (st.n[0] & 1) != 0) return false;
    mx_encode_tiled_fn enc = encode_tiled();
    if (enc == nullptr) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)st.n[0], (cuuint64_t)st.n[1], (cuuint64_t)st.n[2],
                                (cuuint64_t)st.ncomp};
    const cuuint64_t strides[3] = {(cuuint64_t)st.sy * 8, (cuuint64_t)st.sz * 8, (cuuint64_t)st.sn * 8};
    const cuuint32_t box[4] = {32 + 2 * LA, S4_GY + 2 * LA, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(&A.tmap_state, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, st.p, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int PPM, int MODE, bool SPH, bool TMA>
int launch_stream4(const FusedArgs& A, cudaStream_t stream) {
    const size_t smem = sizeof(double) * (size_t)stream4_smem_doubles<PPM>();
    // cudaFuncSetAttribute is per device: remember which devices have it (bit d of the mask)
    static std::atomic<unsigned long long> attr_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & bit)) {
        if (cudaFuncSetAttribute(k_stream4<PPM, MODE, SPH, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return -1;
        attr_set.fetch_or(bit, std::memory_order_relaxed);
    }
    dim3 grid(A.ntx * A.nty, A.nzc);
    k_stream4<PPM, MODE, SPH, TMA><<<grid, 32 * S4_GY, smem, stream>>>(A);
    mx_count_launch();
    if (TMA) mx_count_tma_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <int PPM, bool TMA>
int launch_stream4_tma(const FusedArgs& A, cudaStream_t stream) {
    if (A.mode == 0) return launch_stream4<PPM, 0, false, TMA>(A, stream);
    if (A.mode == 1)
        return A.f.spherical ? launch_stream4<PPM, 1, true, TMA>(A, stream)
                             : launch_stream4<PPM, 1, false, TMA>(A, stream);
    return A.f.spherical ? launch_stream4<PPM, 2, true, TMA>(A, stream)
                         : launch_stream4<PPM, 2, false, TMA>(A, stream);
}

// MXADV_NO_TMA=1 forces the cp.async staging (A/B runs)
template <int PPM>
int launch_stream4_mode(FusedArgs& A, cudaStream_t stream) {
    if (!getenv("MXADV_NO_TMA") && stream4_tmap<PPM>(A)) return launch_stream4_tma<PPM, true>(A, stream);
    return launch_stream4_tma<PPM, false>(A, stream);
}

#endif
