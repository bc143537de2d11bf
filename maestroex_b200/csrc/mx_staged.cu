// mx_staged.cu — general ("staged") CUDA path of Maestro::MakeEdgeScal for sm_100a.
//
// One kernel per stage of the reference algorithm, intermediates in a device workspace, any
// option combination the reference supports (2-D / 3-D, ppm_type 0/1/2, slope_order 0/2/4,
// conservative form, ppm_trace_forces, is_vel, every BC code).  This is the coverage path;
// the fused single-pass kernels in mx_fused.cu are the performance path for the 3-D
// non-conservative scalar case and produce bit-identical results.
//
// Stage -> reference:
//   k_divu        MakeDivU                 Source/MaestroMakeEdgeScalars.cpp:648-660
//   k_parabola    Slopex/y/z | PPM         Source/MaestroSlopes.cpp, Source/MaestroPPM.cpp
//   k_predict     MakeEdgeScalPredictor    MaestroMakeEdgeScalars.cpp:306-452 / :662-863
//   k_transverse  MakeEdgeScalTransverse   :865-1359
//   k_edges       MakeEdgeScalEdges        :454-644 / :1361-1667
#include "mx_common.h"
#include "mx_device.cuh"
#include "mx_staged.h"

using namespace mxd;

namespace {

constexpr int TPB = 256;

struct CellIdx {
    int i, j, k;
};

__device__ __forceinline__ bool cell_of(const Scratch& w, long t, CellIdx& c) {
    const long ncell = (long)w.n[0] * w.n[1] * w.n[2];
    if (t >= ncell) return false;
    const int ii = (int)(t % w.n[0]);
    const long r = t / w.n[0];
    c.i = w.lo[0] + ii;
    c.j = w.lo[1] + (int)(r % w.n[1]);
    c.k = w.lo[2] + (int)(r / w.n[1]);
    return true;
}

__device__ __forceinline__ long woff(const Scratch& w, int i, int j, int k) {
    return (long)(i - w.lo[0]) + (long)(j - w.lo[1]) * w.sy + (long)(k - w.lo[2]) * w.sz;
}

__global__ void k_divu(KGeom g, DFab u, DFab v, DFab wv, Scratch w) {
    CellIdx c;
    if (!cell_of(w, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    // QUIRK: whole sum divided by dx[0] (MaestroMakeEdgeScalars.cpp:656-658)
    double d = u.p[u.off(c.i + 1, c.j, c.k)] - u.p[u.off(c.i, c.j, c.k)] +
               v.p[v.off(c.i, c.j + 1, c.k)] - v.p[v.off(c.i, c.j, c.k)] +
               wv.p[wv.off(c.i, c.j, c.k + 1)] - wv.p[wv.off(c.i, c.j, c.k)];
    w.divu[woff(w, c.i, c.j, c.k)] = d / g.dx[0];
}

// slopes or PPM parabola + characteristic integration, all directions, one component of
// `src` (the state, or the force when tracing forces)
__global__ void k_parabola(KGeom g, KEdge e, DFab src, DFab u0, DFab u1, DFab u2, Scratch w,
                           int to_force) {
    CellIdx c;
    if (!cell_of(w, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    const double* sp_ = src.comp(e.comp);
    const long o = src.off(c.i, c.j, c.k);
    const long wo = woff(w, c.i, c.j, c.k);
    const DFab* vel[3] = {&u0, &u1, &u2};
    for (int d = 0; d < g.dim; ++d) {
        const long st = d == 0 ? 1 : (d == 1 ? src.sy : src.sz);
        const int I = d == 0 ? c.i : (d == 1 ? c.j : c.k);
        const bool lobc = e.bclo[d] == BC_EXT_DIR || e.bclo[d] == BC_HOEXTRAP;
        const bool hibc = e.bchi[d] == BC_EXT_DIR || e.bchi[d] == BC_HOEXTRAP;
        double S[7];
        S[0] = S[6] = 0.0;
#pragma unroll
        for (int q = -2; q <= 2; ++q) S[q + 3] = sp_[o + q * st];
        if (e.ppm_type == 2) {
            S[0] = sp_[o - 3 * st];
            S[6] = sp_[o + 3 * st];
        }
        if (e.ppm_type == 0) {
            double r;
            if (e.slope_order == 0) {
                r = 0.0;
            } else if (e.slope_order == 2) {
                r = slope2(S + 1, I, g.dlo[d], g.dhi[d], lobc, hibc);
            } else {
                r = slope4(S + 1, I, g.dlo[d], g.dhi[d], lobc, hibc, d == 0);
            }
            w.Ip[d][wo] = r;
        } else {
            double sm, sp;
            if (e.ppm_type == 1) {
                ppm1_parabola<true>(S, I, g.dlo[d], g.dhi[d], lobc, hibc, sm, sp);
            } else {
                ppm2_parabola<true>(S, I, g.dlo[d], g.dhi[d], lobc, hibc, d == 0, sm, sp);
            }
            const DFab& V = *vel[d];
            const long vo = V.off(c.i, c.j, c.k);
            const long vst = d == 0 ? 1 : (d == 1 ? V.sy : V.sz);
            const double um = V.p[vo], up = V.p[vo + vst];
            const double s0 = S[3];
            const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
            const double sig_p = fabs(up) * e.dt / g.dx[d];
            const double sig_m = fabs(um) * e.dt / g.dx[d];
            double* Ip = to_force ? w.Ipf[d] : w.Ip[d];
            double* Im = to_force ? w.Imf[d] : w.Im[d];
            Ip[wo] = ppm_Ip(sm, sp, s0, s6, up, sig_p, e.eps);
            Im[wo] = ppm_Im(sm, sp, s0, s6, um, sig_m, e.eps);
        }
    }
}

__global__ void k_predict(KGeom g, KEdge e, DFab st, DFab u0, DFab u1, DFab u2, Scratch w) {
    CellIdx c;
    if (!cell_of(w, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    const double* s = st.comp(e.comp);
    const long so = st.off(c.i, c.j, c.k);
    const long wo = woff(w, c.i, c.j, c.k);
    const DFab* vel[3] = {&u0, &u1, &u2};
    // QUIRK: 3-D predictor upwinds with |u| > 0 (:746), 2-D with rel_eps (:389)
    const double eps = g.dim == 3 ? 0.0 : e.eps;
    for (int d = 0; d < g.dim; ++d) {
        const int I = d == 0 ? c.i : (d == 1 ? c.j : c.k);
        if (I < w.lo[d] + 1) continue;
        const long sst = d == 0 ? 1 : (d == 1 ? st.sy : st.sz);
        const long wst = d == 0 ? 1 : (d == 1 ? w.sy : w.sz);
        const double u = vel[d]->p[vel[d]->off(c.i, c.j, c.k)];
        const double sL = s[so - sst], sR = s[so];
        double l, r;
        if (e.ppm_type == 0) {
            const double slL = w.Ip[d][wo - wst], slR = w.Ip[d][wo];
            const double h = g.dx[d];
            if (g.dim == 3) {
                l = sL + 0.5 * (1.0 - e.dt * u / h) * slL;
                r = sR - 0.5 * (1.0 + e.dt * u / h) * slR;
            } else {
                // QUIRK: 2-D association (:341-345)
                const double dt2 = 0.5 * e.dt;
                l = sL + (0.5 - dt2 * u / h) * slL;
                r = sR - (0.5 + dt2 * u / h) * slR;
            }
        } else {
            l = w.Ip[d][wo - wst];
            r = w.Im[d][wo];
        }
        bc_pair(I, g.dlo[d], g.dhi[d], e.bclo[d], e.bchi[d], e.is_vel && e.comp == d, sL, sR, l,
                r);
        w.sl[d][wo] = l;
        w.sr[d][wo] = r;
        w.simh[d][wo] = upwind(u, eps, l, r);
    }
}

__global__ void k_transverse(KGeom g, KEdge e, DFab st, DFab u0, DFab u1, DFab u2, Scratch w) {
    CellIdx c;
    if (!cell_of(w, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    const double* s = st.comp(e.comp);
    const long so = st.off(c.i, c.j, c.k);
    const long wo = woff(w, c.i, c.j, c.k);
    const DFab* vel[3] = {&u0, &u1, &u2};
    const int idx[3] = {c.i, c.j, c.k};
    const double dt3 = e.dt / 3.0, dt6 = e.dt / 6.0;
    for (int d = 0; d < 3; ++d) {
        for (int t = 0; t < 3; ++t) {
            if (t == d) continue;
            const int o = 3 - d - t;
            // box: valid, d-faces up to vhi+1, grown by one in o (:897-898)
            if (idx[d] < g.vlo[d] || idx[d] > g.vhi[d] + 1) continue;
            if (idx[t] < g.vlo[t] || idx[t] > g.vhi[t]) continue;
            if (idx[o] < g.vlo[o] - 1 || idx[o] > g.vhi[o] + 1) continue;
            const long sst = d == 0 ? 1 : (d == 1 ? st.sy : st.sz);
            const long wsd = d == 0 ? 1 : (d == 1 ? w.sy : w.sz);
            const long wstt = t == 0 ? 1 : (t == 1 ? w.sy : w.sz);
            const DFab& Vt = *vel[t];
            const long vtt = t == 0 ? 1 : (t == 1 ? Vt.sy : Vt.sz);
            const long vtd = d == 0 ? 1 : (d == 1 ? Vt.sy : Vt.sz);
            const long vR = Vt.off(c.i, c.j, c.k), vL = vR - vtd;
            const double vLp = Vt.p[vL + vtt], vLm = Vt.p[vL], vRp = Vt.p[vR + vtt],
                         vRm = Vt.p[vR];
            const long wL = wo - wsd;
            const double hL_p = w.simh[t][wL + wstt], hL_m = w.simh[t][wL];
            const double hR_p = w.simh[t][wo + wstt], hR_m = w.simh[t][wo];
            const double sL = s[so - sst], sR = s[so];
            const double ht = g.dx[t];
            double l, r;
            if (e.is_cons) {
                // :909-921
                l = w.sl[d][wo] - (dt3 / ht) * (hL_p * vLp - hL_m * vLm) -
                    dt3 * sL * w.divu[wL] + (dt3 / ht) * sL * (vLp - vLm);
                r = w.sr[d][wo] - (dt3 / ht) * (hR_p * vRp - hR_m * vRm) -
                    dt3 * sR * w.divu[wo] + (dt3 / ht) * sR * (vRp - vRm);
            } else {
                // :924-929
                l = w.sl[d][wo] - (dt6 / ht) * (vLp + vLm) * (hL_p - hL_m);
                r = w.sr[d][wo] - (dt6 / ht) * (vRp + vRm) * (hR_p - hR_m);
            }
            bc_pair(idx[d], g.dlo[d], g.dhi[d], e.bclo[d], e.bchi[d], e.is_vel && e.comp == d,
                    sL, sR, l, r);
            const double ud = vel[d]->p[vel[d]->off(c.i, c.j, c.k)];
            w.tr[d][t][wo] = upwind(ud, e.eps, l, r);
        }
    }
}

__global__ void k_edges(KGeom g, KEdge e, DFab st, DFab fo, DFab u0, DFab u1, DFab u2,
                        DFab se0, DFab se1, DFab se2, Scratch w) {
    CellIdx c;
    if (!cell_of(w, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    const double* s = st.comp(e.comp);
    const double* f = fo.comp(e.comp);
    const long so = st.off(c.i, c.j, c.k);
    const long fo_ = fo.off(c.i, c.j, c.k);
    const long wo = woff(w, c.i, c.j, c.k);
    const DFab* vel[3] = {&u0, &u1, &u2};
    const DFab* sed[3] = {&se0, &se1, &se2};
    const int idx[3] = {c.i, c.j, c.k};
    const double dt2 = 0.5 * e.dt, dt4 = 0.25 * e.dt;
    for (int d = 0; d < g.dim; ++d) {
        bool in = true;
        for (int q = 0; q < g.dim; ++q) {
            const int hi = g.vhi[q] + (q == d ? 1 : 0);
            if (idx[q] < g.vlo[q] || idx[q] > hi) in = false;
        }
        if (!in) continue;
        const long sst = d == 0 ? 1 : (d == 1 ? st.sy : st.sz);
        const long fst = d == 0 ? 1 : (d == 1 ? fo.sy : fo.sz);
        const long wsd = d == 0 ? 1 : (d == 1 ? w.sy : w.sz);
        const long wL = wo - wsd;
        const double sL = s[so - sst], sR = s[so];
        const double fl = e.trace ? w.Ipf[d][wL] : f[fo_ - fst];
        const double fr = e.trace ? w.Imf[d][wo] : f[fo_];
        const DFab& Vd = *vel[d];
        const long vdd = d == 0 ? 1 : (d == 1 ? Vd.sy : Vd.sz);
        const long vdo = Vd.off(c.i, c.j, c.k);
        const double ud = Vd.p[vdo];
        double el = w.sl[d][wo], er = w.sr[d][wo];
        // transverse directions in increasing order: the reference subtracts them in that order
        const int nt = g.dim - 1;
        for (int q = 0; q < nt; ++q) {
            const int t = (q == 0) ? (d == 0 ? 1 : 0) : (d == 2 ? 1 : 2);
            const int o = 3 - d - t;  // 3-D only
            const double* H = (g.dim == 3) ? w.tr[t][o] : w.simh[t];
            const long wstt = t == 0 ? 1 : (t == 1 ? w.sy : w.sz);
            const DFab& Vt = *vel[t];
            const long vtt = t == 0 ? 1 : (t == 1 ? Vt.sy : Vt.sz);
            const long vtd = d == 0 ? 1 : (d == 1 ? Vt.sy : Vt.sz);
            const long vR = Vt.off(c.i, c.j, c.k), vL = vR - vtd;
            const double vLp = Vt.p[vL + vtt], vLm = Vt.p[vL], vRp = Vt.p[vR + vtt],
                         vRm = Vt.p[vR];
            const double ht = g.dx[t];
            if (e.is_cons) {
                el = el - (dt2 / ht) * (H[wL + wstt] * vLp - H[wL] * vLm);
                er = er - (dt2 / ht) * (H[wo + wstt] * vRp - H[wo] * vRm);
            } else {
                el = el - (dt4 / ht) * (vLp + vLm) * (H[wL + wstt] - H[wL]);
                er = er - (dt4 / ht) * (vRp + vRm) * (H[wo + wstt] - H[wo]);
            }
        }
        if (e.is_cons) {
            const double hd = g.dx[d];
            el = el - (dt2 / hd) * sL * (ud - Vd.p[vdo - vdd]);
            er = er - (dt2 / hd) * sR * (Vd.p[vdo + vdd] - ud);
        }
        el = el + dt2 * fl;
        er = er + dt2 * fr;
        double val = upwind(ud, e.eps, el, er);
        val = bc_final(idx[d], g.dlo[d], g.dhi[d], e.bclo[d], e.bchi[d], e.is_vel && e.comp == d,
                       sL, sR, el, er, val);
        const DFab& O = *sed[d];
        O.comp(e.comp)[O.off(c.i, c.j, c.k)] = val;
    }
}

}  // namespace

size_t staged_scratch_doubles(const KGeom& g) {
    long n = 1;
    for (int d = 0; d < g.dim; ++d) n *= (g.vhi[d] - g.vlo[d] + 3);
    return (size_t)n * 31;
}

void staged_make_edge_scal(const KGeom& g, const KEdge& e0, const MxEdgeParams* par,
                           const DFab& state, const DFab* sedge, const DFab* umac,
                           const DFab& force, double* scratch, cudaStream_t stream) {
    Scratch w;
    long ncell = 1;
    for (int d = 0; d < 3; ++d) {
        const bool in = d < g.dim;
        w.lo[d] = in ? g.vlo[d] - 1 : 0;
        w.n[d] = in ? g.vhi[d] - g.vlo[d] + 3 : 1;
        ncell *= w.n[d];
    }
    w.sy = w.n[0];
    w.sz = (long)w.n[0] * w.n[1];
    double* p = scratch;
    auto take = [&]() { double* r = p; p += ncell; return r; };
    for (int d = 0; d < 3; ++d) {
        w.Ip[d] = take(); w.Im[d] = take(); w.Ipf[d] = take(); w.Imf[d] = take();
        w.sl[d] = take(); w.sr[d] = take(); w.simh[d] = take();
        for (int t = 0; t < 3; ++t) w.tr[d][t] = (t == d) ? nullptr : take();
    }
    w.divu = take();

    const int nblk = (int)((ncell + TPB - 1) / TPB);
    DFab u2 = g.dim == 3 ? umac[2] : umac[0];
    DFab se2 = g.dim == 3 ? sedge[2] : sedge[0];
    if (g.dim == 3 && e0.is_cons) {
        k_divu<<<nblk, TPB, 0, stream>>>(g, umac[0], umac[1], u2, w);
        mx_count_launch();
    }
    for (int n = 0; n < par->num_comp; ++n) {
        KEdge e = e0;
        e.comp = par->start_scomp + n;
        const int* bc = par->bc + 6 * (par->start_bccomp + n);
        for (int d = 0; d < 3; ++d) {
            e.bclo[d] = bc[d];
            e.bchi[d] = bc[3 + d];
        }
        k_parabola<<<nblk, TPB, 0, stream>>>(g, e, state, umac[0], umac[1], u2, w, 0);
        mx_count_launch();
        if (e.ppm_type != 0 && e.trace) {
            k_parabola<<<nblk, TPB, 0, stream>>>(g, e, force, umac[0], umac[1], u2, w, 1);
            mx_count_launch();
        }
        k_predict<<<nblk, TPB, 0, stream>>>(g, e, state, umac[0], umac[1], u2, w);
        mx_count_launch();
        if (g.dim == 3) {
            k_transverse<<<nblk, TPB, 0, stream>>>(g, e, state, umac[0], umac[1], u2, w);
            mx_count_launch();
        }
        k_edges<<<nblk, TPB, 0, stream>>>(g, e, state, force, umac[0], umac[1], u2, sedge[0],
                                         sedge[1], se2, w);
        mx_count_launch();
    }
}
