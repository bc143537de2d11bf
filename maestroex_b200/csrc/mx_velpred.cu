// mx_velpred.cu — the MAC-velocity predictor either side of the scalar path (SURVEY.md section
// 8f, rank 4): Maestro::MakeUtrans and Maestro::VelPred for one box, 2-D and 3-D, ppm_type 0/1/2,
// planar and spherical, every phys_bc code.
//
//   Source/MaestroMakeUtrans.cpp:6-648     MakeUtrans            -> vel_make_utrans
//   Source/MaestroVelPred.cpp:5-294        VelPred (driver)      -> vel_pred
//   :296-542 (2-D) / :735-1128 (3-D)       VelPredInterface      -> interface stage
//   :1130-1595                             VelPredTransverse     -> transverse stage (3-D)
//   :544-733 (2-D) / :1597-1917 (3-D)      VelPredVelocities     -> velocity stage
//
// Decomposition: the reference builds Ip / Im (or slopes) of every velocity component in every
// direction as MultiFabs first; here a face forms the parabola (or slope) of its two
// neighbouring cells on the fly from the shared reconstruction of mx_device.cuh
// (PPM with is_umac = false: both characteristic integrals use the cell-centred velocity,
// MaestroPPM.cpp:233-251); in VelPred every cell forms each parabola once and scatters Ip / Im
// straight into the interface states of its two faces, so only the interface states (ul, ur,
// uimh) and the six corner-coupled states go through device workspace.  One thread per zone of the grown box,
// i fastest; four launches per VelPred call (three in 2-D), one per direction for MakeUtrans.
// Kernels are instantiated per dimension with the direction / component loops unrolled, so that
// fab descriptors are read from the kernel parameters and the small per-thread arrays live in
// registers.  256^3, ppm_type 1: MakeUtrans 1.25 ms, VelPred 4.6 ms = 190x the reference on 16
// host threads (profiles/r2_velpred.json).
//
// Parity contract as in mx_device.cuh: reference expression trees (association included: the
// 2-D and 3-D slope extrapolations differ, MaestroVelPred.cpp:336-345 vs :785-796), -fmad=false,
// IEEE divisions.  Reference quirks that are kept are marked QUIRK.
#include "mx_common.h"
#include "mx_device.cuh"
#include "mx_velpred.h"

using namespace mxd;

namespace {

// amrex::PhysBCType (AMReX_BC_TYPES.H:61-64)
constexpr int P_INTERIOR = 0, P_INFLOW = 1, P_OUTFLOW = 2, P_SYMMETRY = 3, P_SLIPWALL = 4,
              P_NOSLIPWALL = 5;

struct Box3 {
    int lo[3], hi[3];
};

template <class F>
__global__ void __launch_bounds__(256) k_vbox(Box3 b, F f) {
    const int i = b.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = b.lo[1] + blockIdx.y;
    const int k = b.lo[2] + blockIdx.z;
    if (i <= b.hi[0]) f(i, j, k);
}

// one thread per zone, i fastest.  The block width is chosen per box: a row of 258 zones (a 256^3
// box grown by one) would leave a third of three 128-thread blocks idle, three 96-thread blocks
// lose 10 % (256^3 ran 15 % slower per zone than 250^3 with fixed 128-thread blocks)
template <class F>
void for_box(const Box3& b, cudaStream_t st, F f) {
    const int nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1, nz = b.hi[2] - b.lo[2] + 1;
    if (nx <= 0 || ny <= 0 || nz <= 0) return;
    int bw = 128, best = 1 << 30;
    for (int w = 64; w <= 256; w += 32) {
        const int padded = (nx + w - 1) / w * w;
        if (padded < best || (padded == best && w > bw)) {
            best = padded;
            bw = w;
        }
    }
    dim3 grid((nx + bw - 1) / bw, ny, nz);
    k_vbox<<<grid, bw, 0, st>>>(b, f);
    mx_count_launch();
}

// amrex::grow(mfi.tilebox(), 1) in the problem's dimensions
Box3 grown_box(const KGeom& g) {
    Box3 b;
    for (int d = 0; d < 3; ++d) {
        const int ng = d < g.dim ? 1 : 0;
        b.lo[d] = g.vlo[d] - ng;
        b.hi[d] = g.vhi[d] + ng;
    }
    return b;
}

struct I3 {
    int v[3];
};
MXD I3 mk(int i, int j, int k) { return I3{{i, j, k}}; }
MXD I3 sh(I3 c, int d, int s) {
    c.v[d] += s;
    return c;
}
MXD double& at(const DFab& f, const I3& c, int n = 0) {
    return f.p[f.off(c.v[0], c.v[1], c.v[2]) + (long)n * f.sn];
}
MXD bool inside(const Box3& b, const I3& c) {
    return c.v[0] >= b.lo[0] && c.v[0] <= b.hi[0] && c.v[1] >= b.lo[1] && c.v[1] <= b.hi[1] &&
           c.v[2] >= b.lo[2] && c.v[2] <= b.hi[2];
}

// three fabs captured by value in a kernel lambda: with the direction loops unrolled every
// index is a compile-time constant and the fab is read from the kernel parameters (per-thread
// arrays of fabs cost ~200 B of local-memory stores per thread: that was most of the first
// version's run time, profiles/r2_velpred.json)
struct F3 {
    DFab f[3];
};

// workspace arrays over the grown box
struct WS {
    double* p;
    Box3 b;
    long sy, sz, n;
    MXD long off(const I3& c) const {
        return (long)(c.v[0] - b.lo[0]) + (long)(c.v[1] - b.lo[1]) * sy + (long)(c.v[2] - b.lo[2]) * sz;
    }
    // ul / ur / uimh of direction d, component c; corner-coupled state of component c on d faces
    MXD double& ul(int d, int c, const I3& x) const { return p[(0 + d * 3 + c) * n + off(x)]; }
    MXD double& ur(int d, int c, const I3& x) const { return p[(9 + d * 3 + c) * n + off(x)]; }
    MXD double& imh(int d, int c, const I3& x) const { return p[(18 + d * 3 + c) * n + off(x)]; }
    MXD double& tr(int c, int d, const I3& x) const { return p[(27 + c * 3 + d) * n + off(x)]; }
};
constexpr int WS_ARRAYS = 36;

WS make_ws(const KGeom& g, double* p) {
    WS w;
    w.p = p;
    w.b = grown_box(g);
    w.sy = w.b.hi[0] - w.b.lo[0] + 1;
    w.sz = w.sy * (w.b.hi[1] - w.b.lo[1] + 1);
    w.n = w.sz * (w.b.hi[2] - w.b.lo[2] + 1);
    return w;
}

// x / h_d, correctly rounded (== the IEEE division of the reference): the sequence of div_h in
// mx_stream.cuh (tests/test_divconst.py), ~5 fp64 instructions instead of a software division
MXD double div_hd(const KVel& e, int d, double x) {
    const double y = e.inv_h[d];
    double q = x * y;
    if (!e.inv_exact[d]) {
        const double h = e.h[d];
        double r = fma_(-h, q, x);
        q = fma_(r, y, q);
        r = fma_(-h, q, x);
        q = fma_(r, y, q);
    }
    return q;
}

// Slopex/y/z (ppm_type 0: `ip` is the slope) or PPM with is_umac = false of component n of `s`
// in direction d at cell c; bcs_u[n] decides the one-sided rows (MaestroSlopes.cpp:17-22,
// MaestroPPM.cpp:30-31)
struct RC {
    double ip, im;
};
MXD RC recon(const KGeom& g, const KVel& e, const DFab& s, const DFab& uf, int n, int d, const I3& c) {
    const long st = d == 0 ? 1 : (d == 1 ? s.sy : s.sz);
    const double* sp_ = s.comp(n) + s.off(c.v[0], c.v[1], c.v[2]);
    const bool lobc = e.bclo[n][d] == BC_EXT_DIR || e.bclo[n][d] == BC_HOEXTRAP;
    const bool hibc = e.bchi[n][d] == BC_EXT_DIR || e.bchi[n][d] == BC_HOEXTRAP;
    double S[7];
    S[0] = S[6] = 0.0;
#pragma unroll
    for (int q = -2; q <= 2; ++q) S[q + 3] = sp_[q * st];
    RC r;
    if (e.ppm_type == 0) {
        r.ip = slope_any(e.slope_order, S + 1, c.v[d], g.dlo[d], g.dhi[d], lobc, hibc, d == 0);
        r.im = 0.0;
        return r;
    }
    double sm, sp;
    if (e.ppm_type == 1) {
        ppm1_parabola<true>(S, c.v[d], g.dlo[d], g.dhi[d], lobc, hibc, sm, sp);
    } else {
        S[0] = sp_[-3 * st];
        S[6] = sp_[3 * st];
        ppm2_parabola<true>(S, c.v[d], g.dlo[d], g.dhi[d], lobc, hibc, d == 0, sm, sp);
    }
    const double u = at(uf, c, d);
    const double s0 = S[3];
    const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
    const double sig = div_hd(e, d, fabs(u) * e.dt);
    r.ip = ppm_Ip(sm, sp, s0, s6, u, sig, e.eps);
    r.im = ppm_Im(sm, sp, s0, s6, u, sig, e.eps);
    return r;
}

// the w0 term of the Riemann problems: w0mac on spherical grids, w0_cart's last component in
// the last direction of planar ones, nothing otherwise (MaestroVelPred.cpp:1651-1667, :1860-1884)
template <int DIM>
MXD double w0_term(const KVel& e, const F3& w0mac, const DFab& w0c, int d, const I3& f) {
    if (e.spherical) return at(w0mac.f[d], f);
    if (d == DIM - 1) return at(w0c, f, DIM - 1);
    return 0.0;
}

// "solve Riemann problem using full velocity"
MXD double riemann_full(double l, double r, double w, double eps) {
    const bool test = (l + w <= 0.0 && r + w >= 0.0) || (fabs(l + r + 2.0 * w) < eps);
    const double v = 0.5 * (l + r) + w > 0.0 ? l : r;
    return test ? 0.0 : v;
}

// "upwind using full velocity": trans > 0 ? l : r, average where |trans| < rel_eps
MXD double upwind_trans(double t, double l, double r, double eps) {
    const double v = t > 0.0 ? l : r;
    return fabs(t) < eps ? 0.5 * (l + r) : v;
}

}  // namespace

size_t vel_pred_workspace(const KGeom& g) {
    const WS w = make_ws(g, nullptr);
    return (size_t)w.n * WS_ARRAYS;
}

// ---- Maestro::MakeUtrans -------------------------------------------------------------------
template <int DIM, int D>
static void utrans_dir(const KGeom& g, const KVel& e, const DFab& utilde, const DFab& ufull,
                       const DFab& out, const DFab& w0c, const DFab& wmac_d, cudaStream_t st) {
    Box3 fb;
    for (int q = 0; q < 3; ++q) {
        fb.lo[q] = g.vlo[q];
        fb.hi[q] = g.vhi[q] + (q == D ? 1 : 0);
    }
    for_box(fb, st, [=] __device__(int i, int j, int k) {
        constexpr int d = D;
        const I3 f = mk(i, j, k), cl = sh(f, d, -1);
        const double dt2 = 0.5 * e.dt;
        double l, r;
        const RC L = recon(g, e, utilde, ufull, d, d, cl), R = recon(g, e, utilde, ufull, d, d, f);
        if (e.ppm_type == 0) {
            // MaestroMakeUtrans.cpp:94-103
            l = at(utilde, cl, d) + (0.5 - (dt2 / g.dx[d]) * dmax_(0.0, at(ufull, cl, d))) * L.ip;
            r = at(utilde, f, d) - (0.5 + (dt2 / g.dx[d]) * dmin_(0.0, at(ufull, f, d))) * R.ip;
        } else {
            l = L.ip;
            r = R.im;
        }
        const int I = f.v[d];
        if (I == g.dlo[d]) {
            const int bc = e.pbc_lo[d];
            if (bc == P_INFLOW) {
                l = r = at(utilde, cl, d);
            } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                // QUIRK: the 2-D y sweep zeroes only the right state (:220-221)
                if (!(DIM == 2 && d == 1)) l = 0.0;
                r = 0.0;
            } else if (bc == P_OUTFLOW) {
                l = dmin_(r, 0.0);
                r = l;
            }
        } else if (I == g.dhi[d] + 1) {
            const int bc = e.pbc_hi[d];
            if (bc == P_INFLOW) {
                // QUIRK: 3-D reads one cell further out (:350-351, :466-467, :581-582)
                const I3 src = DIM == 3 ? sh(f, d, 1) : f;
                l = r = at(utilde, src, d);
            } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                l = 0.0;
                r = 0.0;
            } else if (bc == P_OUTFLOW) {
                l = dmax_(l, 0.0);
                r = l;
            }
        }
        double w = 0.0;  // w0_term
        if (e.spherical)
            w = at(wmac_d, f);
        else if (d == DIM - 1)
            w = at(w0c, f, DIM - 1);
        at(out, f) = riemann_full(l, r, w, e.eps);
    });
}

void vel_make_utrans(const KGeom& g, const KVel& e, const DFab& utilde, const DFab& ufull,
                     const DFab* utrans, const DFab& w0c, const DFab* w0mac_in, cudaStream_t st) {
    DFab wm[3];
    for (int d = 0; d < 3; ++d) wm[d] = (e.spherical && w0mac_in) ? w0mac_in[d] : w0c;
    if (g.dim == 2) {
        utrans_dir<2, 0>(g, e, utilde, ufull, utrans[0], w0c, wm[0], st);
        utrans_dir<2, 1>(g, e, utilde, ufull, utrans[1], w0c, wm[1], st);
    } else {
        utrans_dir<3, 0>(g, e, utilde, ufull, utrans[0], w0c, wm[0], st);
        utrans_dir<3, 1>(g, e, utilde, ufull, utrans[1], w0c, wm[1], st);
        utrans_dir<3, 2>(g, e, utilde, ufull, utrans[2], w0c, wm[2], st);
    }
}

// ---- Maestro::VelPred ------------------------------------------------------------------------
template <int DIM>
static void vel_pred_dim(const KGeom& g, const KVel& e, const DFab& utilde, const DFab& ufull,
                         const DFab* utrans_in, const DFab* umac_out, const DFab* w0mac_in,
                         const DFab& w0c, const DFab& force, double* wsp, cudaStream_t st) {
    const WS W = make_ws(g, wsp);
    const Box3 obx = W.b;
    constexpr int dim = DIM;
    F3 utr, um, wmac;
    for (int d = 0; d < 3; ++d) {
        utr.f[d] = utrans_in[d < DIM ? d : DIM - 1];
        um.f[d] = umac_out[d < DIM ? d : DIM - 1];
        wmac.f[d] = (e.spherical && w0mac_in) ? w0mac_in[d] : w0c;
    }
    Box3 vb;
    for (int q = 0; q < 3; ++q) {
        vb.lo[q] = g.vlo[q];
        vb.hi[q] = g.vhi[q];
    }

    // ---- VelPredInterface, first half: every cell forms the parabola (or slope) of each component
    // in each direction ONCE and scatters the two traces: Im is the right state of the cell's low
    // face, Ip the left state of its high face (ppm_type 0: the two slope extrapolations) ------
    for_box(obx, st, [=] __device__(int i, int j, int k) {
        const I3 c = mk(i, j, k);
        const double dt2 = 0.5 * e.dt;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
            const double h = g.dx[d];
            double fl = 0.0, fr = 0.0;
            if (e.ppm_type == 0) {
                const double uC = at(ufull, c, d);
                const double uL = dmax_(0.0, uC), uR = dmin_(0.0, uC);
                // 3-D: :785-796; 2-D association differs: :336-345
                fl = dim == 3 ? 0.5 - dt2 * uL / h : 0.5 - (dt2 / h) * uL;
                fr = dim == 3 ? 0.5 + dt2 * uR / h : 0.5 + (dt2 / h) * uR;
            }
            const bool has_hi = c.v[d] < obx.hi[d];
            const I3 fh = sh(c, d, 1);
#pragma unroll
            for (int n = 0; n < dim; ++n) {
                const RC r = recon(g, e, utilde, ufull, n, d, c);
                double vl, vr;
                if (e.ppm_type == 0) {
                    const double s0 = at(utilde, c, n);
                    vl = s0 + fl * r.ip;
                    vr = s0 - fr * r.ip;
                } else {
                    vl = r.ip;
                    vr = r.im;
                }
                W.ur(d, n, c) = vr;
                if (has_hi) W.ul(d, n, fh) = vl;
            }
        }
    });

    // ---- VelPredInterface, second half: phys_bc on the faces of mxbx / mybx / mzbx (obx with the
    // low side of the normal direction pulled in) and the transverse components of uimh ---------
    for_box(obx, st, [=] __device__(int i, int j, int k) {
        const I3 f = mk(i, j, k);
#pragma unroll
        for (int d = 0; d < dim; ++d) {
            if (f.v[d] < obx.lo[d] + 1) continue;
            const I3 cl = sh(f, d, -1);
            double ul[3], ur[3];
#pragma unroll
            for (int n = 0; n < dim; ++n) {
                ul[n] = W.ul(d, n, f);
                ur[n] = W.ur(d, n, f);
            }
            const int I = f.v[d];
            const bool on_lo = I == g.dlo[d], on_hi = I == g.dhi[d] + 1;
            if (on_lo) {
                const int bc = e.pbc_lo[d];
                if (bc == P_INFLOW) {
#pragma unroll
                    for (int n = 0; n < dim; ++n) ul[n] = ur[n] = at(utilde, cl, n);
                } else if (bc == P_SLIPWALL || bc == P_SYMMETRY) {
#pragma unroll
                    for (int n = 0; n < dim; ++n) {
                        if (n == d) {
                            ul[n] = 0.0;
                            ur[n] = 0.0;
                        } else {
                            ul[n] = ur[n];
                        }
                    }
                } else if (bc == P_NOSLIPWALL) {
#pragma unroll
                    for (int n = 0; n < dim; ++n) ul[n] = ur[n] = 0.0;
                } else if (bc == P_OUTFLOW) {
                    ur[d] = dmin_(ur[d], 0.0);
                    if (dim == 2 && d == 0) {
                        // QUIRK: the 2-D x sweep copies left over right here (:380-385)
#pragma unroll
                        for (int n = 0; n < dim; ++n) ur[n] = ul[n];
                    } else {
#pragma unroll
                        for (int n = 0; n < dim; ++n) ul[n] = ur[n];
                    }
                }
            } else if (on_hi) {
                const int bc = e.pbc_hi[d];
                if (bc == P_INFLOW) {
#pragma unroll
                    for (int n = 0; n < dim; ++n) ul[n] = ur[n] = at(utilde, f, n);
                } else if (bc == P_SLIPWALL || bc == P_SYMMETRY) {
#pragma unroll
                    for (int n = 0; n < dim; ++n) {
                        if (n == d) {
                            ul[n] = 0.0;
                            ur[n] = 0.0;
                        } else {
                            ur[n] = ul[n];
                        }
                    }
                } else if (bc == P_NOSLIPWALL) {
#pragma unroll
                    for (int n = 0; n < dim; ++n) ul[n] = ur[n] = 0.0;
                } else if (bc == P_OUTFLOW) {
                    ul[d] = dmax_(ul[d], 0.0);
#pragma unroll
                    for (int n = 0; n < dim; ++n) ur[n] = ul[n];
                }
            }
            const double t = at(utr.f[d], f);
#pragma unroll
            for (int n = 0; n < dim; ++n) {
                if (on_lo || on_hi) {  // only boundary faces change
                    W.ul(d, n, f) = ul[n];
                    W.ur(d, n, f) = ur[n];
                }
                if (n != d) W.imh(d, n, f) = upwind_trans(t, ul[n], ur[n], e.eps);
            }
        }
    });

    // ---- VelPredTransverse (3-D): component c on d faces, corrected by the e derivative ----
    if constexpr (DIM == 3) {
        for_box(obx, st, [=] __device__(int i, int j, int k) {
                const I3 f = mk(i, j, k);
            const double dt6 = e.dt / 6.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    if (d == c) continue;
                    const int t = 3 - c - d;  // the transverse direction
                    // amrex::grow(tilebox, c, 1) with growHi(d, 1)
                    Box3 b = vb;
                    b.lo[c] -= 1;
                    b.hi[c] += 1;
                    b.hi[d] += 1;
                    if (!inside(b, f)) continue;
                    const I3 cl = sh(f, d, -1);
                    const double ht = g.dx[t];
                    double l = W.ul(d, c, f) - (dt6 / ht) * (at(utr.f[t], sh(cl, t, 1)) + at(utr.f[t], cl)) *
                                                  (W.imh(t, c, sh(cl, t, 1)) - W.imh(t, c, cl));
                    double r = W.ur(d, c, f) - (dt6 / ht) * (at(utr.f[t], sh(f, t, 1)) + at(utr.f[t], f)) *
                                                  (W.imh(t, c, sh(f, t, 1)) - W.imh(t, c, f));
                    const int I = f.v[d];
                    if (I == g.dlo[d]) {
                        const int bc = e.pbc_lo[d];
                        if (bc == P_INFLOW) {
                            l = r = at(utilde, cl, c);
                        } else if (bc == P_SLIPWALL || bc == P_SYMMETRY || bc == P_OUTFLOW) {
                            l = r;
                        } else if (bc == P_NOSLIPWALL) {
                            l = r = 0.0;
                        }
                    } else if (I == g.dhi[d] + 1) {
                        const int bc = e.pbc_hi[d];
                        if (bc == P_INFLOW) {
                            l = r = at(utilde, f, c);
                        } else if (bc == P_SLIPWALL || bc == P_SYMMETRY || bc == P_OUTFLOW) {
                            r = l;
                        } else if (bc == P_NOSLIPWALL) {
                            l = r = 0.0;
                        }
                    }
                    W.tr(c, d, f) = upwind_trans(at(utr.f[d], f), l, r, e.eps);
                }
            }
        });
    }

    // ---- VelPredVelocities: the normal component on the faces of the valid box ---------------
    Box3 fb = vb;
    for (int d = 0; d < dim; ++d) fb.hi[d] += 1;
    for_box(fb, st, [=] __device__(int i, int j, int k) {
        const I3 f = mk(i, j, k);
        const double dt2 = 0.5 * e.dt, dt4 = 0.25 * e.dt;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
            // mfi.nodaltilebox(d)
            bool in = true;
#pragma unroll
            for (int q = 0; q < dim; ++q)
                if (q != d && f.v[q] > vb.hi[q]) in = false;
            if (!in) continue;
            const I3 cl = sh(f, d, -1);
            double fl, fr;
            if (e.trace == 0) {
                fl = at(force, cl, d);
                fr = at(force, f, d);
            } else {
                fl = recon(g, e, force, ufull, d, d, cl).ip;
                fr = recon(g, e, force, ufull, d, d, f).im;
            }
            double l = W.ul(d, d, f), r = W.ur(d, d, f);
#pragma unroll
            for (int t = 0; t < dim; ++t) {  // transverse directions in ascending order
                if (t == d) continue;
                const double ht = g.dx[t];
                double dl, dr;
                if (dim == 3) {
                    dl = W.tr(d, t, sh(cl, t, 1)) - W.tr(d, t, cl);
                    dr = W.tr(d, t, sh(f, t, 1)) - W.tr(d, t, f);
                } else {
                    dl = W.imh(t, d, sh(cl, t, 1)) - W.imh(t, d, cl);
                    dr = W.imh(t, d, sh(f, t, 1)) - W.imh(t, d, f);
                }
                l = l - (dt4 / ht) * (at(utr.f[t], sh(cl, t, 1)) + at(utr.f[t], cl)) * dl;
                r = r - (dt4 / ht) * (at(utr.f[t], sh(f, t, 1)) + at(utr.f[t], f)) * dr;
            }
            l = l + dt2 * fl;
            r = r + dt2 * fr;
            double v = riemann_full(l, r, w0_term<DIM>(e, wmac, w0c, d, f), e.eps);
            const int I = f.v[d];
            if (I == g.dlo[d]) {
                const int bc = e.pbc_lo[d];
                if (bc == P_INFLOW) {
                    v = at(utilde, cl, d);
                } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                    v = 0.0;
                } else if (bc == P_OUTFLOW) {
                    v = dmin_(r, 0.0);
                }
            } else if (I == g.dhi[d] + 1) {
                const int bc = e.pbc_hi[d];
                if (bc == P_INFLOW) {
                    v = at(utilde, f, d);
                } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                    v = 0.0;
                } else if (bc == P_OUTFLOW) {
                    v = dmax_(l, 0.0);
                }
            }
            at(um.f[d], f) = v;
        }
    });
}

void vel_pred(const KGeom& g, const KVel& e, const DFab& utilde, const DFab& ufull,
              const DFab* utrans, const DFab* umac, const DFab* w0mac, const DFab& w0c,
              const DFab& force, double* ws, cudaStream_t st) {
    if (g.dim == 2)
        vel_pred_dim<2>(g, e, utilde, ufull, utrans, umac, w0mac, w0c, force, ws, st);
    else
        vel_pred_dim<3>(g, e, utilde, ufull, utrans, umac, w0mac, w0c, force, ws, st);
}
