// mx_device.cuh — fp64 device arithmetic of the Godunov edge-state predictor, written once
// and direction-generic, shared by the staged (mx_staged.cu) and fused (mx_fused.cu) kernels.
//
// Parity contract: every function reproduces the reference expression tree (association
// included) so that with FMA contraction off (-fmad=false; the reference CPU build has no
// FMA either) results are bit-identical to
//   Source/MaestroSlopes.cpp, Source/MaestroPPM.cpp, Source/MaestroMakeEdgeScalars.cpp.
// Deliberate, result-preserving departures from the literal source:
//   * x / 6.0 and x / 3.0 use a 3-instruction correctly-rounded sequence (div_c) instead of
//     the ~25-instruction IEEE division routine; explicit __fma_rn, independent of -fmad;
//   * copysign(1.0, a) * m with m >= 0 is formed as copysign(m, a) on the integer pipe;
//   * |x| uses the free operand modifier; min/max follow std::min/std::max (amrex::min/max),
//     not fmin/fmax, so ties and signed zeros select the same operand as the reference.
#ifndef MX_DEVICE_CUH
#define MX_DEVICE_CUH

#include <cuda_runtime.h>

#define MXD __device__ __forceinline__

namespace mxd {

constexpr int BC_REFLECT_ODD = -1;
constexpr int BC_REFLECT_EVEN = 1;
constexpr int BC_FOEXTRAP = 2;
constexpr int BC_EXT_DIR = 3;
constexpr int BC_HOEXTRAP = 4;

MXD double dmin_(double a, double b) { return (b < a) ? b : a; }
MXD double dmax_(double a, double b) { return (a < b) ? b : a; }

// select / fused multiply-add / multiply spelled as functions so that the reconstruction
// templates below read the same for any value type
MXD double sel(bool m, double x, double y) { return m ? x : y; }
MXD double fma_(double x, double y, double z) { return __fma_rn(x, y, z); }
MXD double mul_(double x, double y) { return __dmul_rn(x, y); }

// copysign(1.0, sgn) * mag for mag >= +0 (exact: +-1 * mag only flips the sign bit)
MXD double signed_mag(double mag, double sgn) {
    return __hiloint2double((__double2hiint(mag) & 0x7fffffff) | (__double2hiint(sgn) & 0x80000000),
                            __double2loint(mag));
}
MXD double sign1(double sgn) { return signed_mag(1.0, sgn); }

// (prod > 0.0) ? r : +0.0 as a select.  Written in PTX because nvcc otherwise sinks the
// computation of `r` under a divergent branch (BSSY / BRA / BSYNC per limiter: measured slower
// than computing it unconditionally on every lane).
MXD double sel_pos(double prod, double r) {
    double out;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, 0d0000000000000000;\n\t"
        "selp.f64 %0, %1, 0d0000000000000000, p;\n\t}"
        : "=d"(out)
        : "d"(r), "d"(prod));
    return out;
}


// Correctly rounded x / c for the compile-time constants c = 6 and c = 3 (Markstein-style:
// q0 = RN(x * RN(1/c)), r = x - c*q0 exactly via FMA, q = RN(q0 + r * RN(1/c))).  Verified
// against IEEE division on 5.5e8 random and structured operands per divisor
// (tests/test_divconst.py restates the check on the host); no over/underflow in range.
template <class T>
MXD T div6(T x) {
    const double y = 1.0 / 6.0;
    const T q = mul_(x, y);
    const T r = fma_(-6.0, q, x);
    return fma_(r, y, q);
}
MXD double div3(double x) {
    const double y = 1.0 / 3.0;
    const double q = __dmul_rn(x, y);
    const double r = __fma_rn(-3.0, q, x);
    return __fma_rn(r, y, q);
}

// ---------------------------------------------------------------------------------------
// Slopes (ppm_type == 0): Source/MaestroSlopes.cpp:6-601.  S[o+2] = s at offset o in [-2,2]
// along the sweep direction; I = cell index along that direction; lo/hi = domain bounds;
// lobc/hibc = (bc == ext_dir || bc == hoextrap); xquirk = sweep direction is x.
// ---------------------------------------------------------------------------------------
MXD double mc_lim(double dpls, double dmn) {
    const double slim = dmin_(fabs(dpls), fabs(dmn));
    return sel_pos(dpls * dmn, slim);
}

MXD double slope2(const double* S, int I, int lo, int hi, bool lobc, bool hibc) {
    // MaestroSlopes.cpp:43-90
    const double sm1 = S[1], s0 = S[2], sp1 = S[3];
    double del = 0.5 * (sp1 - sm1);
    const double dpls = 2.0 * (sp1 - s0);
    const double dmn = 2.0 * (s0 - sm1);
    const double slim = mc_lim(dpls, dmn);
    double r = signed_mag(dmin_(slim, fabs(del)), del);
    if (lobc) {
        if (I == lo - 1) {
            r = 0.0;
        } else if (I == lo) {
            del = div3(sp1 + 3.0 * s0 - 4.0 * sm1);
            r = signed_mag(dmin_(slim, fabs(del)), del);
        }
    }
    if (hibc) {
        if (I == hi + 1) {
            r = 0.0;
        } else if (I == hi) {
            del = -div3(sm1 + 3.0 * s0 - 4.0 * sp1);
            r = signed_mag(dmin_(slim, fabs(del)), del);
        }
    }
    return r;
}

MXD double slope4(const double* S, int I, int lo, int hi, bool lobc, bool hibc, bool xquirk) {
    // MaestroSlopes.cpp:95-201
    const double sm2 = S[0], sm1 = S[1], s0 = S[2], sp1 = S[3], sp2 = S[4];
    // left
    double dcen = 0.5 * (s0 - sm2);
    double dmn = 2.0 * (sm1 - sm2);
    double dpls = 2.0 * (s0 - sm1);
    double dlim = mc_lim(dpls, dmn);
    double dxl = signed_mag(dmin_(dlim, fabs(dcen)), dcen);
    // right
    dcen = 0.5 * (sp2 - s0);
    dmn = 2.0 * (sp1 - s0);
    dpls = 2.0 * (sp2 - sp1);
    dlim = mc_lim(dpls, dmn);
    double dxr = signed_mag(dmin_(dlim, fabs(dcen)), dcen);
    // centre
    dcen = 0.5 * (sp1 - sm1);
    dmn = 2.0 * (s0 - sm1);
    dpls = 2.0 * (sp1 - s0);
    dlim = mc_lim(dpls, dmn);
    double ds = 4.0 / 3.0 * dcen - div6(dxr + dxl);
    double r = signed_mag(dmin_(fabs(ds), dlim), dcen);

    if (lobc || hibc) {
        // one-sided rows.  QUIRK (MaestroSlopes.cpp:142,156,178,192 vs :341.. and :539..):
        // Slopex keeps the centred `dlim` where Slopey/z use the freshly computed `slim`.
        if (lobc) {
            if (I == lo - 1) {
                r = 0.0;
            } else if (I == lo) {
                const double del = -16.0 / 15.0 * sm1 + 0.5 * s0 + 2.0 / 3.0 * sp1 - 0.1 * sp2;
                const double a = 2.0 * (s0 - sm1), b = 2.0 * (sp1 - s0);
                double slim = dmin_(fabs(b), fabs(a));
                slim = b * a > 0.0 ? (xquirk ? dlim : slim) : 0.0;
                r = signed_mag(dmin_(slim, fabs(del)), del);
            } else if (I == lo + 1) {
                const double del = -16.0 / 15.0 * sm2 + 0.5 * sm1 + 2.0 / 3.0 * s0 - 0.1 * sp1;
                const double a = 2.0 * (sm1 - sm2), b = 2.0 * (s0 - sm1);
                double slim = dmin_(fabs(b), fabs(a));
                slim = b * a > 0.0 ? (xquirk ? dlim : slim) : 0.0;
                dxl = signed_mag(dmin_(slim, fabs(del)), del);
                ds = 4.0 / 3.0 * dcen - div6(dxr + dxl);
                r = signed_mag(dmin_(fabs(ds), dlim), dcen);
            }
        }
        if (hibc) {
            if (I == hi + 1) {
                r = 0.0;
            } else if (I == hi) {
                const double del = -(-16.0 / 15.0 * sp1 + 0.5 * s0 + 2.0 / 3.0 * sm1 - 0.1 * sm2);
                const double a = 2.0 * (s0 - sm1), b = 2.0 * (sp1 - s0);
                double slim = dmin_(fabs(b), fabs(a));
                slim = b * a > 0.0 ? (xquirk ? dlim : slim) : 0.0;
                r = signed_mag(dmin_(slim, fabs(del)), del);
            } else if (I == hi - 1) {
                const double del = -(-16.0 / 15.0 * sp2 + 0.5 * sp1 + 2.0 / 3.0 * s0 - 0.1 * sm1);
                const double a = 2.0 * (sp1 - s0), b = 2.0 * (sp2 - sp1);
                double slim = dmin_(fabs(b), fabs(a));
                slim = b * a > 0.0 ? (xquirk ? dlim : slim) : 0.0;
                dxr = signed_mag(dmin_(slim, fabs(del)), del);
                ds = 4.0 / 3.0 * dcen - div6(dxl + dxr);
                r = signed_mag(dmin_(fabs(ds), dlim), dcen);
            }
        }
    }
    return r;
}

MXD double slope_any(int order, const double* S, int I, int lo, int hi, bool lobc, bool hibc,
                     bool xquirk) {
    if (order == 0) return 0.0;
    if (order == 2) return slope2(S, I, lo, hi, lobc, hibc);
    return slope4(S, I, lo, hi, lobc, hibc, xquirk);
}

// ---------------------------------------------------------------------------------------
// PPM: Source/MaestroPPM.cpp
// ---------------------------------------------------------------------------------------
// |x| on the integer pipe (clears the sign bit of the high word; the fp64 pipe is the scarce one)
MXD double abs_(double x) {
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}

// van Leer slope (MaestroPPM.cpp:42-49): with dsc = (sp1 - sm1)/2, dsl = 2 (s0 - sm1),
// dsr = 2 (sp1 - s0) the reference forms sign(dsc) min(|dsc|, |dsl|, |dsr|) if dsl dsr > 0, else 0.
// Scaling by 2 and 1/2 is exact, so the same result is obtained from the raw differences
// c, l, r: the minimum over |c|/2, 2|l|, 2|r| is decided by comparing |c| with 4 min(|l|, |r|)
// and scaled once, and dsl dsr > 0 <=> l r > 0.  This saves a third of the fp64 instructions
// of a slope and lets neighbouring slopes share l / r.  (Equivalent bit for bit unless an
// intermediate over- or underflows, i.e. for differences beyond 1e153 or below 1e-154.)
template <class T>
MXD T vanleer(T sm1, T s0, T sp1) {
    const T c = sp1 - sm1, l = s0 - sm1, r = sp1 - s0;
    const T al = abs_(l), ar = abs_(r), ac = abs_(c);
    const T mn = dmin_(al, ar);        // dmin_(|dsl|, |dsr|) / 2, same tie rule
    const auto inner = 4.0 * mn < ac;  // min(|dsl|, |dsr|) < |dsc|
    const T m = sel(inner, mn, ac) * sel(inner, 2.0, 0.5);
    return sel_pos(l * r, signed_mag(m, c));
}

// 4th-order edge between sL and sR with van Leer slopes dL, dR, clamped (MaestroPPM.cpp:62-67)
template <class T>
MXD T vl_edge(T sL, T sR, T dL, T dR) {
    T e = 0.5 * (sR + sL) - div6(dR - dL);
    // min(sR, sL) and max(sR, sL) from one comparison (for sR == sL either operand is the value)
    const auto p = sL < sR;
    const T lo = sel(p, sL, sR), hi = sel(p, sR, sL);
    e = dmax_(e, lo);
    e = dmin_(e, hi);
    return e;
}

template <class T>
MXD void cw_limit(T s0, T& sm, T& sp) {
    // MaestroPPM.cpp:106-115, written with selects (both arms are cheap and side-effect free)
    const T dp = sp - s0, dm = sm - s0;
    const auto flat = dp * dm >= 0.0;  // dp (s0 - sm) <= 0: s0 - sm = -dm exactly
    const T adp = abs_(dp), adm = abs_(dm);
    const auto c1 = adp >= 2.0 * adm;
    const auto c2 = adm >= 2.0 * adp;
    const T s3 = 3.0 * s0;
    const T sp_c1 = s3 - 2.0 * sm;
    const T sm_c2 = s3 - 2.0 * sp;
    const T sp_n = sel(c1, sp_c1, sp);
    const T sm_n = sel(!c1 && c2, sm_c2, sm);
    sp = sel(flat, s0, sp_n);
    sm = sel(flat, s0, sm_n);
}

template <class T>
MXD T bc_edge(T a, T b, T c, T d, T m0, T m1) {
    // MaestroPPM.cpp:124-131
    T v = -0.2 * a + 0.75 * b + 0.5 * c - 0.05 * d;
    v = dmax_(v, dmin_(m0, m1));
    v = dmin_(v, dmax_(m0, m1));
    return v;
}

// ppm_type 1, second half: CW84 limiter + BC rows applied to the unlimited edge values
// sm_u = edge(i-1/2), sp_u = edge(i+1/2) (MaestroPPM.cpp:100-205).  S[o+3], o in [-2,2] read.
template <bool BC, class T>
MXD void ppm1_finish(const T* S, int I, int lo, int hi, bool lobc, bool hibc, T sm_u, T sp_u,
                     T& sm, T& sp) {
    const T sm2 = S[1], sm1 = S[2], s0 = S[3], sp1 = S[4], sp2 = S[5];
    sm = sm_u;
    sp = sp_u;
    const T sedgel = sp, sedger = sm;
    cw_limit(s0, sm, sp);
    // (lobc / hibc are uniform over a warp in the streaming kernel, which clears them away
    // from the domain faces: the whole block below is then skipped by a uniform branch)
    if (BC && (lobc || hibc)) {
        if (I == lo) {
            if (lobc) {
                sm = sm1;
                sp = bc_edge(sm1, s0, sp1, sp2, sp1, s0);
            }
        } else if (I == lo + 1) {
            if (lobc) {
                sm = bc_edge(sm2, sm1, s0, sp1, s0, sm1);
                sp = sedgel;
                cw_limit(s0, sm, sp);
            }
        } else if (I == hi) {
            if (hibc) {
                sp = sp1;
                sm = bc_edge(sp1, s0, sm1, sm2, sm1, s0);
            }
        } else if (I == hi - 1) {
            if (hibc) {
                sp = bc_edge(sp2, sp1, s0, sm1, s0, sp1);
                sm = sedger;
                cw_limit(s0, sm, sp);
            }
        }
    }
}

// ppm_type 1: sm, sp of the limited parabola in the cell.  S[o+3], o in [-3,3] (only -2..2
// read).  MaestroPPM.cpp:34-205.
template <bool BC, class T>
MXD void ppm1_parabola(const T* S, int I, int lo, int hi, bool lobc, bool hibc, T& sm, T& sp) {
    const T sm2 = S[1], sm1 = S[2], s0 = S[3], sp1 = S[4], sp2 = S[5];
    const T dl = vanleer(sm2, sm1, s0);
    const T dc = vanleer(sm1, s0, sp1);
    const T dr = vanleer(s0, sp1, sp2);
    ppm1_finish<BC>(S, I, lo, hi, lobc, hibc, vl_edge(sm1, s0, dl, dc), vl_edge(s0, sp1, dc, dr),
                    sm, sp);
}

// ppm_type 2 edge (MaestroPPM.cpp:258-279): a,b,c,d = the four cells around the edge b|c
MXD double cs_edge(double a, double b, double c, double d) {
    const double C = 1.25;
    double e = (7.0 / 12.0) * (b + c) - (1.0 / 12.0) * (a + d);
    if ((e - b) * (c - e) < 0.0) {
        const double D2 = 3.0 * (b - 2.0 * e + c);
        const double D2L = a - 2.0 * b + c;
        const double D2R = b - 2.0 * c + d;
        const double sgn = sign1(D2);
        const double D2LIM = sgn * dmax_(dmin_(C * sgn * D2L, dmin_(C * sgn * D2R, sgn * D2)), 0.0);
        e = 0.5 * (b + c) - div6(D2LIM);
    }
    return e;
}

// Colella-Sekora limiter (MaestroPPM.cpp:356-441)
MXD void cs_limit(double sedgel, double sedge, double sedger, double sedgerr, double sm2,
                  double sm1, double s0, double sp1, double sp2, double bigm_thresh, double& sm,
                  double& sp) {
    const double C = 1.25;
    double alphap = sedger - s0;
    double alpham = sedge - s0;
    const bool bigp = fabs(alphap) > 2.0 * fabs(alpham);
    const bool bigm = fabs(alpham) > 2.0 * fabs(alphap);
    bool extremum = false;
    if (alpham * alphap >= 0.0) {
        extremum = true;
    } else if (bigp || bigm) {
        const double dafacem = sedge - sedgel;
        const double dafacep = sedgerr - sedger;
        const double dabarm = s0 - sm1;
        const double dabarp = sp1 - s0;
        const double dafacemin = dmin_(fabs(dafacem), fabs(dafacep));
        const double dabarmin = dmin_(fabs(dabarm), fabs(dabarp));
        double dachkm, dachkp;
        if (dafacemin >= dabarmin) {
            dachkm = dafacem;
            dachkp = dafacep;
        } else {
            dachkm = dabarm;
            dachkp = dabarp;
        }
        extremum = (dachkm * dachkp <= 0.0);
    }
    if (extremum) {
        const double D2 = 6.0 * (alpham + alphap);
        const double D2L = sm2 - 2.0 * sm1 + s0;
        const double D2R = s0 - 2.0 * sp1 + sp2;
        const double D2C = sm1 - 2.0 * s0 + sp1;
        const double sgn = sign1(D2);
        const double D2LIM = dmax_(
            dmin_(sgn * D2, dmin_(C * sgn * D2L, dmin_(C * sgn * D2R, C * sgn * D2C))), 0.0);
        const double D2ABS = dmax_(fabs(D2), 1.e-10);
        alpham = alpham * D2LIM / D2ABS;
        alphap = alphap * D2LIM / D2ABS;
    } else {
        if (bigp) {
            const double sgn = sign1(alpham);
            const double amax = -alphap * alphap / (4.0 * (alpham + alphap));
            const double delam = sm1 - s0;
            if (sgn * amax >= sgn * delam) {
                if (sgn * (delam - alpham) >= 1.e-10) {
                    alphap = (-2.0 * delam - 2.0 * sgn * sqrt(delam * delam - delam * alpham));
                } else {
                    alphap = -2.0 * alpham;
                }
            }
        }
        if (bigm) {
            const double sgn = sign1(alphap);
            const double amax = -alpham * alpham / (4.0 * (alpham + alphap));
            const double delap = sp1 - s0;
            if (sgn * amax >= sgn * delap) {
                if (sgn * (delap - alphap) >= bigm_thresh) {
                    alpham = (-2.0 * delap - 2.0 * sgn * sqrt(delap * delap - delap * alphap));
                } else {
                    alpham = -2.0 * alphap;
                }
            }
        }
    }
    sm = s0 + alpham;
    sp = s0 + alphap;
}

// ppm_type 2, second half: limiter + BC rows given the four unlimited edges around the cell
// (sedgel = edge(i-3/2) .. sedgerr = edge(i+3/2)).  MaestroPPM.cpp:356-719.
template <bool BC>
MXD void ppm2_finish(const double* S, int I, int lo, int hi, bool lobc, bool hibc, bool xquirk,
                     double sedgel, double sedge, double sedger, double sedgerr, double& sm,
                     double& sp) {
    const double sm3 = S[0], sm2 = S[1], sm1 = S[2], s0 = S[3], sp1 = S[4], sp2 = S[5],
                 sp3 = S[6];
    cs_limit(sedgel, sedge, sedger, sedgerr, sm2, sm1, s0, sp1, sp2, 1.e-10, sm, sp);
    if (BC) {
        if (lobc) {
            if (I == lo) {
                sm = sm1;
                sp = bc_edge(sm1, s0, sp1, sp2, sp1, s0);
            } else if (I == lo + 1) {
                sedgel = sm2;
                sedge = bc_edge(sm2, sm1, s0, sp1, s0, sm1);
            } else if (I == lo + 2) {
                sedgel = bc_edge(sm3, sm2, sm1, s0, sm1, sm2);
            }
            if (I == lo + 1 || I == lo + 2) {
                // QUIRK: x-direction lo side tests 1.e10 (MaestroPPM.cpp:566), y/z 1.e-10
                cs_limit(sedgel, sedge, sedger, sedgerr, sm2, sm1, s0, sp1, sp2,
                         xquirk ? 1.e10 : 1.e-10, sm, sp);
            }
        }
        if (hibc) {
            if (I == hi) {
                sp = sp1;
                sm = bc_edge(sp1, s0, sm1, sm2, sm1, s0);
            } else if (I == hi - 1) {
                // QUIRK: y and z reset sedgerr = s(+2) (MaestroPPM.cpp:1343,:2087), x does not
                if (!xquirk) sedgerr = sp2;
                sedger = bc_edge(sp2, sp1, s0, sm1, s0, sp1);
            } else if (I == hi - 2) {
                sedgerr = bc_edge(sp3, sp2, sp1, s0, sp1, sp2);
            }
            if (I == hi - 1 || I == hi - 2) {
                cs_limit(sedgel, sedge, sedger, sedgerr, sm2, sm1, s0, sp1, sp2, 1.e-10, sm, sp);
            }
        }
    }
}

// ppm_type 2 parabola: S[o+3], o in [-3,3].  MaestroPPM.cpp:255-719.
template <bool BC>
MXD void ppm2_parabola(const double* S, int I, int lo, int hi, bool lobc, bool hibc, bool xquirk,
                       double& sm, double& sp) {
    const double sm3 = S[0], sm2 = S[1], sm1 = S[2], s0 = S[3], sp1 = S[4], sp2 = S[5],
                 sp3 = S[6];
    ppm2_finish<BC>(S, I, lo, hi, lobc, hibc, xquirk, cs_edge(sm3, sm2, sm1, s0),
                    cs_edge(sm2, sm1, s0, sp1), cs_edge(sm1, s0, sp1, sp2),
                    cs_edge(s0, sp1, sp2, sp3), sm, sp);
}

// characteristic integration, MaestroPPM.cpp:210-231 (is_umac).  sig = |u| dt / h already
// formed per face by the caller ((|u| * dt) / h, IEEE division).
template <class T>
MXD T ppm_Ip(T sm, T sp, T s0, T s6, T up, T sig_p, double eps) {
    const T v = sp - 0.5 * sig_p * (sp - sm - (1.0 - 2.0 / 3.0 * sig_p) * s6);
    return sel(up > eps, v, s0);
}
template <class T>
MXD T ppm_Im(T sm, T sp, T s0, T s6, T um, T sig_m, double eps) {
    const T v = sm + 0.5 * sig_m * (sp - sm + (1.0 - 2.0 / 3.0 * sig_m) * s6);
    return sel(um < -eps, v, s0);
}

// ---------------------------------------------------------------------------------------
// Riemann / BC helpers of MakeEdgeScalPredictor / Transverse / Edges
// ---------------------------------------------------------------------------------------
MXD double upwind(double u, double eps, double sl, double sr) {
    // MaestroMakeEdgeScalars.cpp:745-748
    const double r = (u > 0.0) ? sl : sr;
    return (fabs(u) > eps) ? r : 0.5 * (sl + sr);
}

// the same Riemann choice with the velocity predicates formed once per face: pos = u > 0,
// nz = |u| > eps; `slow` is a warp-uniform "some lane of this warp has !nz" flag, so the
// averaging arm costs nothing where the flow is nowhere stagnant
MXD double upw(bool pos, bool nz, bool slow, double sl, double sr) {
    double r = pos ? sl : sr;
    if (slow) {
        const double avg = 0.5 * (sl + sr);
        r = nz ? r : avg;
    }
    return r;
}

// the stagnation arm of the Riemann choice on its own: `r` is the upwind state already chosen
MXD double upw_avg(bool nz, double r, double sl, double sr) { return nz ? r : 0.5 * (sl + sr); }

// :710-742 and the identical blocks in the transverse stage
MXD void bc_pair(int I, int lo, int hi, int bclo, int bchi, bool velclamp, double s_lo,
                 double s_hi, double& sl, double& sr) {
    if (I == lo) {
        if (bclo == BC_EXT_DIR) {
            sl = s_lo;
            sr = s_lo;
        } else if (bclo == BC_FOEXTRAP || bclo == BC_HOEXTRAP) {
            if (velclamp) sr = dmin_(sr, 0.0);
            sl = sr;
        } else if (bclo == BC_REFLECT_EVEN) {
            sl = sr;
        } else if (bclo == BC_REFLECT_ODD) {
            sl = 0.0;
            sr = 0.0;
        }
    } else if (I == hi + 1) {
        if (bchi == BC_EXT_DIR) {
            sl = s_hi;
            sr = s_hi;
        } else if (bchi == BC_FOEXTRAP || bchi == BC_HOEXTRAP) {
            if (velclamp) sl = dmax_(sl, 0.0);
            sr = sl;
        } else if (bchi == BC_REFLECT_EVEN) {
            sr = sl;
        } else if (bchi == BC_REFLECT_ODD) {
            sl = 0.0;
            sr = 0.0;
        }
    }
}

// :1455-1485
MXD double bc_final(int I, int lo, int hi, int bclo, int bchi, bool velclamp, double s_lo,
                    double s_hi, double el, double er, double val) {
    if (I == lo) {
        if (bclo == BC_EXT_DIR) {
            val = s_lo;
        } else if (bclo == BC_FOEXTRAP || bclo == BC_HOEXTRAP) {
            val = velclamp ? dmin_(er, 0.0) : er;
        } else if (bclo == BC_REFLECT_EVEN) {
            val = er;
        } else if (bclo == BC_REFLECT_ODD) {
            val = 0.0;
        }
    } else if (I == hi + 1) {
        if (bchi == BC_EXT_DIR) {
            val = s_hi;
        } else if (bchi == BC_FOEXTRAP || bchi == BC_HOEXTRAP) {
            val = velclamp ? dmax_(el, 0.0) : el;
        } else if (bchi == BC_REFLECT_EVEN) {
            val = el;
        } else if (bchi == BC_REFLECT_ODD) {
            val = 0.0;
        }
    }
    return val;
}

}  // namespace mxd
#endif
