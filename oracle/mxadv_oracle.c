/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle for the MAESTROeX scalar-advection hot path.
 *
 * A plain-C restatement (our own code, direction-generic, flat arrays) of the reference
 * algorithm.  It is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg; the product (libmxadv.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks it bit-for-bit against
 * oracle/_ref/libmaestro_ref{2,3}d.so, which is the reference's own
 * Source/Maestro{MakeEdgeScalars,PPM,Slopes,MakeFlux,CelltoEdge,Makew0,MakeGrav}.cpp compiled
 * unmodified
 * (oracle/Makefile), and against the golden vectors under tests/golden/ generated from that
 * library (tests/golden/make_golden.py).
 *
 * Compile with -ffp-contract=off (no FMA): the reference CPU build has none either
 * (g++ -O3 without -march), and every expression below keeps the reference's association.
 *
 * Reference files followed (relative to the MAESTROeX tree):
 *   Source/MaestroSlopes.cpp:6-601          -> slope_dir()
 *   Source/MaestroPPM.cpp:6-2263            -> ppm_dir()
 *   Source/MaestroMakeEdgeScalars.cpp:6-302 -> orc_make_edge_scal() (driver)
 *     :306-452 / :662-863                   -> predictor_dir()  (2-D / 3-D)
 *     :648-660                              -> MakeDivU in orc_make_edge_scal()
 *     :865-1359                             -> transverse()
 *     :454-644 / :1361-1667                 -> edges_dir()      (2-D / 3-D)
 *   Source/MaestroMakeFlux.cpp:18-465       -> orc_make_rhoX_flux()
 *   Source/MaestroMakeFlux.cpp:467-1067     -> orc_make_rhoh_flux()
 *   Source/MaestroCelltoEdge.cpp:5-55       -> orc_cell_to_edge()
 *   Source/MaestroAdvection.cpp:86-251      -> orc_update_scal() (zone loop; EOS zones flagged)
 *   Source/MaestroMakew0.cpp:5-58,60-272    -> orc_make_w0_planar() (one radial level)
 *   Source/MaestroMakew0.cpp:537-740,942-969 + Source/MaestroMakeGrav.cpp:303-322
 *                                           -> orc_make_w0_sphr()
 *   Source/MaestroMakeUtrans.cpp:6-648      -> orc_make_utrans()
 *   Source/MaestroVelPred.cpp:5-1917        -> orc_vel_pred()  (tests/test_velpred.py pins both
 *                                              to MaestroMakeUtrans.cpp / MaestroVelPred.cpp in _ref)
 * Direction-dependent quirks of the reference that are reproduced on purpose are marked
 * "QUIRK" with the file:line they come from.
 */
#define _GNU_SOURCE /* M_PI under -std=c11 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mxadv.h"

#define BC_REFLECT_ODD (-1)
#define BC_REFLECT_EVEN 1
#define BC_FOEXTRAP 2
#define BC_EXT_DIR 3
#define BC_HOEXTRAP 4

/* std::min / std::max semantics (amrex::min/max), not fmin/fmax */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }

typedef struct {
    int lo[3], hi[3];
} IBox;

static IBox grow_all(IBox b, int n, int dim) {
    for (int d = 0; d < dim; ++d) { b.lo[d] -= n; b.hi[d] += n; }
    return b;
}

/* strided view of one component of a fab */
typedef struct {
    const double* p; /* points at (lo) of the chosen component */
    long st[3];
    int lo[3];
} CView;

static CView cview(const MxFab* f, int comp) {
    CView v;
    long nx = f->hi[0] - f->lo[0] + 1, ny = f->hi[1] - f->lo[1] + 1,
         nz = f->hi[2] - f->lo[2] + 1;
    v.p = f->p + (long)comp * nx * ny * nz;
    v.st[0] = 1; v.st[1] = nx; v.st[2] = nx * ny;
    for (int d = 0; d < 3; ++d) v.lo[d] = f->lo[d];
    return v;
}
static inline long voff(const CView* v, int i, int j, int k) {
    return (i - v->lo[0]) * v->st[0] + (j - v->lo[1]) * v->st[1] + (k - v->lo[2]) * v->st[2];
}
#define RD(v, i, j, k) ((v).p[voff(&(v), (i), (j), (k))])

/* scratch array on a box */
typedef struct {
    double* p;
    long st[3];
    int lo[3];
} Tmp;
static Tmp tmp_alloc(IBox b) {
    Tmp t;
    long n0 = b.hi[0] - b.lo[0] + 1, n1 = b.hi[1] - b.lo[1] + 1, n2 = b.hi[2] - b.lo[2] + 1;
    t.p = (double*)calloc((size_t)(n0 * n1 * n2), sizeof(double));
    t.st[0] = 1; t.st[1] = n0; t.st[2] = n0 * n1;
    for (int d = 0; d < 3; ++d) t.lo[d] = b.lo[d];
    return t;
}
static inline long toff(const Tmp* t, int i, int j, int k) {
    return (i - t->lo[0]) + (j - t->lo[1]) * t->st[1] + (k - t->lo[2]) * t->st[2];
}
#define T(t, i, j, k) ((t).p[toff(&(t), (i), (j), (k))])

typedef struct {
    int dim;
    IBox valid, obx;
    int domlo[3], domhi[3];
    double dx[3];
    double dt, rel_eps;
    int ppm_type, slope_order, ppm_trace_forces, is_vel, is_cons;
} Ctx;

/* ------------------------------------------------------------------------------------ */
/* Slopex/y/z: Source/MaestroSlopes.cpp:6-203 (x), :205-401 (y), :403-601 (z)             */
/* ------------------------------------------------------------------------------------ */
static void slope_dir(const Ctx* c, int d, const CView* s, Tmp* out, int bclo, int bchi) {
    const int e0 = d == 0, e1 = d == 1, e2 = d == 2;
    const int lo = c->domlo[d], hi = c->domhi[d];
    const int lobc = (bclo == BC_EXT_DIR || bclo == BC_HOEXTRAP);
    const int hibc = (bchi == BC_EXT_DIR || bchi == BC_HOEXTRAP);
    for (int k = c->obx.lo[2]; k <= c->obx.hi[2]; ++k)
    for (int j = c->obx.lo[1]; j <= c->obx.hi[1]; ++j)
    for (int i = c->obx.lo[0]; i <= c->obx.hi[0]; ++i) {
#define S(o) RD(*s, i + (o) * e0, j + (o) * e1, k + (o) * e2)
        const int I = d == 0 ? i : (d == 1 ? j : k);
        double r;
        if (c->slope_order == 0) {
            r = 0.0; /* MaestroSlopes.cpp:35-38 */
        } else if (c->slope_order == 2) {
            /* MaestroSlopes.cpp:43-90 */
            double del = 0.5 * (S(1) - S(-1));
            double dpls = 2.0 * (S(1) - S(0));
            double dmn = 2.0 * (S(0) - S(-1));
            double slim = dmin(fabs(dpls), fabs(dmn));
            slim = dpls * dmn > 0.0 ? slim : 0.0;
            double sflag = copysign(1.0, del);
            r = sflag * dmin(slim, fabs(del));
            if (lobc) {
                if (I == lo - 1) {
                    r = 0.0;
                } else if (I == lo) {
                    del = (S(1) + 3.0 * S(0) - 4.0 * S(-1)) / 3.0;
                    dpls = 2.0 * (S(1) - S(0));
                    dmn = 2.0 * (S(0) - S(-1));
                    slim = dmin(fabs(dpls), fabs(dmn));
                    slim = dpls * dmn > 0.0 ? slim : 0.0;
                    sflag = copysign(1.0, del);
                    r = sflag * dmin(slim, fabs(del));
                }
            }
            if (hibc) {
                if (I == hi + 1) {
                    r = 0.0;
                } else if (I == hi) {
                    del = -(S(-1) + 3.0 * S(0) - 4.0 * S(1)) / 3.0;
                    dpls = 2.0 * (S(1) - S(0));
                    dmn = 2.0 * (S(0) - S(-1));
                    slim = dmin(fabs(dpls), fabs(dmn));
                    slim = dpls * dmn > 0.0 ? slim : 0.0;
                    sflag = copysign(1.0, del);
                    r = sflag * dmin(slim, fabs(del));
                }
            }
        } else {
            /* 4th order, MaestroSlopes.cpp:95-201 */
            double dcen = 0.5 * (S(0) - S(-2));
            double dmn = 2.0 * (S(-1) - S(-2));
            double dpls = 2.0 * (S(0) - S(-1));
            double dlim = dmin(fabs(dmn), fabs(dpls));
            dlim = dpls * dmn > 0.0 ? dlim : 0.0;
            double dflag = copysign(1.0, dcen);
            double dxl = dflag * dmin(dlim, fabs(dcen));

            dcen = 0.5 * (S(2) - S(0));
            dmn = 2.0 * (S(1) - S(0));
            dpls = 2.0 * (S(2) - S(1));
            dlim = dmin(fabs(dmn), fabs(dpls));
            dlim = dpls * dmn > 0.0 ? dlim : 0.0;
            dflag = copysign(1.0, dcen);
            double dxr = dflag * dmin(dlim, fabs(dcen));

            dcen = 0.5 * (S(1) - S(-1));
            dmn = 2.0 * (S(0) - S(-1));
            dpls = 2.0 * (S(1) - S(0));
            dlim = dmin(fabs(dmn), fabs(dpls));
            dlim = dpls * dmn > 0.0 ? dlim : 0.0;
            dflag = copysign(1.0, dcen);

            double ds = 4.0 / 3.0 * dcen - (dxr + dxl) / 6.0;
            r = dflag * dmin(fabs(ds), dlim);

            /* QUIRK: in the one-sided rows Slopex selects `dlim` (the centred limiter
             * computed above) where Slopey/Slopez select the freshly computed `slim`
             * (MaestroSlopes.cpp:142,156,178,192 vs :341,355,377,391 and :539,...). */
#define SEL_LIM(slim) (dpls * dmn > 0.0 ? (d == 0 ? dlim : (slim)) : 0.0)
            if (lobc) {
                if (I == lo - 1) {
                    r = 0.0;
                } else if (I == lo) {
                    double del = -16.0 / 15.0 * S(-1) + 0.5 * S(0) + 2.0 / 3.0 * S(1) -
                                 0.1 * S(2);
                    dmn = 2.0 * (S(0) - S(-1));
                    dpls = 2.0 * (S(1) - S(0));
                    double slim = dmin(fabs(dpls), fabs(dmn));
                    slim = SEL_LIM(slim);
                    double sflag = copysign(1.0, del);
                    r = sflag * dmin(slim, fabs(del));
                } else if (I == lo + 1) {
                    double del = -16.0 / 15.0 * S(-2) + 0.5 * S(-1) + 2.0 / 3.0 * S(0) -
                                 0.1 * S(1);
                    dmn = 2.0 * (S(-1) - S(-2));
                    dpls = 2.0 * (S(0) - S(-1));
                    double slim = dmin(fabs(dpls), fabs(dmn));
                    slim = SEL_LIM(slim);
                    double sflag = copysign(1.0, del);
                    dxl = sflag * dmin(slim, fabs(del));
                    ds = 4.0 / 3.0 * dcen - (dxr + dxl) / 6.0;
                    r = dflag * dmin(fabs(ds), dlim);
                }
            }
            if (hibc) {
                if (I == hi + 1) {
                    r = 0.0;
                } else if (I == hi) {
                    double del = -(-16.0 / 15.0 * S(1) + 0.5 * S(0) + 2.0 / 3.0 * S(-1) -
                                   0.1 * S(-2));
                    dmn = 2.0 * (S(0) - S(-1));
                    dpls = 2.0 * (S(1) - S(0));
                    double slim = dmin(fabs(dpls), fabs(dmn));
                    slim = SEL_LIM(slim);
                    double sflag = copysign(1.0, del);
                    r = sflag * dmin(slim, fabs(del));
                } else if (I == hi - 1) {
                    double del = -(-16.0 / 15.0 * S(2) + 0.5 * S(1) + 2.0 / 3.0 * S(0) -
                                   0.1 * S(-1));
                    dmn = 2.0 * (S(1) - S(0));
                    dpls = 2.0 * (S(2) - S(1));
                    double slim = dmin(fabs(dpls), fabs(dmn));
                    slim = SEL_LIM(slim);
                    double sflag = copysign(1.0, del);
                    dxr = sflag * dmin(slim, fabs(del));
                    ds = 4.0 / 3.0 * dcen - (dxl + dxr) / 6.0;
                    r = dflag * dmin(fabs(ds), dlim);
                }
            }
#undef SEL_LIM
        }
        T(*out, i, j, k) = r;
#undef S
    }
}

/* ------------------------------------------------------------------------------------ */
/* PPM: Source/MaestroPPM.cpp:27-771 (x), :773-1513 (y), :1515-2262 (z), is_umac = true   */
/* ------------------------------------------------------------------------------------ */
static inline double vanleer(double sm1, double s0, double sp1) {
    /* MaestroPPM.cpp:42-49 */
    double dsc = 0.5 * (sp1 - sm1);
    double dsl = 2.0 * (s0 - sm1);
    double dsr = 2.0 * (sp1 - s0);
    if (dsl * dsr > 0.0) return copysign(1.0, dsc) * dmin(fabs(dsc), dmin(fabs(dsl), fabs(dsr)));
    return 0.0;
}

static inline void cw_limit(double s0, double* sm, double* sp) {
    /* MaestroPPM.cpp:106-115 */
    if ((*sp - s0) * (s0 - *sm) <= 0.0) {
        *sp = s0;
        *sm = s0;
    } else if (fabs(*sp - s0) >= 2.0 * fabs(*sm - s0)) {
        *sp = 3.0 * s0 - 2.0 * *sm;
    } else if (fabs(*sm - s0) >= 2.0 * fabs(*sp - s0)) {
        *sm = 3.0 * s0 - 2.0 * *sp;
    }
}

static inline double bc_edge(double a, double b, double cc, double dd, double mn0, double mn1) {
    /* -0.2 a + 0.75 b + 0.5 c - 0.05 d clamped between mn0, mn1: MaestroPPM.cpp:124-131 */
    double v = -0.2 * a + 0.75 * b + 0.5 * cc - 0.05 * dd;
    v = dmax(v, dmin(mn0, mn1));
    v = dmin(v, dmax(mn0, mn1));
    return v;
}

/* 4th-order edge with 2nd-derivative limiter, MaestroPPM.cpp:258-279; sm2..sp1 are the
 * four cells around the edge between s_m1 and s_0 */
static inline double cs_edge(double sm2, double sm1, double s0, double sp1) {
    const double C = 1.25;
    double e = (7.0 / 12.0) * (sm1 + s0) - (1.0 / 12.0) * (sm2 + sp1);
    if ((e - sm1) * (s0 - e) < 0.0) {
        double D2 = 3.0 * (sm1 - 2.0 * e + s0);
        double D2L = sm2 - 2.0 * sm1 + s0;
        double D2R = sm1 - 2.0 * s0 + sp1;
        double sgn = copysign(1.0, D2);
        double D2LIM = sgn * dmax(dmin(C * sgn * D2L, dmin(C * sgn * D2R, sgn * D2)), 0.0);
        e = 0.5 * (sm1 + s0) - D2LIM / 6.0;
    }
    return e;
}

/* Colella-Sekora limiter, MaestroPPM.cpp:356-441 (and the BC re-applications :488-580,
 * :627-718).  `big_thresh` is 1.e-10 except for the QUIRK noted at the call site. */
static inline void cs_limit(double sedgel, double sedge, double sedger, double sedgerr,
                            double sm2, double sm1, double s0, double sp1, double sp2,
                            double bigm_thresh, double* sm_out, double* sp_out) {
    const double C = 1.25;
    double alphap = sedger - s0;
    double alpham = sedge - s0;
    int bigp = fabs(alphap) > 2.0 * fabs(alpham);
    int bigm = fabs(alpham) > 2.0 * fabs(alphap);
    int extremum = 0;
    if (alpham * alphap >= 0.0) {
        extremum = 1;
    } else if (bigp || bigm) {
        double dafacem = sedge - sedgel;
        double dafacep = sedgerr - sedger;
        double dabarm = s0 - sm1;
        double dabarp = sp1 - s0;
        double dafacemin = dmin(fabs(dafacem), fabs(dafacep));
        double dabarmin = dmin(fabs(dabarm), fabs(dabarp));
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
        double D2 = 6.0 * (alpham + alphap);
        double D2L = sm2 - 2.0 * sm1 + s0;
        double D2R = s0 - 2.0 * sp1 + sp2;
        double D2C = sm1 - 2.0 * s0 + sp1;
        double sgn = copysign(1.0, D2);
        double D2LIM =
            dmax(dmin(sgn * D2, dmin(C * sgn * D2L, dmin(C * sgn * D2R, C * sgn * D2C))), 0.0);
        double D2ABS = dmax(fabs(D2), 1.e-10);
        alpham = alpham * D2LIM / D2ABS;
        alphap = alphap * D2LIM / D2ABS;
    } else {
        if (bigp) {
            double sgn = copysign(1.0, alpham);
            double amax = -alphap * alphap / (4.0 * (alpham + alphap));
            double delam = sm1 - s0;
            if (sgn * amax >= sgn * delam) {
                if (sgn * (delam - alpham) >= 1.e-10) {
                    alphap = (-2.0 * delam - 2.0 * sgn * sqrt(delam * delam - delam * alpham));
                } else {
                    alphap = -2.0 * alpham;
                }
            }
        }
        if (bigm) {
            double sgn = copysign(1.0, alphap);
            double amax = -alpham * alpham / (4.0 * (alpham + alphap));
            double delap = sp1 - s0;
            if (sgn * amax >= sgn * delap) {
                if (sgn * (delap - alphap) >= bigm_thresh) {
                    alpham = (-2.0 * delap - 2.0 * sgn * sqrt(delap * delap - delap * alphap));
                } else {
                    alpham = -2.0 * alphap;
                }
            }
        }
    }
    *sm_out = s0 + alpham;
    *sp_out = s0 + alphap;
}

static void ppm_dir(const Ctx* c, int d, const CView* s, const CView* vel, Tmp* Ip, Tmp* Im,
                    int bclo, int bchi, int is_umac) {
    const int e0 = d == 0, e1 = d == 1, e2 = d == 2;
    const int lo = c->domlo[d], hi = c->domhi[d];
    const int lobc = (bclo == BC_EXT_DIR || bclo == BC_HOEXTRAP);
    const int hibc = (bchi == BC_EXT_DIR || bchi == BC_HOEXTRAP);
    const double dt = c->dt, h = c->dx[d], eps = c->rel_eps;
    for (int k = c->obx.lo[2]; k <= c->obx.hi[2]; ++k)
    for (int j = c->obx.lo[1]; j <= c->obx.hi[1]; ++j)
    for (int i = c->obx.lo[0]; i <= c->obx.hi[0]; ++i) {
#define S(o) RD(*s, i + (o) * e0, j + (o) * e1, k + (o) * e2)
        const int I = d == 0 ? i : (d == 1 ? j : k);
        double sm, sp;
        const double s0 = S(0);
        if (c->ppm_type == 1) {
            /* MaestroPPM.cpp:34-205 */
            double dl = vanleer(S(-2), S(-1), S(0));
            double dc = vanleer(S(-1), S(0), S(1));
            double dr = vanleer(S(0), S(1), S(2));
            sm = 0.5 * (S(0) + S(-1)) - (dc - dl) / 6.0;
            sm = dmax(sm, dmin(S(0), S(-1)));
            sm = dmin(sm, dmax(S(0), S(-1)));
            sp = 0.5 * (S(1) + S(0)) - (dr - dc) / 6.0;
            sp = dmax(sp, dmin(S(1), S(0)));
            sp = dmin(sp, dmax(S(1), S(0)));
            const double sedgel = sp, sedger = sm;
            cw_limit(s0, &sm, &sp);
            if (I == lo) {
                if (lobc) {
                    sm = S(-1);
                    sp = bc_edge(S(-1), S(0), S(1), S(2), S(1), S(0));
                }
            } else if (I == lo + 1) {
                if (lobc) {
                    sm = bc_edge(S(-2), S(-1), S(0), S(1), S(0), S(-1));
                    sp = sedgel;
                    cw_limit(s0, &sm, &sp);
                }
            } else if (I == hi) {
                if (hibc) {
                    sp = S(1);
                    sm = bc_edge(S(1), S(0), S(-1), S(-2), S(-1), S(0));
                }
            } else if (I == hi - 1) {
                if (hibc) {
                    sp = bc_edge(S(2), S(1), S(0), S(-1), S(0), S(1));
                    sm = sedger;
                    cw_limit(s0, &sm, &sp);
                }
            }
        } else {
            /* ppm_type == 2, MaestroPPM.cpp:255-719 */
            double sedgel = cs_edge(S(-3), S(-2), S(-1), S(0));
            double sedge = cs_edge(S(-2), S(-1), S(0), S(1));
            double sedger = cs_edge(S(-1), S(0), S(1), S(2));
            double sedgerr = cs_edge(S(0), S(1), S(2), S(3));
            cs_limit(sedgel, sedge, sedger, sedgerr, S(-2), S(-1), S(0), S(1), S(2), 1.e-10,
                     &sm, &sp);
            if (lobc) {
                if (I == lo) {
                    sm = S(-1);
                    sp = bc_edge(S(-1), S(0), S(1), S(2), S(1), S(0));
                } else if (I == lo + 1) {
                    sedgel = S(-2);
                    sedge = bc_edge(S(-2), S(-1), S(0), S(1), S(0), S(-1));
                } else if (I == lo + 2) {
                    sedgel = bc_edge(S(-3), S(-2), S(-1), S(0), S(-1), S(-2));
                }
                if (I == lo + 1 || I == lo + 2) {
                    /* QUIRK: the x-direction lo-side re-application tests `>= 1.e10` in the
                     * bigm branch (MaestroPPM.cpp:566); y and z test 1.e-10 (:1300, :2044) */
                    cs_limit(sedgel, sedge, sedger, sedgerr, S(-2), S(-1), S(0), S(1), S(2),
                             d == 0 ? 1.e10 : 1.e-10, &sm, &sp);
                }
            }
            if (hibc) {
                if (I == hi) {
                    sp = S(1);
                    sm = bc_edge(S(1), S(0), S(-1), S(-2), S(-1), S(0));
                } else if (I == hi - 1) {
                    /* QUIRK: y and z set sedgerr = s(+2) here (MaestroPPM.cpp:1343, :2087);
                     * the x block (:598-608) does not */
                    if (d != 0) sedgerr = S(2);
                    sedger = bc_edge(S(2), S(1), S(0), S(-1), S(0), S(1));
                } else if (I == hi - 2) {
                    sedgerr = bc_edge(S(3), S(2), S(1), S(0), S(1), S(2));
                }
                if (I == hi - 1 || I == hi - 2) {
                    cs_limit(sedgel, sedge, sedger, sedgerr, S(-2), S(-1), S(0), S(1), S(2),
                             1.e-10, &sm, &sp);
                }
            }
        }
        /* MaestroPPM.cpp:210-231 (is_umac: MAC velocity, edge-based indexing) / :233-251 (cell-
         * centred velocity: both integrals use u(i,j,k); MakeUtrans and VelPred) */
        const double s6 = 6.0 * s0 - 3.0 * (sm + sp);
        const double up = is_umac ? RD(*vel, i + e0, j + e1, k + e2) : RD(*vel, i, j, k);
        const double um = RD(*vel, i, j, k);
        double sigma = fabs(up) * dt / h;
        if (up > eps) {
            T(*Ip, i, j, k) = sp - 0.5 * sigma * (sp - sm - (1.0 - 2.0 / 3.0 * sigma) * s6);
        } else {
            T(*Ip, i, j, k) = s0;
        }
        sigma = fabs(um) * dt / h;
        if (um < -eps) {
            T(*Im, i, j, k) = sm + 0.5 * sigma * (sp - sm + (1.0 - 2.0 / 3.0 * sigma) * s6);
        } else {
            T(*Im, i, j, k) = s0;
        }
#undef S
    }
}

/* physical-BC override of a left/right pair at a domain face; the pattern shared by
 * MaestroMakeEdgeScalars.cpp:710-742 (predictor) and :933-965 etc. (transverse) */
static inline void bc_pair(int I, int lo, int hi, int bclo, int bchi, int velclamp,
                           double s_lo_ghost, double s_hi_ghost, double* sl, double* sr) {
    if (I == lo) {
        if (bclo == BC_EXT_DIR) {
            *sl = s_lo_ghost;
            *sr = s_lo_ghost;
        } else if (bclo == BC_FOEXTRAP || bclo == BC_HOEXTRAP) {
            if (velclamp) *sr = dmin(*sr, 0.0);
            *sl = *sr;
        } else if (bclo == BC_REFLECT_EVEN) {
            *sl = *sr;
        } else if (bclo == BC_REFLECT_ODD) {
            *sl = 0.0;
            *sr = 0.0;
        }
    } else if (I == hi + 1) {
        if (bchi == BC_EXT_DIR) {
            *sl = s_hi_ghost;
            *sr = s_hi_ghost;
        } else if (bchi == BC_FOEXTRAP || bchi == BC_HOEXTRAP) {
            if (velclamp) *sl = dmax(*sl, 0.0);
            *sr = *sl;
        } else if (bchi == BC_REFLECT_EVEN) {
            *sr = *sl;
        } else if (bchi == BC_REFLECT_ODD) {
            *sl = 0.0;
            *sr = 0.0;
        }
    }
}

static inline double upwind(double u, double eps, double sl, double sr) {
    double r = (u > 0.0) ? sl : sr;
    return (fabs(u) > eps) ? r : 0.5 * (sl + sr);
}

/* ------------------------------------------------------------------------------------ */
/* MakeEdgeScalPredictor: 3-D MaestroMakeEdgeScalars.cpp:662-863, 2-D :306-452           */
/* faces of direction d: index d in [vlo, vhi+1], other directions in [vlo-1, vhi+1]      */
/* ------------------------------------------------------------------------------------ */
static void predictor_dir(const Ctx* c, int d, const CView* s, const CView* vel,
                          const Tmp* Ip_or_slope, const Tmp* Im, Tmp* sl, Tmp* sr, Tmp* simh,
                          int bclo, int bchi, int comp) {
    const int e0 = d == 0, e1 = d == 1, e2 = d == 2;
    IBox b = c->obx;
    b.lo[d] += 1;
    const double h = c->dx[d], dt = c->dt, dt2 = 0.5 * c->dt;
    /* QUIRK: the 3-D predictor upwinds with |u| > 0.0 (MaestroMakeEdgeScalars.cpp:746,803,
     * 859); the 2-D one with rel_eps (:389,448) */
    const double eps = (c->dim == 3) ? 0.0 : c->rel_eps;
    for (int k = b.lo[2]; k <= b.hi[2]; ++k)
    for (int j = b.lo[1]; j <= b.hi[1]; ++j)
    for (int i = b.lo[0]; i <= b.hi[0]; ++i) {
        const int I = d == 0 ? i : (d == 1 ? j : k);
        const double u = RD(*vel, i, j, k);
        const double sL = RD(*s, i - e0, j - e1, k - e2), sR = RD(*s, i, j, k);
        double l, r;
        if (c->ppm_type == 0) {
            const double slpL = T(*Ip_or_slope, i - e0, j - e1, k - e2);
            const double slpR = T(*Ip_or_slope, i, j, k);
            if (c->dim == 3) {
                /* :698-703 */
                l = sL + 0.5 * (1.0 - dt * u / h) * slpL;
                r = sR - 0.5 * (1.0 + dt * u / h) * slpR;
            } else {
                /* QUIRK: 2-D association differs (:341-345) */
                l = sL + (0.5 - dt2 * u / h) * slpL;
                r = sR - (0.5 + dt2 * u / h) * slpR;
            }
        } else {
            l = T(*Ip_or_slope, i - e0, j - e1, k - e2);
            r = T(*Im, i, j, k);
        }
        bc_pair(I, c->domlo[d], c->domhi[d], bclo, bchi, c->is_vel && comp == d, sL, sR, &l,
                &r);
        T(*sl, i, j, k) = l;
        T(*sr, i, j, k) = r;
        T(*simh, i, j, k) = upwind(u, eps, l, r);
    }
}

/* ------------------------------------------------------------------------------------ */
/* MakeEdgeScalTransverse (3-D): MaestroMakeEdgeScalars.cpp:865-1359.  simh_{d|t}: state  */
/* on d-faces corrected by the t-direction transverse term; the box is the valid box with */
/* the d-faces [vlo, vhi+1] and grown by one in the third direction o (:897-898 etc.).    */
/* ------------------------------------------------------------------------------------ */
static void transverse(const Ctx* c, int d, int t, const CView* s, const CView* vel_d,
                       const CView* vel_t, const Tmp* sl, const Tmp* sr, const Tmp* simh_t,
                       const Tmp* divu, Tmp* out, int bclo, int bchi, int comp) {
    const int o = 3 - d - t;
    const int e[3] = {d == 0, d == 1, d == 2};
    const int f[3] = {t == 0, t == 1, t == 2};
    IBox b = c->valid;
    b.lo[o] -= 1;
    b.hi[o] += 1;
    b.hi[d] += 1;
    const double dt3 = c->dt / 3.0, dt6 = c->dt / 6.0, ht = c->dx[t];
    for (int k = b.lo[2]; k <= b.hi[2]; ++k)
    for (int j = b.lo[1]; j <= b.hi[1]; ++j)
    for (int i = b.lo[0]; i <= b.hi[0]; ++i) {
        const int I = d == 0 ? i : (d == 1 ? j : k);
        /* left cell L = (i,j,k) - e_d, right cell R = (i,j,k) */
        const int li = i - e[0], lj = j - e[1], lk = k - e[2];
        double l, r;
        const double sL = RD(*s, li, lj, lk), sR = RD(*s, i, j, k);
        if (c->is_cons) {
            /* :909-921 */
            l = T(*sl, i, j, k) -
                (dt3 / ht) * (T(*simh_t, li + f[0], lj + f[1], lk + f[2]) *
                                  RD(*vel_t, li + f[0], lj + f[1], lk + f[2]) -
                              T(*simh_t, li, lj, lk) * RD(*vel_t, li, lj, lk)) -
                dt3 * sL * T(*divu, li, lj, lk) +
                (dt3 / ht) * sL *
                    (RD(*vel_t, li + f[0], lj + f[1], lk + f[2]) - RD(*vel_t, li, lj, lk));
            r = T(*sr, i, j, k) -
                (dt3 / ht) * (T(*simh_t, i + f[0], j + f[1], k + f[2]) *
                                  RD(*vel_t, i + f[0], j + f[1], k + f[2]) -
                              T(*simh_t, i, j, k) * RD(*vel_t, i, j, k)) -
                dt3 * sR * T(*divu, i, j, k) +
                (dt3 / ht) * sR *
                    (RD(*vel_t, i + f[0], j + f[1], k + f[2]) - RD(*vel_t, i, j, k));
        } else {
            /* :924-929 */
            l = T(*sl, i, j, k) -
                (dt6 / ht) *
                    (RD(*vel_t, li + f[0], lj + f[1], lk + f[2]) + RD(*vel_t, li, lj, lk)) *
                    (T(*simh_t, li + f[0], lj + f[1], lk + f[2]) - T(*simh_t, li, lj, lk));
            r = T(*sr, i, j, k) -
                (dt6 / ht) *
                    (RD(*vel_t, i + f[0], j + f[1], k + f[2]) + RD(*vel_t, i, j, k)) *
                    (T(*simh_t, i + f[0], j + f[1], k + f[2]) - T(*simh_t, i, j, k));
        }
        bc_pair(I, c->domlo[d], c->domhi[d], bclo, bchi, c->is_vel && comp == d, sL, sR, &l,
                &r);
        T(*out, i, j, k) = upwind(RD(*vel_d, i, j, k), c->rel_eps, l, r);
    }
}

/* final physical-BC override of the edge state, MaestroMakeEdgeScalars.cpp:1455-1485 */
static inline double bc_final(int I, int lo, int hi, int bclo, int bchi, int velclamp,
                              double s_lo_ghost, double s_hi_ghost, double el, double er,
                              double val) {
    if (I == lo) {
        if (bclo == BC_EXT_DIR) {
            val = s_lo_ghost;
        } else if (bclo == BC_FOEXTRAP || bclo == BC_HOEXTRAP) {
            val = velclamp ? dmin(er, 0.0) : er;
        } else if (bclo == BC_REFLECT_EVEN) {
            val = er;
        } else if (bclo == BC_REFLECT_ODD) {
            val = 0.0;
        }
    } else if (I == hi + 1) {
        if (bchi == BC_EXT_DIR) {
            val = s_hi_ghost;
        } else if (bchi == BC_FOEXTRAP || bchi == BC_HOEXTRAP) {
            val = velclamp ? dmax(el, 0.0) : el;
        } else if (bchi == BC_REFLECT_EVEN) {
            val = el;
        } else if (bchi == BC_REFLECT_ODD) {
            val = 0.0;
        }
    }
    return val;
}

/* ------------------------------------------------------------------------------------ */
/* MakeEdgeScalEdges: 3-D MaestroMakeEdgeScalars.cpp:1361-1667, 2-D :454-644.             */
/* tr1/tr2: the transverse states for the two other directions t1 < t2 (3-D: simh_{t1|t2} */
/* and simh_{t2|t1}; 2-D: simh_{t1} only).                                                */
/* ------------------------------------------------------------------------------------ */
static void edges_dir(const Ctx* c, int d, const CView* s, const CView* vel[3],
                      const CView* force, const Tmp* Ipf, const Tmp* Imf, const Tmp* sl,
                      const Tmp* sr, const Tmp* tr1, const Tmp* tr2, int t1, int t2,
                      const MxFab* sedge, int comp, int bclo, int bchi) {
    const int e[3] = {d == 0, d == 1, d == 2};
    const int f1[3] = {t1 == 0, t1 == 1, t1 == 2};
    const int f2[3] = {t2 == 0, t2 == 1, t2 == 2};
    IBox b = c->valid;
    b.hi[d] += 1;
    const double dt2 = 0.5 * c->dt, dt4 = 0.25 * c->dt;
    const double hd = c->dx[d], h1 = c->dx[t1], h2 = (c->dim == 3) ? c->dx[t2] : 0.0;
    const long nx = sedge->hi[0] - sedge->lo[0] + 1, ny = sedge->hi[1] - sedge->lo[1] + 1,
               nz = sedge->hi[2] - sedge->lo[2] + 1;
    double* out = sedge->p + (long)comp * nx * ny * nz;
    const CView* vd = vel[d];
    const CView* v1 = vel[t1];
    const CView* v2 = (c->dim == 3) ? vel[t2] : NULL;
    for (int k = b.lo[2]; k <= b.hi[2]; ++k)
    for (int j = b.lo[1]; j <= b.hi[1]; ++j)
    for (int i = b.lo[0]; i <= b.hi[0]; ++i) {
        const int I = d == 0 ? i : (d == 1 ? j : k);
        const int li = i - e[0], lj = j - e[1], lk = k - e[2];
        const double sL = RD(*s, li, lj, lk), sR = RD(*s, i, j, k);
        const double fl = (c->ppm_trace_forces == 0) ? RD(*force, li, lj, lk) : T(*Ipf, li, lj, lk);
        const double fr = (c->ppm_trace_forces == 0) ? RD(*force, i, j, k) : T(*Imf, i, j, k);
        double el, er;
#define P1(q, a, b_, c_) q(a + f1[0], b_ + f1[1], c_ + f1[2])
#define V1(a, b_, c_) RD(*v1, a, b_, c_)
#define V2(a, b_, c_) RD(*v2, a, b_, c_)
#define VD(a, b_, c_) RD(*vd, a, b_, c_)
#define TR1(a, b_, c_) T(*tr1, a, b_, c_)
#define TR2(a, b_, c_) T(*tr2, a, b_, c_)
        if (c->dim == 3) {
            if (c->is_cons) {
                /* :1412-1429 */
                el = T(*sl, i, j, k) -
                     (dt2 / h1) * (TR1(li + f1[0], lj + f1[1], lk + f1[2]) *
                                       V1(li + f1[0], lj + f1[1], lk + f1[2]) -
                                   TR1(li, lj, lk) * V1(li, lj, lk)) -
                     (dt2 / h2) * (TR2(li + f2[0], lj + f2[1], lk + f2[2]) *
                                       V2(li + f2[0], lj + f2[1], lk + f2[2]) -
                                   TR2(li, lj, lk) * V2(li, lj, lk)) -
                     (dt2 / hd) * sL * (VD(i, j, k) - VD(li, lj, lk)) + dt2 * fl;
                er = T(*sr, i, j, k) -
                     (dt2 / h1) * (TR1(i + f1[0], j + f1[1], k + f1[2]) *
                                       V1(i + f1[0], j + f1[1], k + f1[2]) -
                                   TR1(i, j, k) * V1(i, j, k)) -
                     (dt2 / h2) * (TR2(i + f2[0], j + f2[1], k + f2[2]) *
                                       V2(i + f2[0], j + f2[1], k + f2[2]) -
                                   TR2(i, j, k) * V2(i, j, k)) -
                     (dt2 / hd) * sR * (VD(i + e[0], j + e[1], k + e[2]) - VD(i, j, k)) +
                     dt2 * fr;
            } else {
                /* :1431-1443 */
                el = T(*sl, i, j, k) -
                     (dt4 / h1) *
                         (V1(li + f1[0], lj + f1[1], lk + f1[2]) + V1(li, lj, lk)) *
                         (TR1(li + f1[0], lj + f1[1], lk + f1[2]) - TR1(li, lj, lk)) -
                     (dt4 / h2) *
                         (V2(li + f2[0], lj + f2[1], lk + f2[2]) + V2(li, lj, lk)) *
                         (TR2(li + f2[0], lj + f2[1], lk + f2[2]) - TR2(li, lj, lk)) +
                     dt2 * fl;
                er = T(*sr, i, j, k) -
                     (dt4 / h1) * (V1(i + f1[0], j + f1[1], k + f1[2]) + V1(i, j, k)) *
                         (TR1(i + f1[0], j + f1[1], k + f1[2]) - TR1(i, j, k)) -
                     (dt4 / h2) * (V2(i + f2[0], j + f2[1], k + f2[2]) + V2(i, j, k)) *
                         (TR2(i + f2[0], j + f2[1], k + f2[2]) - TR2(i, j, k)) +
                     dt2 * fr;
            }
        } else {
            if (c->is_cons) {
                /* :499-511 */
                el = T(*sl, i, j, k) -
                     (dt2 / h1) * (TR1(li + f1[0], lj + f1[1], lk + f1[2]) *
                                       V1(li + f1[0], lj + f1[1], lk + f1[2]) -
                                   TR1(li, lj, lk) * V1(li, lj, lk)) -
                     (dt2 / hd) * sL * (VD(i, j, k) - VD(li, lj, lk)) + dt2 * fl;
                er = T(*sr, i, j, k) -
                     (dt2 / h1) * (TR1(i + f1[0], j + f1[1], k + f1[2]) *
                                       V1(i + f1[0], j + f1[1], k + f1[2]) -
                                   TR1(i, j, k) * V1(i, j, k)) -
                     (dt2 / hd) * sR * (VD(i + e[0], j + e[1], k + e[2]) - VD(i, j, k)) +
                     dt2 * fr;
            } else {
                /* :513-520 */
                el = T(*sl, i, j, k) -
                     (dt4 / h1) *
                         (V1(li + f1[0], lj + f1[1], lk + f1[2]) + V1(li, lj, lk)) *
                         (TR1(li + f1[0], lj + f1[1], lk + f1[2]) - TR1(li, lj, lk)) +
                     dt2 * fl;
                er = T(*sr, i, j, k) -
                     (dt4 / h1) * (V1(i + f1[0], j + f1[1], k + f1[2]) + V1(i, j, k)) *
                         (TR1(i + f1[0], j + f1[1], k + f1[2]) - TR1(i, j, k)) +
                     dt2 * fr;
            }
        }
#undef P1
#undef V1
#undef V2
#undef VD
#undef TR1
#undef TR2
        double val = upwind(RD(*vd, i, j, k), c->rel_eps, el, er);
        val = bc_final(I, c->domlo[d], c->domhi[d], bclo, bchi, c->is_vel && comp == d, sL, sR,
                       el, er, val);
        out[(i - sedge->lo[0]) + nx * ((j - sedge->lo[1]) + ny * (long)(k - sedge->lo[2]))] = val;
    }
}

static void fill_ctx(Ctx* c, const MxGeom* g, const MxEdgeParams* p) {
    memset(c, 0, sizeof(*c));
    c->dim = g->dim;
    for (int d = 0; d < 3; ++d) {
        c->valid.lo[d] = d < g->dim ? g->valid_lo[d] : 0;
        c->valid.hi[d] = d < g->dim ? g->valid_hi[d] : 0;
        c->domlo[d] = d < g->dim ? g->dom_lo[d] : 0;
        c->domhi[d] = d < g->dim ? g->dom_hi[d] : 0;
        c->dx[d] = d < g->dim ? g->dx[d] : 0.0;
    }
    c->obx = grow_all(c->valid, 1, g->dim);
    c->dt = p->dt;
    c->rel_eps = p->rel_eps;
    c->ppm_type = p->ppm_type;
    c->slope_order = p->slope_order;
    c->ppm_trace_forces = p->ppm_trace_forces;
    c->is_vel = p->is_vel;
    c->is_cons = p->is_conservative;
}

/* Maestro::MakeEdgeScal, Source/MaestroMakeEdgeScalars.cpp:6-302 (one box, one level) */
int orc_make_edge_scal(const MxGeom* g, const MxFab* state, const MxFab* sedge,
                       const MxFab* umac, const MxFab* force, const MxEdgeParams* p) {
    Ctx c;
    fill_ctx(&c, g, p);
    const int dim = g->dim;
    CView vel[3];
    const CView* velp[3] = {&vel[0], &vel[1], &vel[2]};
    for (int d = 0; d < dim; ++d) vel[d] = cview(&umac[d], 0);

    Tmp Ip[3], Im[3], Ipf[3], Imf[3], sl[3], sr[3], simh[3], tr[3][3], divu;
    for (int d = 0; d < dim; ++d) {
        Ip[d] = tmp_alloc(c.obx); Im[d] = tmp_alloc(c.obx);
        Ipf[d] = tmp_alloc(c.obx); Imf[d] = tmp_alloc(c.obx);
        sl[d] = tmp_alloc(c.obx); sr[d] = tmp_alloc(c.obx); simh[d] = tmp_alloc(c.obx);
        for (int t = 0; t < dim; ++t) tr[d][t] = tmp_alloc(c.obx);
    }
    divu = tmp_alloc(c.obx);

    if (dim == 3 && c.is_cons) {
        /* MakeDivU, :648-660.  QUIRK: the whole sum is divided by dx[0] (:658) */
        for (int k = c.obx.lo[2]; k <= c.obx.hi[2]; ++k)
        for (int j = c.obx.lo[1]; j <= c.obx.hi[1]; ++j)
        for (int i = c.obx.lo[0]; i <= c.obx.hi[0]; ++i) {
            double v = RD(vel[0], i + 1, j, k) - RD(vel[0], i, j, k) + RD(vel[1], i, j + 1, k) -
                       RD(vel[1], i, j, k) + RD(vel[2], i, j, k + 1) - RD(vel[2], i, j, k);
            v /= c.dx[0];
            T(divu, i, j, k) = v;
        }
    }

    for (int n = 0; n < p->num_comp; ++n) {
        const int comp = p->start_scomp + n;
        const int bccomp = p->start_bccomp + n;
        const int* bc = p->bc + 6 * bccomp;
        const CView s = cview(state, comp);
        const CView fv = cview(force, comp);
        for (int d = 0; d < dim; ++d) {
            if (c.ppm_type == 0) {
                slope_dir(&c, d, &s, &Ip[d], bc[d], bc[3 + d]); /* :207-216 */
            } else {
                ppm_dir(&c, d, &s, &vel[d], &Ip[d], &Im[d], bc[d], bc[3 + d], 1); /* :219-221 */
                if (c.ppm_trace_forces == 1) {
                    ppm_dir(&c, d, &fv, &vel[d], &Ipf[d], &Imf[d], bc[d], bc[3 + d], 1);
                }
            }
        }
        for (int d = 0; d < dim; ++d) {
            predictor_dir(&c, d, &s, &vel[d], &Ip[d], &Im[d], &sl[d], &sr[d], &simh[d], bc[d],
                          bc[3 + d], comp);
        }
        if (dim == 3) {
            for (int d = 0; d < 3; ++d)
                for (int t = 0; t < 3; ++t)
                    if (t != d)
                        transverse(&c, d, t, &s, &vel[d], &vel[t], &sl[d], &sr[d], &simh[t],
                                   &divu, &tr[d][t], bc[d], bc[3 + d], comp);
            for (int d = 0; d < 3; ++d) {
                const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
                /* sedge_d uses simh_{t1|t2} and simh_{t2|t1} (:1414-1417 etc.) */
                edges_dir(&c, d, &s, velp, &fv, &Ipf[d], &Imf[d], &sl[d], &sr[d], &tr[t1][t2],
                          &tr[t2][t1], t1, t2, &sedge[d], comp, bc[d], bc[3 + d]);
            }
        } else {
            for (int d = 0; d < 2; ++d) {
                const int t1 = 1 - d;
                edges_dir(&c, d, &s, velp, &fv, &Ipf[d], &Imf[d], &sl[d], &sr[d], &simh[t1],
                          NULL, t1, t1, &sedge[d], comp, bc[d], bc[3 + d]);
            }
        }
    }

    for (int d = 0; d < dim; ++d) {
        free(Ip[d].p); free(Im[d].p); free(Ipf[d].p); free(Imf[d].p);
        free(sl[d].p); free(sr[d].p); free(simh[d].p);
        for (int t = 0; t < dim; ++t) free(tr[d][t].p);
    }
    free(divu.p);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* fluxes                                                                                */
/* ------------------------------------------------------------------------------------ */
static inline double* fab_at(const MxFab* f, int i, int j, int k, int n) {
    const long nx = f->hi[0] - f->lo[0] + 1, ny = f->hi[1] - f->lo[1] + 1,
               nz = f->hi[2] - f->lo[2] + 1;
    return f->p + (i - f->lo[0]) + nx * ((j - f->lo[1]) + ny * ((k - f->lo[2]) + nz * (long)n));
}
static inline double base_at(const MxBase* b, int lev, int r) {
    return b->p[(lev * b->len + r) * b->nvar];
}

/* Maestro::MakeRhoXFlux, Source/MaestroMakeFlux.cpp:18-465 (one box) */
int orc_make_rhoX_flux(const MxGeom* g, const MxFab* sflux, const MxFab* etarhoflux,
                       const MxFab* sedge, const MxFab* umac, const MxBase* rho0_old,
                       const MxBase* rho0_edge_old, const MxFab* r0mac_old,
                       const MxBase* rho0_new, const MxBase* rho0_edge_new,
                       const MxFab* r0mac_new, const MxBase* rho0_predicted_edge,
                       const MxBase* w0, const MxFluxParams* p) {
    const int dim = g->dim, lev = p->lev, vert = dim - 1;
    for (int d = 0; d < dim; ++d) {
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int q = 0; q < dim; ++q) { lo[q] = g->valid_lo[q]; hi[q] = g->valid_hi[q]; }
        hi[d] += 1;
        for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
            const int r = (dim == 3) ? k : j; /* radial index of this face / cell */
            double* frho = fab_at(&sflux[d], i, j, k, p->rho_comp);
            *frho = 0.0; /* :47-51 */
            const double u = *fab_at((MxFab*)&umac[d], i, j, k, 0);
            for (int comp = p->start_comp; comp < p->start_comp + p->num_comp; ++comp) {
                double rho0_edge;
                if (dim == 3 && p->spherical) {
                    /* :65-70 LinComb, then :326-329 */
                    rho0_edge = 0.5 * *fab_at((MxFab*)&r0mac_old[d], i, j, k, 0) +
                                0.5 * *fab_at((MxFab*)&r0mac_new[d], i, j, k, 0);
                } else if (d == vert) {
                    rho0_edge = 0.5 * (base_at(rho0_edge_old, lev, r) + base_at(rho0_edge_new, lev, r));
                } else {
                    rho0_edge = 0.5 * (base_at(rho0_old, lev, r) + base_at(rho0_new, lev, r));
                }
                double* F = fab_at(&sflux[d], i, j, k, comp);
                const double se = *fab_at((MxFab*)&sedge[d], i, j, k, comp);
                const double ser = *fab_at((MxFab*)&sedge[d], i, j, k, p->rho_comp);
                if (p->species_pred_type == MX_PRED_RHOPRIME_AND_X) {
                    *F = u * (rho0_edge + ser) * se;
                } else if (p->species_pred_type == MX_PRED_RHOX) {
                    *F = u * se;
                } else if (p->species_pred_type == MX_PRED_RHO_AND_X) {
                    *F = u * ser * se;
                }
                if (d == vert && !(dim == 3 && p->spherical)) {
                    /* :288-301 */
                    if (p->evolve_base_state && !p->use_exact_base_state) {
                        double* eta = fab_at((MxFab*)etarhoflux, i, j, k, 0);
                        if (comp >= p->first_spec && comp <= p->first_spec + p->num_spec - 1) {
                            *eta += *F;
                        }
                        if (comp == p->first_spec + p->num_spec - 1) {
                            *eta -= base_at(w0, lev, r) * base_at(rho0_predicted_edge, lev, r);
                        }
                    }
                }
                *frho += *F; /* :223 */
            }
        }
    }
    return 0;
}

/* Maestro::MakeRhoHFlux, Source/MaestroMakeFlux.cpp:467-1067 (one box) */
int orc_make_rhoh_flux(const MxGeom* g, const MxFab* sflux, const MxFab* sedge,
                       const MxFab* umac, const MxBase* rho0_old, const MxBase* rho0_edge_old,
                       const MxFab* r0mac_old, const MxBase* rho0_new,
                       const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                       const MxBase* rhoh0_old, const MxBase* rhoh0_edge_old,
                       const MxFab* rh0mac_old, const MxBase* rhoh0_new,
                       const MxBase* rhoh0_edge_new, const MxFab* rh0mac_new,
                       const MxFab* h0mac_old, const MxFab* h0mac_new, const MxFluxParams* p) {
    const int dim = g->dim, lev = p->lev, vert = dim - 1;
    const int ept = p->enthalpy_pred_type, spt = p->species_pred_type;
    const int have_h = ept == MX_PRED_H || ept == MX_PRED_T_THEN_H || ept == MX_PRED_TPRIME_THEN_H;
    const int have_hprime = ept == MX_PRED_HPRIME;
    const int have_rhoh = ept == MX_PRED_RHOH;
    const int RhoC = p->rho_comp, RhoHC = p->rhoh_comp;
    for (int d = 0; d < dim; ++d) {
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int q = 0; q < dim; ++q) { lo[q] = g->valid_lo[q]; hi[q] = g->valid_hi[q]; }
        hi[d] += 1;
        for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
            const int r = (dim == 3) ? k : j;
            double* F = fab_at(&sflux[d], i, j, k, RhoHC);
            const double u = *fab_at((MxFab*)&umac[d], i, j, k, 0);
            const double seh = *fab_at((MxFab*)&sedge[d], i, j, k, RhoHC);
            const double ser = *fab_at((MxFab*)&sedge[d], i, j, k, RhoC);
            if (!(dim == 3 && p->spherical)) {
                /* planar: 2-D :600-669, 3-D :674-786 */
                if (have_h) {
                    if (spt == MX_PRED_RHOPRIME_AND_X) {
                        double rho0_edge =
                            (d == vert)
                                ? 0.5 * (base_at(rho0_edge_old, lev, r) + base_at(rho0_edge_new, lev, r))
                                : 0.5 * (base_at(rho0_old, lev, r) + base_at(rho0_new, lev, r));
                        *F = u * (rho0_edge + ser) * seh;
                    } else if (spt == MX_PRED_RHO_AND_X || spt == MX_PRED_RHOX) {
                        *F = u * ser * seh;
                    }
                } else if (have_rhoh) {
                    *F = u * seh;
                } else if (ept == MX_PRED_RHOHPRIME || ept == MX_PRED_T_THEN_RHOHPRIME) {
                    if (d == vert) {
                        double rhoh0_edge = 0.5 * (base_at(rhoh0_edge_old, lev, r) +
                                                   base_at(rhoh0_edge_new, lev, r));
                        *F = u * (seh + rhoh0_edge); /* :782-784 */
                    } else {
                        double rhoh0_edge =
                            0.5 * (base_at(rhoh0_old, lev, r) + base_at(rhoh0_new, lev, r));
                        *F = u * (rhoh0_edge + seh); /* :706-708 */
                    }
                }
            } else if (p->use_exact_base_state) {
                /* spherical, exact base state :788-858 */
                /* QUIRK: the x-face test is `have_h || have_rhoh` (:797), so predict_rhoh
                 * leaves the x-flux untouched; y and z test `have_h || have_hprime` */
                const int skip = (d == 0) ? (have_h || have_rhoh) : (have_h || have_hprime);
                if (skip) {
                } else if (have_rhoh) {
                    *F = u * seh;
                } else {
                    double rhoh0_edge = 0.5 * *fab_at((MxFab*)&rh0mac_old[d], i, j, k, 0) +
                                        0.5 * *fab_at((MxFab*)&rh0mac_new[d], i, j, k, 0);
                    *F = u * (rhoh0_edge + seh);
                }
            } else {
                /* spherical :860-1021 */
                const double rho0_edge = 0.5 * *fab_at((MxFab*)&r0mac_old[d], i, j, k, 0) +
                                         0.5 * *fab_at((MxFab*)&r0mac_new[d], i, j, k, 0);
                const double h0_edge = 0.5 * *fab_at((MxFab*)&h0mac_old[d], i, j, k, 0) +
                                       0.5 * *fab_at((MxFab*)&h0mac_new[d], i, j, k, 0);
                if (have_h) {
                    if (spt == MX_PRED_RHOPRIME_AND_X) {
                        *F = u * (rho0_edge + ser) * seh;
                    } else if (spt == MX_PRED_RHO_AND_X || spt == MX_PRED_RHOX) {
                        *F = u * ser * seh;
                    }
                } else if (have_hprime) {
                    if (spt == MX_PRED_RHOPRIME_AND_X) {
                        *F = u * (ser + rho0_edge) * (seh + h0_edge);
                    }
                } else if (have_rhoh) {
                    *F = u * seh;
                } else {
                    *F = u * (rho0_edge * h0_edge + seh);
                }
            }
        }
    }
    return 0;
}

/* Maestro::CelltoEdge, Source/MaestroCelltoEdge.cpp:5-55; one level, one chunk [r_lo, r_hi] */
int orc_cell_to_edge(const double* cell, double* edge, int nr, int r_lo, int r_hi) {
    const int lo = r_lo, hi = r_hi + 1;
    for (int r = lo; r <= hi; ++r) {
        if (r == 0) {
            edge[r] = cell[r];
        } else if (r == 1 || r == lo) {
            edge[r] = 0.5 * (cell[r - 1] + cell[r]);
        } else if (r == nr) {
            edge[r] = cell[r - 1];
        } else if (r == nr - 1 || r == hi) {
            edge[r] = 0.5 * (cell[r] + cell[r - 1]);
        } else {
            double tmp = 7.0 / 12.0 * (cell[r] + cell[r - 1]) - 1.0 / 12.0 * (cell[r + 1] + cell[r - 2]);
            double s0min = dmin(cell[r], cell[r - 1]);
            double s0max = dmax(cell[r], cell[r - 1]);
            edge[r] = dmin(dmax(tmp, s0min), s0max);
        }
    }
    return 0;
}

/* w0_force, Source/MaestroMakew0.cpp:241-268 (planar) and :727-738 (spherical) */
static void w0_force_cells(const MxW0Params* p, const double* w0_old, const double* w0,
                           double* w0_force) {
    const double dt_in = p->dt_in, dtold_in = p->dtold_in;
    const double dt_avg = 0.5 * (dt_in + dtold_in);
    for (int r = 0; r < p->nr_fine; ++r) {
        double w0_old_cen = 0.5 * (w0_old[r] + w0_old[r + 1]);
        double w0_new_cen = 0.5 * (w0[r] + w0[r + 1]);
        double w0_avg = 0.5 * (dt_in * w0_old_cen + dtold_in * w0_new_cen) / dt_avg;
        double div_avg = 0.5 *
                         (dt_in * (w0_old[r + 1] - w0_old[r]) + dtold_in * (w0[r + 1] - w0[r])) /
                         dt_avg;
        w0_force[r] = (w0_new_cen - w0_old_cen) / dt_avg + w0_avg * div_avg / p->dr;
    }
}

/* Maestro::Makew0 -> Makew0Planar, Source/MaestroMakew0.cpp:5-58,60-272; max_radial_level 0,
 * one chunk [0, nr_fine-1] (RestrictBase / FillGhostBase are no-ops at one level) */
int orc_make_w0_planar(const MxW0Params* p, const double* w0_old, const double* Sbar,
                       const double* p0_old, const double* p0_new, const double* g1o,
                       const double* g1n, const double* pmp, const double* etarho_cc,
                       double* w0, double* w0_force, double* p0bdot) {
    const int nr = p->nr_fine, cut = p->base_cutoff_density_coord;
    const double dr = p->dr;
    double* psi = (double*)calloc((size_t)nr, sizeof(double));
    double* int1 = (double*)calloc((size_t)nr + 1, sizeof(double));
    for (int r = 0; r <= nr; ++r) w0[r] = 0.0;
    for (int r = 0; r < nr; ++r)
        if (r < cut) psi[r] = etarho_cc[r] * fabs(p->grav_const); /* :132-139 */
    w0[0] = 0.0;
    int1[0] = 0.0;
    for (int r = 1; r <= nr; ++r) { /* :143-173 */
        double gamma1bar_p0_avg = (g1o[r - 1] + g1n[r - 1]) * (p0_old[r - 1] + p0_new[r - 1]) / 4.0;
        double delta_chi_w0 = 0.0;
        if (r < cut) {
            if (p->is_predictor) {
                delta_chi_w0 = p->dpdt_factor * pmp[r - 1] / (g1o[r - 1] * p0_old[r - 1] * p->dt);
            } else {
                delta_chi_w0 += p->dpdt_factor * pmp[r - 1] / (g1n[r - 1] * p0_new[r - 1] * p->dt);
            }
        }
        int1[r] = int1[r - 1] + (1.0 / gamma1bar_p0_avg) * dr;
        w0[r] = w0[r - 1] + Sbar[r - 1] * dr - psi[r - 1] / gamma1bar_p0_avg * dr -
                delta_chi_w0 * dr;
    }
    if (p->add_pb) { /* :175-184 */
        double pb = w0[nr] / int1[nr];
        if (p0bdot) *p0bdot = pb;
        for (int r = 1; r <= nr; ++r) w0[r] -= pb * int1[r];
    }
    w0_force_cells(p, w0_old, w0, w0_force);
    free(psi);
    free(int1);
    return 0;
}

/* Maestro::Makew0 -> Makew0Sphr, Source/MaestroMakew0.cpp:537-740, with MakeGravEdge's
 * spherical branch (Source/MaestroMakeGrav.cpp:303-322) and Tridiag (:942-969) inlined */
int orc_make_w0_sphr(const MxW0Params* p, const double* w0_old, const double* Sbar,
                     const double* rho0_old, const double* rho0_new, const double* p0_old,
                     const double* p0_new, const double* g1o, const double* g1n,
                     const double* pmp, const double* etarho_cc, const double* etarho_ec,
                     const double* rcc, const double* re, double* w0, double* w0_force) {
    const double Gconst = 6.67428e-8; /* Source/Maestro.H:1864 */
    const int nr = p->nr_fine;
    const double dr0 = p->dr, dt_in = p->dt_in;
    double* buf = (double*)calloc((size_t)(11 * (nr + 2)), sizeof(double));
    double *g_nph = buf, *p_nph = g_nph + nr + 2, *rho_nph = p_nph + nr + 2,
           *wS = rho_nph + nr + 2, *grav = wS + nr + 2, *A = grav + nr + 2, *B = A + nr + 2,
           *C = B + nr + 2, *F = C + nr + 2, *u = F + nr + 2, *gam = u + nr + 2;
    for (int r = 0; r < nr; ++r) { /* :595-600 */
        p_nph[r] = 0.5 * (p0_old[r] + p0_new[r]);
        rho_nph[r] = 0.5 * (rho0_old[r] + rho0_new[r]);
        g_nph[r] = 0.5 * (g1o[r] + g1n[r]);
    }
    wS[0] = 0.0;
    for (int r = 1; r <= nr; ++r) { /* :608-622 */
        double volume_discrepancy = rho0_old[r - 1] > p->base_cutoff_density
                                        ? p->dpdt_factor * pmp[r - 1] / dt_in
                                        : 0.0;
        wS[r] = wS[r - 1] + dr0 * Sbar[r - 1] * rcc[r - 1] * rcc[r - 1];
        if (volume_discrepancy != 0.0) {
            wS[r] -= dr0 * volume_discrepancy * rcc[r - 1] * rcc[r - 1] /
                     (g_nph[r - 1] * p_nph[r - 1]);
        }
    }
    for (int r = 1; r <= nr; ++r) wS[r] /= (re[r] * re[r]);
    /* MakeGravEdge, spherical */
    grav[0] = 0.0;
    double mencl = 0.0;
    for (int r = 1; r <= nr; ++r) {
        if (rho_nph[r - 1] > p->base_cutoff_density) {
            mencl += 4.0 / 3.0 * M_PI * (re[r] - re[r - 1]) *
                     (re[r] * re[r] + re[r] * re[r - 1] + re[r - 1] * re[r - 1]) * rho_nph[r - 1];
        }
        grav[r] = -Gconst * mencl / (re[r] * re[r]);
    }
    int max_cutoff = p->base_cutoff_density_coord < nr - 1 ? p->base_cutoff_density_coord : nr - 1;
    for (int r = 1; r <= max_cutoff; ++r) { /* :653-681 */
        A[r] = g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1]);
        A[r] /= dr0 * dr0;
        B[r] = -(g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1]) +
                 g_nph[r] * p_nph[r] / (rcc[r] * rcc[r])) /
               (dr0 * dr0);
        double dpdr = (p_nph[r] - p_nph[r - 1]) / dr0;
        B[r] -= 4.0 * dpdr / (re[r] * re[r] * re[r]);
        C[r] = g_nph[r] * p_nph[r] / (rcc[r] * rcc[r]);
        C[r] /= dr0 * dr0;
        F[r] = 4.0 * dpdr * wS[r] / re[r] -
               grav[r] *
                   (rcc[r] * rcc[r] * etarho_cc[r] - rcc[r - 1] * rcc[r - 1] * etarho_cc[r - 1]) /
                   (dr0 * re[r] * re[r]) -
               4.0 * M_PI * Gconst * 0.5 * (rho_nph[r] + rho_nph[r - 1]) * etarho_ec[r];
    }
    A[0] = 0.0; B[0] = 1.0; C[0] = 0.0; F[0] = 0.0;
    A[max_cutoff + 1] = -1.0; B[max_cutoff + 1] = 1.0; C[max_cutoff + 1] = 0.0;
    F[max_cutoff + 1] = 0.0;
    { /* Tridiag(A, B, C, F, u, max_cutoff + 2) */
        const int n = max_cutoff + 2;
        if (B[0] == 0) { free(buf); return -1; }
        double bet = B[0];
        u[0] = F[0] / bet;
        for (int j = 1; j < n; j++) {
            gam[j] = C[j - 1] / bet;
            bet = B[j] - A[j] * gam[j];
            if (bet == 0) { free(buf); return -1; }
            u[j] = (F[j] - A[j] * u[j - 1]) / bet;
        }
        for (int j = n - 2; j >= 0; --j) u[j] -= gam[j + 1] * u[j + 1];
    }
    w0[0] = wS[0];
    for (int r = 1; r <= max_cutoff + 1; ++r) w0[r] = u[r] / (re[r] * re[r]) + wS[r];
    for (int r = max_cutoff + 2; r <= nr; ++r) {
        w0[r] = w0[max_cutoff + 1] * re[max_cutoff + 1] * re[max_cutoff + 1] / (re[r] * re[r]);
    }
    w0_force_cells(p, w0_old, w0, w0_force);
    free(buf);
    return 0;
}

/* Maestro::Makew0 -> Makew0SphrIrreg, Source/MaestroMakew0.cpp:742-940 (use_exact_base_state:
 * irregular radial bins), with MakeGravEdge's spherical branch and Tridiag inlined.
 * w0_force(0): the reference reads r_edge_loc(0,-1) (out of bounds, :926); the first cell's own
 * width stands in here and is not compared by the tests. */
int orc_make_w0_sphr_irreg(const MxW0Params* p, const double* w0_old, const double* Sbar,
                     const double* rho0_old, const double* rho0_new, const double* p0_old,
                     const double* p0_new, const double* g1o, const double* g1n,
                     const double* pmp, const double* etarho_cc, const double* etarho_ec,
                     const double* rcc, const double* re, double* w0, double* w0_force) {
    const double Gconst = 6.67428e-8; /* Source/Maestro.H:1864 */
    const int nr = p->nr_fine;
    const double dt_in = p->dt_in;
    double* buf = (double*)calloc((size_t)(11 * (nr + 2)), sizeof(double));
    double *g_nph = buf, *p_nph = g_nph + nr + 2, *rho_nph = p_nph + nr + 2,
           *wS = rho_nph + nr + 2, *grav = wS + nr + 2, *A = grav + nr + 2, *B = A + nr + 2,
           *C = B + nr + 2, *F = C + nr + 2, *u = F + nr + 2, *gam = u + nr + 2;
    for (int r = 0; r < nr; ++r) { /* :595-600 */
        p_nph[r] = 0.5 * (p0_old[r] + p0_new[r]);
        rho_nph[r] = 0.5 * (rho0_old[r] + rho0_new[r]);
        g_nph[r] = 0.5 * (g1o[r] + g1n[r]);
    }
    wS[0] = 0.0;
    for (int r = 1; r <= nr; ++r) { /* :608-622 */
        double volume_discrepancy = rho0_old[r - 1] > p->base_cutoff_density
                                        ? p->dpdt_factor * pmp[r - 1] / dt_in
                                        : 0.0;
        double dr1 = re[r] - re[r - 1];
        wS[r] = wS[r - 1] + dr1 * Sbar[r - 1] * rcc[r - 1] * rcc[r - 1] -
                dr1 * volume_discrepancy * rcc[r - 1] * rcc[r - 1] / (g_nph[r - 1] * p_nph[r - 1]);
    }
    for (int r = 1; r <= nr; ++r) wS[r] /= (re[r] * re[r]);
    /* MakeGravEdge, spherical */
    grav[0] = 0.0;
    double mencl = 0.0;
    for (int r = 1; r <= nr; ++r) {
        if (rho_nph[r - 1] > p->base_cutoff_density) {
            mencl += 4.0 / 3.0 * M_PI * (re[r] - re[r - 1]) *
                     (re[r] * re[r] + re[r] * re[r - 1] + re[r - 1] * re[r - 1]) * rho_nph[r - 1];
        }
        grav[r] = -Gconst * mencl / (re[r] * re[r]);
    }
    int max_cutoff = p->base_cutoff_density_coord < nr - 1 ? p->base_cutoff_density_coord : nr - 1;
    for (int r = 1; r <= max_cutoff; ++r) { /* :843-878 */
        double dr1 = re[r] - re[r - 1];
        double dr2 = re[r + 1] - re[r];
        double dr3 = rcc[r] - rcc[r - 1];
        A[r] = g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1]);
        A[r] /= dr1 * dr3;
        B[r] = -(g_nph[r - 1] * p_nph[r - 1] / (rcc[r - 1] * rcc[r - 1] * dr1) +
                 g_nph[r] * p_nph[r] / (rcc[r] * rcc[r] * dr2)) /
               dr3;
        double dpdr = (p_nph[r] - p_nph[r - 1]) / dr3;
        B[r] -= 4.0 * dpdr / (re[r] * re[r] * re[r]);
        C[r] = g_nph[r] * p_nph[r] / (rcc[r] * rcc[r]);
        C[r] /= dr2 * dr3;
        F[r] = 4.0 * dpdr * wS[r] / re[r] -
               grav[r] *
                   (rcc[r] * rcc[r] * etarho_cc[r] - rcc[r - 1] * rcc[r - 1] * etarho_cc[r - 1]) /
                   (dr3 * re[r] * re[r]) -
               4.0 * M_PI * Gconst * 0.5 * (rho_nph[r] + rho_nph[r - 1]) * etarho_ec[r];
    }
    A[0] = 0.0; B[0] = 1.0; C[0] = 0.0; F[0] = 0.0;
    A[max_cutoff + 1] = -1.0; B[max_cutoff + 1] = 1.0; C[max_cutoff + 1] = 0.0;
    F[max_cutoff + 1] = 0.0;
    { /* Tridiag(A, B, C, F, u, max_cutoff + 2) */
        const int n = max_cutoff + 2;
        if (B[0] == 0) { free(buf); return -1; }
        double bet = B[0];
        u[0] = F[0] / bet;
        for (int j = 1; j < n; j++) {
            gam[j] = C[j - 1] / bet;
            bet = B[j] - A[j] * gam[j];
            if (bet == 0) { free(buf); return -1; }
            u[j] = (F[j] - A[j] * u[j - 1]) / bet;
        }
        for (int j = n - 2; j >= 0; --j) u[j] -= gam[j + 1] * u[j + 1];
    }
    w0[0] = wS[0];
    for (int r = 1; r <= max_cutoff + 1; ++r) w0[r] = u[r] / (re[r] * re[r]) + wS[r];
    for (int r = max_cutoff + 2; r <= nr; ++r) {
        w0[r] = w0[max_cutoff + 1] * re[max_cutoff + 1] * re[max_cutoff + 1] / (re[r] * re[r]);
    }
    { /* :924-938 */
        const double dtold_in = p->dtold_in;
        const double dt_avg = 0.5 * (dt_in + dtold_in);
        for (int r = 0; r < nr; ++r) {
            double dr1 = r > 0 ? re[r] - re[r - 1] : re[1] - re[0];
            double w0_old_cen = 0.5 * (w0_old[r] + w0_old[r + 1]);
            double w0_new_cen = 0.5 * (w0[r] + w0[r + 1]);
            double w0_avg = 0.5 * (dt_in * w0_old_cen + dtold_in * w0_new_cen) / dt_avg;
            double div_avg =
                0.5 * (dt_in * (w0_old[r + 1] - w0_old[r]) + dtold_in * (w0[r + 1] - w0[r])) / dt_avg;
            w0_force[r] = (w0_new_cen - w0_old_cen) / dt_avg + w0_avg * div_avg / dr1;
        }
    }
    free(buf);
    return 0;
}

/* Maestro::UpdateScal, zone loop of one box: Source/MaestroAdvection.cpp:86-251.  The EOS call
 * (:137-159, third-party Microphysics) is not restated: those zones are flagged in eos_mask. */
static inline long fidx(const MxFab* f, int i, int j, int k) {
    long nx = f->hi[0] - f->lo[0] + 1, ny = f->hi[1] - f->lo[1] + 1;
    return (long)(i - f->lo[0]) + nx * ((long)(j - f->lo[1]) + ny * (long)(k - f->lo[2]));
}
static inline long fcomp(const MxFab* f) {
    return (long)(f->hi[0] - f->lo[0] + 1) * (f->hi[1] - f->lo[1] + 1) * (f->hi[2] - f->lo[2] + 1);
}

int orc_update_scal(const MxGeom* g, const MxFab* sold, const MxFab* snew, const MxFab* sflux,
                    const MxFab* force, const MxUpdateParams* p, unsigned char* eos_mask,
                    long long* n_eos) {
    const int dim = g->dim;
    const int klo = dim == 3 ? g->valid_lo[2] : 0, khi = dim == 3 ? g->valid_hi[2] : 0;
    const long so = fcomp(sold), sn = fcomp(snew), fo = fcomp(force);
    const long fx = fcomp(&sflux[0]), fy = fcomp(&sflux[1]), fz = dim == 3 ? fcomp(&sflux[2]) : 0;
    const double dt = p->dt;
    const int nx = g->valid_hi[0] - g->valid_lo[0] + 1, ny = g->valid_hi[1] - g->valid_lo[1] + 1;
    long long cnt = 0;
    if (p->start_comp != p->rhoh_comp && p->start_comp != p->first_spec) return -1; /* :253 */
    for (int k = klo; k <= khi; ++k)
        for (int j = g->valid_lo[1]; j <= g->valid_hi[1]; ++j)
            for (int i = g->valid_lo[0]; i <= g->valid_hi[0]; ++i) {
                const long io = fidx(sold, i, j, k), in = fidx(snew, i, j, k);
                const long ifo = fidx(force, i, j, k);
                const long x0 = fidx(&sflux[0], i, j, k), x1 = fidx(&sflux[0], i + 1, j, k);
                const long y0 = fidx(&sflux[1], i, j, k), y1 = fidx(&sflux[1], i, j + 1, k);
                const long z0 = dim == 3 ? fidx(&sflux[2], i, j, k) : 0;
                const long z1 = dim == 3 ? fidx(&sflux[2], i, j, k + 1) : 0;
                if (p->start_comp == p->rhoh_comp) {
                    const int c = p->rhoh_comp;
                    double divterm = (sflux[0].p[c * fx + x1] - sflux[0].p[c * fx + x0]) / g->dx[0] +
                                     (sflux[1].p[c * fy + y1] - sflux[1].p[c * fy + y0]) / g->dx[1];
                    if (dim == 3)
                        divterm = divterm +
                                  (sflux[2].p[c * fz + z1] - sflux[2].p[c * fz + z0]) / g->dx[2];
                    snew->p[c * sn + in] =
                        sold->p[c * so + io] + dt * (-divterm + force->p[c * fo + ifo]);
                    const int flag = p->do_eos_h_above_cutoff &&
                                     snew->p[p->rho_comp * sn + in] <= p->base_cutoff_density;
                    if (eos_mask) {
                        eos_mask[(long)(i - g->valid_lo[0]) +
                                 nx * ((long)(j - g->valid_lo[1]) + (long)ny * (k - klo))] =
                            (unsigned char)flag;
                    }
                    cnt += flag;
                    continue;
                }
                /* species, :161-183 */
                for (int n = 0; n < p->num_spec; ++n) {
                    const int c = p->first_spec + n;
                    double divterm = (sflux[0].p[c * fx + x1] - sflux[0].p[c * fx + x0]) / g->dx[0];
                    divterm += (sflux[1].p[c * fy + y1] - sflux[1].p[c * fy + y0]) / g->dx[1];
                    if (dim == 3)
                        divterm += (sflux[2].p[c * fz + z1] - sflux[2].p[c * fz + z0]) / g->dx[2];
                    snew->p[c * sn + in] =
                        sold->p[c * so + io] + dt * (-divterm + force->p[c * fo + ifo]);
                }
                /* density, floor, negative species, :186-251 */
                double* rho = &snew->p[p->rho_comp * sn + in];
                *rho = sold->p[p->rho_comp * so + io];
                int has_negative_species = 0;
                for (int c = p->start_comp; c < p->start_comp + p->num_spec; ++c) {
                    *rho += snew->p[c * sn + in] - sold->p[c * so + io];
                    if (snew->p[c * sn + in] < 0.0) has_negative_species = 1;
                }
                if (*rho < 0.5 * p->base_cutoff_density) {
                    for (int c = p->start_comp; c < p->start_comp + p->num_spec; ++c)
                        snew->p[c * sn + in] *= 0.5 * p->base_cutoff_density / *rho;
                    *rho = 0.5 * p->base_cutoff_density;
                }
                if (has_negative_species) {
                    for (int c = p->start_comp; c < p->start_comp + p->num_spec; ++c) {
                        if (snew->p[c * sn + in] < 0.0) {
                            double delta = -snew->p[c * sn + in];
                            double sumX = 0.0;
                            for (int c2 = p->start_comp; c2 < p->start_comp + p->num_spec; ++c2)
                                if (c2 != c && snew->p[c2 * sn + in] >= 0.0)
                                    sumX += snew->p[c2 * sn + in];
                            for (int c2 = p->start_comp; c2 < p->start_comp + p->num_spec; ++c2)
                                if (c2 != c && snew->p[c2 * sn + in] >= 0.0) {
                                    double frac = snew->p[c2 * sn + in] / sumX;
                                    snew->p[c2 * sn + in] -= frac * delta;
                                }
                            snew->p[c * sn + in] = 0.0;
                        }
                    }
                }
            }
    if (n_eos) *n_eos = cnt;
    return 0;
}

/* ---- multilevel planar Makew0 ---------------------------------------------------------------
 * Maestro::Makew0Planar with max_radial_level > 0 (Source/MaestroMakew0.cpp:60-272) and the
 * RestrictBase / FillGhostBase it calls (Source/MaestroBaseStateGeometry.cpp:236-355).
 * Arrays: level n at offset n*len (len = nr_fine cells or nr_fine+1 edges). */
#define RS(n, j) (rg->r_start_coord[(n) * rg->chunk_len + (j)])
#define RE(n, j) (rg->r_end_coord[(n) * rg->chunk_len + (j)])

static double mc_slope(const double* c, int r) {
    /* MaestroBaseStateGeometry.cpp:281-291: limited slope of the coarse cell r */
    double del = 0.5 * (c[r + 1] - c[r - 1]);
    double dpls = 2.0 * (c[r + 1] - c[r]);
    double dmn = 2.0 * (c[r] - c[r - 1]);
    double slim = dmin(fabs(dpls), fabs(dmn));
    slim = dpls * dmn > 0.0 ? slim : 0.0;
    return copysign(1.0, del) * dmin(slim, fabs(del));
}

static void restrict_base(const MxRadialGeom* rg, double* s0, int len, int cc) {
    for (int n = rg->finest_radial_level; n >= 1; --n)
        for (int i = 1; i <= rg->numdisjointchunks[n]; ++i) {
            double *f = s0 + (long)n * len, *c = s0 + (long)(n - 1) * len;
            if (cc) {
                for (int j = RS(n, i); j < RE(n, i); j += 2) c[j / 2] = 0.5 * (f[j] + f[j + 1]);
            } else {
                for (int j = RS(n, i); j <= RE(n, i) + 1; j += 2) c[j / 2] = f[j];
            }
        }
}

static void fill_ghost_base(const MxRadialGeom* rg, double* s0, int len, int cc) {
    for (int n = rg->finest_radial_level; n >= 1; --n)
        for (int i = 1; i <= rg->numdisjointchunks[n]; ++i) {
            const int lo = RS(n, i), hi = RE(n, i);
            double *f = s0 + (long)n * len, *c = s0 + (long)(n - 1) * len;
            if (cc) {
                if (lo != 0) {
                    int rc = lo / 2 - 1;
                    double sl = mc_slope(c, rc);
                    f[lo - 1] = c[rc] + 0.25 * sl;
                    f[lo - 2] = c[rc] - 0.25 * sl;
                    rc = lo / 2 - 2;
                    sl = mc_slope(c, rc);
                    f[lo - 3] = c[rc] + 0.25 * sl;
                    f[lo - 4] = c[rc] - 0.25 * sl;
                }
                if (hi != rg->nr[n] - 1) {
                    int rc = (hi + 1) / 2;
                    double sl = mc_slope(c, rc);
                    f[hi + 1] = c[rc] - 0.25 * sl;
                    f[hi + 2] = c[rc] + 0.25 * sl;
                    rc = (hi + 1) / 2 + 1;
                    sl = mc_slope(c, rc);
                    f[hi + 3] = c[rc] - 0.25 * sl;
                    f[hi + 4] = c[rc] + 0.25 * sl;
                }
            } else {
                if (lo != 0) {
                    f[lo - 1] = -f[lo + 1] / 3.0 + f[lo] + c[lo / 2 - 1] / 3.0;
                    f[lo - 2] = c[(lo - 2) / 2];
                }
                if (hi + 1 != rg->nr[n]) {
                    f[hi + 2] = -f[hi] / 3.0 + f[hi + 1] + c[(hi + 3) / 2] / 3.0;
                    f[hi + 3] = c[(hi + 3) / 2];
                }
            }
        }
}

int orc_make_w0_planar_ml(const MxW0Params* p, const MxRadialGeom* rg, const double* w0_old,
                          const double* Sbar, const double* p0_old, const double* p0_new,
                          const double* g1o, const double* g1n, const double* pmp,
                          const double* etarho_cc, double* w0, double* w0_force, double* p0bdot) {
    const int nrf = rg->nr_fine, nlev = rg->max_radial_level + 1;
    const int lc = nrf, le = nrf + 1;
    double* psi = (double*)calloc((size_t)nrf + 1, sizeof(double));
    double* int1 = (double*)calloc((size_t)nrf + 2, sizeof(double));
    for (long q = 0; q < (long)nlev * le; ++q) w0[q] = 0.0; /* w0.setVal(0.0), :87 */
    for (long q = 0; q < (long)nlev * lc; ++q) w0_force[q] = 0.0; /* Makew0: w0_force.setVal(0.), :18 */
#define W(n, r) w0[(long)(n) * le + (r)]
#define C(a, n, r) a[(long)(n) * lc + (r)]
    for (int n = 0; n <= rg->max_radial_level; ++n) {
        for (int r = 0; r < nrf; ++r) psi[r] = 0.0;
        const int cut = rg->base_cutoff_density_coord[n];
        const double dr = rg->dr[n];
        for (int j = 1; j <= rg->numdisjointchunks[n]; ++j) {
            if (n == 0) {
                W(0, 0) = 0.0;
                int1[0] = 0.0;
            } else {
                W(n, RS(n, j)) = W(n - 1, RS(n, j) / 2); /* :125-127 */
            }
            for (int r = RS(n, j); r <= RE(n, j); ++r)
                if (r < cut) psi[r] = C(etarho_cc, n, r) * fabs(p->grav_const);
            for (int r = RS(n, j) + 1; r <= RE(n, j) + 1; ++r) {
                double gavg = (C(g1o, n, r - 1) + C(g1n, n, r - 1)) * (C(p0_old, n, r - 1) + C(p0_new, n, r - 1)) / 4.0;
                double dchi = 0.0;
                if (r < cut) {
                    if (p->is_predictor)
                        dchi = p->dpdt_factor * C(pmp, n, r - 1) / (C(g1o, n, r - 1) * C(p0_old, n, r - 1) * p->dt);
                    else
                        dchi += p->dpdt_factor * C(pmp, n, r - 1) / (C(g1n, n, r - 1) * C(p0_new, n, r - 1) * p->dt);
                }
                if (n == 0) int1[r] = int1[r - 1] + (1.0 / gavg) * dr;
                W(n, r) = W(n, r - 1) + C(Sbar, n, r - 1) * dr - psi[r - 1] / gavg * dr - dchi * dr;
            }
            if (n == 0 && p->add_pb) { /* :175-184 */
                const int k = RE(n, j) + 1;
                const double pb = W(n, k) / int1[k];
                if (p0bdot) *p0bdot = pb;
                for (int r = RS(n, j) + 1; r <= RE(n, j) + 1; ++r) W(n, r) -= pb * int1[r];
            }
            if (n > 0) { /* :185-213 */
                const double offset = W(n, RE(n, j) + 1) - W(n - 1, (RE(n, j) + 1) / 2);
                for (int i = n - 1; i >= 0; --i) {
                    const int refrat = 1 << (n - i);
                    /* QUIRK :196: the restriction writes level n, not level i */
                    for (int r = RS(n, j); r <= RE(n, j) + 1; ++r)
                        if (r % refrat == 0) W(n, r / refrat) = W(n, r);
                    for (int r = (RE(n, j) + 1) / refrat + 1; r <= rg->nr[i]; ++r) W(i, r) += offset;
                }
            }
        }
    }
    for (int n = 1; n <= rg->max_radial_level; ++n) /* :216-236 */
        for (int j = 1; j <= rg->numdisjointchunks[n]; ++j) {
            const int lo = RE(n, j) + 2;
            const int hi = j == rg->numdisjointchunks[n] ? rg->nr[n] : RS(n, j + 1) - 1;
            for (int r = lo; r <= hi; ++r) W(n, r) = 0.0;
        }
    restrict_base(rg, w0, le, 0);
    fill_ghost_base(rg, w0, le, 0);
    const double dt_avg = 0.5 * (p->dt_in + p->dtold_in);
    for (int n = 0; n <= rg->max_radial_level; ++n)
        for (int j = 1; j <= rg->numdisjointchunks[n]; ++j)
            for (int r = RS(n, j); r <= RE(n, j); ++r) { /* :241-268 */
                const double* wo = w0_old + (long)n * le;
                double w0_old_cen = 0.5 * (wo[r] + wo[r + 1]);
                double w0_new_cen = 0.5 * (W(n, r) + W(n, r + 1));
                double w0_avg = 0.5 * (p->dt_in * w0_old_cen + p->dtold_in * w0_new_cen) / dt_avg;
                double div_avg = 0.5 * (p->dt_in * (wo[r + 1] - wo[r]) + p->dtold_in * (W(n, r + 1) - W(n, r))) / dt_avg;
                w0_force[(long)n * lc + r] = (w0_new_cen - w0_old_cen) / dt_avg + w0_avg * div_avg / rg->dr[n];
            }
    restrict_base(rg, w0_force, lc, 1);
    fill_ghost_base(rg, w0_force, lc, 1);
#undef W
#undef C
    free(psi);
    free(int1);
    return 0;
}
#undef RS
#undef RE

/* ==================================================================================== */
/* MAC-velocity predictor                                   */
/*   Source/MaestroMakeUtrans.cpp:6-648    -> orc_make_utrans()                           */
/*   Source/MaestroVelPred.cpp:5-1917      -> orc_vel_pred(): VelPredInterface (:296-542  */
/*      2-D, :735-1128 3-D), VelPredTransverse (:1130-1595), VelPredVelocities (:544-733  */
/*      2-D, :1597-1917 3-D)                                                               */
/* Staged like the reference: Ip / Im (or slopes) of every velocity component in every     */
/* direction first, then the interface states, the corner-coupled states, the velocities. */
/* Pinned bit for bit against the reference's own translation units compiled into          */
/* oracle/_ref (tests/test_velpred.py).                                                    */
/* ==================================================================================== */
enum { P_INTERIOR = 0, P_INFLOW = 1, P_OUTFLOW = 2, P_SYMMETRY = 3, P_SLIPWALL = 4, P_NOSLIPWALL = 5 };

typedef struct {
    Ctx c;
    int spherical, trace;
    int pbc_lo[3], pbc_hi[3];
    int bclo[3][3], bchi[3][3]; /* [component][direction] */
} VCtx;

static int fill_vctx(VCtx* v, const MxGeom* g, const MxVelPredParams* p) {
    if (!g || !p || !p->bc) return -1;
    memset(v, 0, sizeof(*v));
    Ctx* c = &v->c;
    c->dim = g->dim;
    for (int d = 0; d < 3; ++d) {
        const int in = d < g->dim;
        c->valid.lo[d] = in ? g->valid_lo[d] : 0;
        c->valid.hi[d] = in ? g->valid_hi[d] : 0;
        c->domlo[d] = in ? g->dom_lo[d] : 0;
        c->domhi[d] = in ? g->dom_hi[d] : 0;
        c->dx[d] = in ? g->dx[d] : 1.0;
    }
    c->obx = grow_all(c->valid, 1, g->dim);
    c->dt = p->dt;
    c->rel_eps = p->rel_eps;
    c->ppm_type = p->ppm_type;
    c->slope_order = p->slope_order;
    v->spherical = p->spherical != 0;
    v->trace = p->ppm_trace_forces != 0;
    for (int d = 0; d < g->dim; ++d) {
        v->pbc_lo[d] = p->phys_bc_lo[d];
        v->pbc_hi[d] = p->phys_bc_hi[d];
        for (int n = 0; n < g->dim; ++n) {
            v->bclo[n][d] = p->bc[6 * n + d];
            v->bchi[n][d] = p->bc[6 * n + 3 + d];
        }
    }
    return 0;
}

/* Slopex/y/z or PPM(is_umac = false) of component n of `s` in direction d over obx */
static void vrecon(const VCtx* v, const MxFab* s, const MxFab* ufull, int n, int d, Tmp* Ip, Tmp* Im) {
    const CView sv = cview(s, n), uv = cview(ufull, d);
    if (v->c.ppm_type == 0)
        slope_dir(&v->c, d, &sv, Ip, v->bclo[n][d], v->bchi[n][d]);
    else
        ppm_dir(&v->c, d, &sv, &uv, Ip, Im, v->bclo[n][d], v->bchi[n][d], 0);
}

/* w0 term of the Riemann problems at face (i,j,k) of direction d */
static double v_w0(const VCtx* v, const MxFab* w0_cart, const MxFab* w0mac, int d, int i, int j, int k) {
    if (v->spherical) return *fab_at(&w0mac[d], i, j, k, 0);
    if (d == v->c.dim - 1) return *fab_at(w0_cart, i, j, k, v->c.dim - 1);
    return 0.0;
}

static double v_riemann(double l, double r, double w, double eps) {
    const int test = (l + w <= 0.0 && r + w >= 0.0) || (fabs(l + r + 2.0 * w) < eps);
    const double val = 0.5 * (l + r) + w > 0.0 ? l : r;
    return test ? 0.0 : val;
}

static double v_upwind(double t, double l, double r, double eps) {
    const double val = t > 0.0 ? l : r;
    return fabs(t) < eps ? 0.5 * (l + r) : val;
}

int orc_make_utrans(const MxGeom* g, const MxFab* utilde, const MxFab* ufull, const MxFab* utrans,
                    const MxFab* w0_cart, const MxFab* w0mac, const MxVelPredParams* p) {
    VCtx v;
    if (fill_vctx(&v, g, p)) return -1;
    const Ctx* c = &v.c;
    const int dim = c->dim;
    const double dt2 = 0.5 * c->dt;
    for (int d = 0; d < dim; ++d) {
        const int e[3] = {d == 0, d == 1, d == 2};
        Tmp Ip = tmp_alloc(c->obx), Im = tmp_alloc(c->obx);
        vrecon(&v, utilde, ufull, d, d, &Ip, &Im);
        IBox fb = c->valid;
        fb.hi[d] += 1;
        for (int k = fb.lo[2]; k <= fb.hi[2]; ++k)
        for (int j = fb.lo[1]; j <= fb.hi[1]; ++j)
        for (int i = fb.lo[0]; i <= fb.hi[0]; ++i) {
            const int il = i - e[0], jl = j - e[1], kl = k - e[2];
            const int I = d == 0 ? i : (d == 1 ? j : k);
            double l, r;
            if (c->ppm_type == 0) { /* MaestroMakeUtrans.cpp:94-103 */
                l = *fab_at(utilde, il, jl, kl, d) +
                    (0.5 - (dt2 / c->dx[d]) * dmax(0.0, *fab_at(ufull, il, jl, kl, d))) * T(Ip, il, jl, kl);
                r = *fab_at(utilde, i, j, k, d) -
                    (0.5 + (dt2 / c->dx[d]) * dmin(0.0, *fab_at(ufull, i, j, k, d))) * T(Ip, i, j, k);
            } else {
                l = T(Ip, il, jl, kl);
                r = T(Im, i, j, k);
            }
            if (I == c->domlo[d]) {
                const int bc = v.pbc_lo[d];
                if (bc == P_INFLOW) {
                    l = r = *fab_at(utilde, il, jl, kl, d);
                } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                    /* QUIRK: the 2-D y sweep zeroes only the right state (:220-221) */
                    if (!(dim == 2 && d == 1)) l = 0.0;
                    r = 0.0;
                } else if (bc == P_OUTFLOW) {
                    l = dmin(r, 0.0);
                    r = l;
                }
            } else if (I == c->domhi[d] + 1) {
                const int bc = v.pbc_hi[d];
                if (bc == P_INFLOW) {
                    /* QUIRK: the 3-D sweeps read one cell further out (:350-351, :466-467, :581-582) */
                    const int q = dim == 3 ? 1 : 0;
                    l = r = *fab_at(utilde, i + q * e[0], j + q * e[1], k + q * e[2], d);
                } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                    l = 0.0;
                    r = 0.0;
                } else if (bc == P_OUTFLOW) {
                    l = dmax(l, 0.0);
                    r = l;
                }
            }
            *fab_at(&utrans[d], i, j, k, 0) =
                v_riemann(l, r, v_w0(&v, w0_cart, w0mac, d, i, j, k), c->rel_eps);
        }
        free(Ip.p);
        free(Im.p);
    }
    return 0;
}

int orc_vel_pred(const MxGeom* g, const MxFab* utilde, const MxFab* ufull, const MxFab* utrans,
                 const MxFab* umac, const MxFab* w0mac, const MxFab* w0_cart, const MxFab* force,
                 const MxVelPredParams* p) {
    VCtx v;
    if (fill_vctx(&v, g, p)) return -1;
    const Ctx* c = &v.c;
    const int dim = c->dim;
    if (v.trace && (dim == 2 || c->ppm_type == 0)) return -2; /* see mx_vel_pred */
    const double dt2 = 0.5 * c->dt, dt4 = 0.25 * c->dt, dt6 = c->dt / 6.0, eps = c->rel_eps;
    const IBox obx = c->obx;
    /* Ip / Im of component n in direction d ([n][d]; ppm_type 0: the slope is in Ip) */
    Tmp Ip[3][3], Im[3][3], Ipf[3], Imf[3], ul[3][3], ur[3][3], imh[3][3], tr[3][3];
    memset(Ipf, 0, sizeof(Ipf));
    memset(Imf, 0, sizeof(Imf));
    memset(tr, 0, sizeof(tr));
    for (int n = 0; n < dim; ++n)
        for (int d = 0; d < dim; ++d) {
            Ip[n][d] = tmp_alloc(obx);
            Im[n][d] = tmp_alloc(obx);
            ul[d][n] = tmp_alloc(obx);
            ur[d][n] = tmp_alloc(obx);
            imh[d][n] = tmp_alloc(obx);
            vrecon(&v, utilde, ufull, n, d, &Ip[n][d], &Im[n][d]);
        }
    if (v.trace)
        for (int d = 0; d < dim; ++d) {
            Ipf[d] = tmp_alloc(obx);
            Imf[d] = tmp_alloc(obx);
            vrecon(&v, force, ufull, d, d, &Ipf[d], &Imf[d]);
        }

    /* ---- VelPredInterface ---- */
    for (int d = 0; d < dim; ++d) {
        const int e[3] = {d == 0, d == 1, d == 2};
        const double h = c->dx[d];
        IBox mb = obx;
        mb.lo[d] += 1;
        CView tv = cview(&utrans[d], 0);
        for (int k = mb.lo[2]; k <= mb.hi[2]; ++k)
        for (int j = mb.lo[1]; j <= mb.hi[1]; ++j)
        for (int i = mb.lo[0]; i <= mb.hi[0]; ++i) {
            const int il = i - e[0], jl = j - e[1], kl = k - e[2];
            const int I = d == 0 ? i : (d == 1 ? j : k);
            double l[3], r[3];
            if (c->ppm_type == 0) {
                const double uL = dmax(0.0, *fab_at(ufull, il, jl, kl, d));
                const double uR = dmin(0.0, *fab_at(ufull, i, j, k, d));
                /* 3-D: :785-796; the 2-D association differs: :336-345 */
                const double fl = dim == 3 ? 0.5 - dt2 * uL / h : 0.5 - (dt2 / h) * uL;
                const double fr = dim == 3 ? 0.5 + dt2 * uR / h : 0.5 + (dt2 / h) * uR;
                for (int n = 0; n < dim; ++n) {
                    l[n] = *fab_at(utilde, il, jl, kl, n) + fl * T(Ip[n][d], il, jl, kl);
                    r[n] = *fab_at(utilde, i, j, k, n) - fr * T(Ip[n][d], i, j, k);
                }
            } else {
                for (int n = 0; n < dim; ++n) {
                    l[n] = T(Ip[n][d], il, jl, kl);
                    r[n] = T(Im[n][d], i, j, k);
                }
            }
            if (I == c->domlo[d]) {
                const int bc = v.pbc_lo[d];
                if (bc == P_INFLOW) {
                    for (int n = 0; n < dim; ++n) l[n] = r[n] = *fab_at(utilde, il, jl, kl, n);
                } else if (bc == P_SLIPWALL || bc == P_SYMMETRY) {
                    for (int n = 0; n < dim; ++n) {
                        if (n == d) {
                            l[n] = 0.0;
                            r[n] = 0.0;
                        } else {
                            l[n] = r[n];
                        }
                    }
                } else if (bc == P_NOSLIPWALL) {
                    for (int n = 0; n < dim; ++n) l[n] = r[n] = 0.0;
                } else if (bc == P_OUTFLOW) {
                    r[d] = dmin(r[d], 0.0);
                    if (dim == 2 && d == 0) { /* QUIRK: the 2-D x sweep copies left over right (:380-385) */
                        for (int n = 0; n < dim; ++n) r[n] = l[n];
                    } else {
                        for (int n = 0; n < dim; ++n) l[n] = r[n];
                    }
                }
            } else if (I == c->domhi[d] + 1) {
                const int bc = v.pbc_hi[d];
                if (bc == P_INFLOW) {
                    for (int n = 0; n < dim; ++n) l[n] = r[n] = *fab_at(utilde, i, j, k, n);
                } else if (bc == P_SLIPWALL || bc == P_SYMMETRY) {
                    for (int n = 0; n < dim; ++n) {
                        if (n == d) {
                            l[n] = 0.0;
                            r[n] = 0.0;
                        } else {
                            r[n] = l[n];
                        }
                    }
                } else if (bc == P_NOSLIPWALL) {
                    for (int n = 0; n < dim; ++n) l[n] = r[n] = 0.0;
                } else if (bc == P_OUTFLOW) {
                    l[d] = dmax(l[d], 0.0);
                    for (int n = 0; n < dim; ++n) r[n] = l[n];
                }
            }
            const double t = RD(tv, i, j, k);
            for (int n = 0; n < dim; ++n) {
                T(ul[d][n], i, j, k) = l[n];
                T(ur[d][n], i, j, k) = r[n];
                if (n != d) T(imh[d][n], i, j, k) = v_upwind(t, l[n], r[n], eps);
            }
        }
    }

    /* ---- VelPredTransverse (3-D): component n on d faces, corrected in direction t ---- */
    if (dim == 3) {
        for (int n = 0; n < 3; ++n)
        for (int d = 0; d < 3; ++d) {
            if (d == n) continue;
            const int t = 3 - n - d;
            const int e[3] = {d == 0, d == 1, d == 2}, q[3] = {t == 0, t == 1, t == 2};
            tr[n][d] = tmp_alloc(obx);
            IBox b = c->valid; /* amrex::grow(tilebox, n, 1), growHi(d, 1) */
            b.lo[n] -= 1;
            b.hi[n] += 1;
            b.hi[d] += 1;
            const CView tt = cview(&utrans[t], 0), td = cview(&utrans[d], 0);
            const double ht = c->dx[t];
            for (int k = b.lo[2]; k <= b.hi[2]; ++k)
            for (int j = b.lo[1]; j <= b.hi[1]; ++j)
            for (int i = b.lo[0]; i <= b.hi[0]; ++i) {
                const int il = i - e[0], jl = j - e[1], kl = k - e[2];
                const int I = d == 0 ? i : (d == 1 ? j : k);
                double l = T(ul[d][n], i, j, k) -
                           (dt6 / ht) * (RD(tt, il + q[0], jl + q[1], kl + q[2]) + RD(tt, il, jl, kl)) *
                               (T(imh[t][n], il + q[0], jl + q[1], kl + q[2]) - T(imh[t][n], il, jl, kl));
                double r = T(ur[d][n], i, j, k) -
                           (dt6 / ht) * (RD(tt, i + q[0], j + q[1], k + q[2]) + RD(tt, i, j, k)) *
                               (T(imh[t][n], i + q[0], j + q[1], k + q[2]) - T(imh[t][n], i, j, k));
                if (I == c->domlo[d]) {
                    const int bc = v.pbc_lo[d];
                    if (bc == P_INFLOW) {
                        l = r = *fab_at(utilde, il, jl, kl, n);
                    } else if (bc == P_SLIPWALL || bc == P_SYMMETRY || bc == P_OUTFLOW) {
                        l = r;
                    } else if (bc == P_NOSLIPWALL) {
                        l = r = 0.0;
                    }
                } else if (I == c->domhi[d] + 1) {
                    const int bc = v.pbc_hi[d];
                    if (bc == P_INFLOW) {
                        l = r = *fab_at(utilde, i, j, k, n);
                    } else if (bc == P_SLIPWALL || bc == P_SYMMETRY || bc == P_OUTFLOW) {
                        r = l;
                    } else if (bc == P_NOSLIPWALL) {
                        l = r = 0.0;
                    }
                }
                T(tr[n][d], i, j, k) = v_upwind(RD(td, i, j, k), l, r, eps);
            }
        }
    }

    /* ---- VelPredVelocities ---- */
    for (int d = 0; d < dim; ++d) {
        const int e[3] = {d == 0, d == 1, d == 2};
        IBox fb = c->valid;
        fb.hi[d] += 1;
        for (int k = fb.lo[2]; k <= fb.hi[2]; ++k)
        for (int j = fb.lo[1]; j <= fb.hi[1]; ++j)
        for (int i = fb.lo[0]; i <= fb.hi[0]; ++i) {
            const int il = i - e[0], jl = j - e[1], kl = k - e[2];
            const int I = d == 0 ? i : (d == 1 ? j : k);
            double fl, fr;
            if (!v.trace) {
                fl = *fab_at(force, il, jl, kl, d);
                fr = *fab_at(force, i, j, k, d);
            } else {
                fl = T(Ipf[d], il, jl, kl);
                fr = T(Imf[d], i, j, k);
            }
            double l = T(ul[d][d], i, j, k), r = T(ur[d][d], i, j, k);
            for (int t = 0; t < dim; ++t) { /* the transverse directions in ascending order */
                if (t == d) continue;
                const int q[3] = {t == 0, t == 1, t == 2};
                const CView tt = cview(&utrans[t], 0);
                const Tmp* a = dim == 3 ? &tr[d][t] : &imh[t][d];
                const double ht = c->dx[t];
                l = l - (dt4 / ht) * (RD(tt, il + q[0], jl + q[1], kl + q[2]) + RD(tt, il, jl, kl)) *
                            (T(*a, il + q[0], jl + q[1], kl + q[2]) - T(*a, il, jl, kl));
                r = r - (dt4 / ht) * (RD(tt, i + q[0], j + q[1], k + q[2]) + RD(tt, i, j, k)) *
                            (T(*a, i + q[0], j + q[1], k + q[2]) - T(*a, i, j, k));
            }
            l = l + dt2 * fl;
            r = r + dt2 * fr;
            double val = v_riemann(l, r, v_w0(&v, w0_cart, w0mac, d, i, j, k), eps);
            if (I == c->domlo[d]) {
                const int bc = v.pbc_lo[d];
                if (bc == P_INFLOW) {
                    val = *fab_at(utilde, il, jl, kl, d);
                } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                    val = 0.0;
                } else if (bc == P_OUTFLOW) {
                    val = dmin(r, 0.0);
                }
            } else if (I == c->domhi[d] + 1) {
                const int bc = v.pbc_hi[d];
                if (bc == P_INFLOW) {
                    val = *fab_at(utilde, i, j, k, d);
                } else if (bc == P_SLIPWALL || bc == P_NOSLIPWALL || bc == P_SYMMETRY) {
                    val = 0.0;
                } else if (bc == P_OUTFLOW) {
                    val = dmax(l, 0.0);
                }
            }
            *fab_at(&umac[d], i, j, k, 0) = val;
        }
    }
    for (int n = 0; n < dim; ++n)
        for (int d = 0; d < dim; ++d) {
            free(Ip[n][d].p);
            free(Im[n][d].p);
            free(ul[d][n].p);
            free(ur[d][n].p);
            free(imh[d][n].p);
            free(tr[n][d].p);
        }
    for (int d = 0; d < 3; ++d) {
        free(Ipf[d].p);
        free(Imf[d].p);
    }
    return 0;
}
