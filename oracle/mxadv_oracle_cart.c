/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle, second file: the passes either side of the
 * edge-state / flux path.  Plain-C restatement (our own code,
 * flat arrays, one box) of
 *   Source/MaestroConvert.cpp:6-114        -> orc_convert_by_rho(), orc_put_in_pert_form()
 *   Source/MaestroFill3dData.cpp:48-337    -> orc_put_1d_array_on_cart()
 *   Source/MaestroFill3dData.cpp:339-355   -> quad_interp()
 *   Source/MaestroFill3dData.cpp:357-461   -> orc_addw0()
 *   Source/MaestroFill3dData.cpp:464-710   -> orc_make_w0mac()
 *   Source/MaestroFill3dData.cpp:712-1188  -> orc_make_s0mac()   (!use_exact_base_state)
 *   Source/MaestroForce.cpp:261-467        -> orc_modify_scal_force()
 *   Source/MaestroAdvection.cpp:299-526    -> orc_update_vel()
 * Parity status: PINNED — tests/test_cart.py checks every function bit-for-bit against
 * oracle/_ref/libmaestro_ref{2,3}d.so, i.e. the reference's own MaestroConvert.cpp,
 * MaestroFill3dData.cpp, MaestroForce.cpp and MaestroAdvection.cpp compiled unmodified against
 * oracle/shim (oracle/Makefile).  Compile with -ffp-contract=off.
 * The product (libmxadv.so) never links, loads or calls this file.
 */
#include <math.h>
#include <stdlib.h>

#include "../include/mxadv.h"

typedef struct {
    double* p;
    long sy, sz, sn;
    int lo[3];
} V;

static V view(const MxFab* f) {
    V v;
    long nx = f->hi[0] - f->lo[0] + 1, ny = f->hi[1] - f->lo[1] + 1, nz = f->hi[2] - f->lo[2] + 1;
    v.p = f->p;
    v.sy = nx;
    v.sz = nx * ny;
    v.sn = nx * ny * nz;
    for (int d = 0; d < 3; ++d) v.lo[d] = f->lo[d];
    return v;
}
#define A(v, i, j, k, n) \
    ((v).p[((i) - (v).lo[0]) + ((j) - (v).lo[1]) * (v).sy + ((k) - (v).lo[2]) * (v).sz + (long)(n) * (v).sn])

typedef struct {
    int lo[3], hi[3];
} B3;

static B3 valid_box(const MxGeom* g) {
    B3 b;
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = d < g->dim ? g->valid_lo[d] : 0;
        b.hi[d] = d < g->dim ? g->valid_hi[d] : 0;
    }
    return b;
}
static B3 nodal_box(const MxGeom* g, int dir, int ng) {
    B3 b = valid_box(g);
    b.hi[dir] += 1;
    for (int d = 0; d < g->dim; ++d) {
        b.lo[d] -= ng;
        b.hi[d] += ng;
    }
    return b;
}
#define FOR_BOX(b)                                 \
    for (int k = (b).lo[2]; k <= (b).hi[2]; ++k)   \
        for (int j = (b).lo[1]; j <= (b).hi[1]; ++j) \
            for (int i = (b).lo[0]; i <= (b).hi[0]; ++i)

/* std::min / std::max semantics (amrex::min/max) */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }

/* QuadInterp with its default limit = true (Source/Maestro.H:65-68: every call site on this path
 * leaves the argument out), Source/MaestroFill3dData.cpp:339-355 */
static double quad_interp(double x, double x0, double x1, double x2, double y0, double y1, double y2) {
    double y = y0 + (y1 - y0) / (x1 - x0) * (x - x0) +
               ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0)) / (x2 - x0) * (x - x0) * (x - x1);
    if (y > dmax(y0, dmax(y1, y2))) y = dmax(y0, dmax(y1, y2));
    if (y < dmin(y0, dmin(y1, y2))) y = dmin(y0, dmin(y1, y2));
    return y;
}

/* s0(r) of a 1-D array with component stride st */
#define S0(r) (s0[(long)(r) * st])

static double interp_edge(const MxSphr* S, const double* s0, int st, int type, double radius) {
    const double drf = S->dr_fine;
    const int nr = S->nr_fine;
    int index = (int)(radius / drf);
    if (type == 1) {
        double rfac = (radius - (double)index * drf) / drf;
        return rfac > 0.5 ? S0(index + 1) : S0(index);
    }
    if (type == 2) {
        double rfac = (radius - (double)index * drf) / drf;
        if (index < nr) return rfac * S0(index + 1) + (1.0 - rfac) * S0(index);
        return S0(nr);
    }
    if (index <= 0) {
        index = 0;
    } else if (index >= nr - 1) {
        index = nr - 2;
    } else if (radius - S->r_edge_loc[index] < S->r_edge_loc[index + 1]) { /* QUIRK :236 */
        index--;
    }
    return quad_interp(radius, S->r_edge_loc[index], S->r_edge_loc[index + 1], S->r_edge_loc[index + 2],
                       S0(index), S0(index + 1), S0(index + 2));
}

static double interp_cc(const MxSphr* S, const double* s0, int st, int type, double radius) {
    const double drf = S->dr_fine;
    const int nr = S->nr_fine;
    const double* rcc = S->r_cc_loc;
    int index = (int)(radius / drf);
    if (type == 1) return S0(index);
    if (type == 2) {
        if (radius >= rcc[index]) {
            if (index >= nr - 1) return S0(nr - 1);
            return S0(index + 1) * (radius - rcc[index]) / drf + S0(index) * (rcc[index + 1] - radius) / drf;
        }
        if (index == 0) return S0(index);
        if (index > nr - 1) return S0(nr - 1);
        return S0(index) * (radius - rcc[index - 1]) / drf + S0(index - 1) * (rcc[index] - radius) / drf;
    }
    if (index == 0) {
        index = 1;
    } else if (index >= nr - 1) {
        index = nr - 2;
    }
    return quad_interp(radius, rcc[index - 1], rcc[index], rcc[index + 1], S0(index - 1), S0(index),
                       S0(index + 1));
}

static double pos(const MxSphr* S, const MxGeom* g, int d, int q, double half) {
    return S->prob_lo[d] + ((double)q + half) * g->dx[d] - S->center[d];
}

int orc_convert_by_rho(const MxGeom* g, const MxFab* scal, int rho, int c0, int nc, int flag) {
    V s = view(scal);
    B3 b = valid_box(g);
    for (int n = c0; n < c0 + nc; ++n) FOR_BOX(b) {
            if (flag)
                A(s, i, j, k, n) = A(s, i, j, k, n) / A(s, i, j, k, rho);
            else
                A(s, i, j, k, n) = A(s, i, j, k, n) * A(s, i, j, k, rho);
        }
    return 0;
}

/* mode 0 store, 1 subtract, 2 add */
static void put_1d(const MxGeom* g, const MxBase* b0, int lev, const MxFab* out, int edge, int vec, int mode,
                   int comp, const MxSphr* S) {
    V o = view(out);
    B3 b = valid_box(g);
    const int st = b0->nvar;
    const int sph = S && S->spherical;
    const double* s0 = b0->p + (long)(sph ? 0 : lev) * b0->len * b0->nvar;
    if (!sph) {
        const int oc = mode == 0 ? (vec ? g->dim - 1 : 0) : comp;
        FOR_BOX(b) {
            const int r = g->dim == 2 ? j : k;
            const double v = edge ? 0.5 * (S0(r) + S0(r + 1)) : S0(r);
            if (mode == 0)
                A(o, i, j, k, oc) = v;
            else if (mode == 1)
                A(o, i, j, k, oc) = A(o, i, j, k, oc) - v;
            else
                A(o, i, j, k, oc) = A(o, i, j, k, oc) + v;
        }
        return;
    }
    FOR_BOX(b) {
        const double x = pos(S, g, 0, i, 0.5), y = pos(S, g, 1, j, 0.5), z = pos(S, g, 2, k, 0.5);
        const double radius = sqrt(x * x + y * y + z * z);
        const double v = edge ? interp_edge(S, s0, st, S->w0_interp_type, radius)
                              : interp_cc(S, s0, st, S->s0_interp_type, radius);
        if (mode == 0) {
            if (vec) {
                A(o, i, j, k, 0) = v * x / radius;
                A(o, i, j, k, 1) = v * y / radius;
                A(o, i, j, k, 2) = v * z / radius;
            } else {
                A(o, i, j, k, 0) = v;
            }
        } else if (mode == 1) {
            A(o, i, j, k, comp) = A(o, i, j, k, comp) - v;
        } else {
            A(o, i, j, k, comp) = A(o, i, j, k, comp) + v;
        }
    }
}

int orc_put_1d_array_on_cart(const MxGeom* g, const MxBase* s0, int lev, const MxFab* s0_cart, int edge,
                             int vec, const MxSphr* S) {
    put_1d(g, s0, lev, s0_cart, edge, vec, 0, 0, S);
    return 0;
}

int orc_put_in_pert_form(const MxGeom* g, const MxFab* scal, const MxBase* s0, int lev, int comp, int flag,
                         const MxSphr* S) {
    put_1d(g, s0, lev, scal, 0, 0, flag ? 1 : 2, comp, S);
    return 0;
}

int orc_modify_scal_force(const MxGeom* g, const MxFab* scal_force, const MxFab* state, const MxFab* umac,
                          const MxBase* s0_edge, int lev, const MxFab* s0_cart, const MxFab* w0_cart,
                          const double* w0, int comp, int fullform, int bcf_hi, const MxFab* s0_edge_cart,
                          const MxSphr* S) {
    V f = view(scal_force), sc = view(state), s0c = view(s0_cart);
    V um = view(&umac[0]), vm = view(&umac[1]), wm = view(&umac[g->dim == 3 ? 2 : 1]);
    B3 b = valid_box(g);
    const double* dx = g->dx;
    if (S && S->spherical) {
        const int nr = S->nr_fine;
        double* divw0 = (double*)calloc((size_t)nr, sizeof(double));
        for (int r = 0; r < nr - 1; ++r) {
            const double* re = S->r_edge_loc;
            double dr_loc = re[r + 1] - re[r];
            divw0[r] = (re[r + 1] * re[r + 1] * w0[r + 1] - re[r] * re[r] * w0[r]) /
                       (dr_loc * S->r_cc_loc[r] * S->r_cc_loc[r]);
        }
        FOR_BOX(b) {
            double divumac = (A(um, i + 1, j, k, 0) - A(um, i, j, k, 0)) / dx[0] +
                             (A(vm, i, j + 1, k, 0) - A(vm, i, j, k, 0)) / dx[1] +
                             (A(wm, i, j, k + 1, 0) - A(wm, i, j, k, 0)) / dx[2];
            const double x = pos(S, g, 0, i, 0.5), y = pos(S, g, 1, j, 0.5), z = pos(S, g, 2, k, 0.5);
            const double divu = interp_cc(S, divw0, 1, S->s0_interp_type, sqrt(x * x + y * y + z * z));
            if (fullform == 1) {
                A(f, i, j, k, comp) -= A(sc, i, j, k, comp) * (divumac + divu);
            } else {
                const double c = A(s0c, i, j, k, 0);
                double s0_xhi = i < g->dom_hi[0] ? 0.5 * (c + A(s0c, i + 1, j, k, 0)) : c;
                double s0_xlo = i > g->dom_lo[0] ? 0.5 * (c + A(s0c, i - 1, j, k, 0)) : c;
                double s0_yhi = j < g->dom_hi[1] ? 0.5 * (c + A(s0c, i, j + 1, k, 0)) : c;
                double s0_ylo = j > g->dom_lo[1] ? 0.5 * (c + A(s0c, i, j - 1, k, 0)) : c;
                double s0_zhi = k < g->dom_hi[2] ? 0.5 * (c + A(s0c, i, j, k + 1, 0)) : c;
                double s0_zlo = k > g->dom_lo[2] ? 0.5 * (c + A(s0c, i, j, k - 1, 0)) : c;
                double divs0u = (A(um, i + 1, j, k, 0) * s0_xhi - A(um, i, j, k, 0) * s0_xlo) / dx[0] +
                                (A(vm, i, j + 1, k, 0) * s0_yhi - A(vm, i, j, k, 0) * s0_ylo) / dx[1] +
                                (A(wm, i, j, k + 1, 0) * s0_zhi - A(wm, i, j, k, 0) * s0_zlo) / dx[2];
                A(f, i, j, k, comp) += -divs0u - (A(sc, i, j, k, comp) - c) * (divumac + divu);
            }
        }
        free(divw0);
        return 0;
    }
    V w0c = view(w0_cart);
    const int last = g->dim - 1;
    const int st = s0_edge_cart ? 1 : s0_edge->nvar;
    const double* s0 = s0_edge_cart ? NULL : s0_edge->p + (long)lev * s0_edge->len * s0_edge->nvar;
    V ec = s0c;
    if (s0_edge_cart) ec = view(s0_edge_cart);
    FOR_BOX(b) {
        const int r = g->dim == 2 ? j : k;
        /* s0_edge_cart: Put1dArrayOnCart(s0_edge, ., false, false, bcs_f) on a 1-ghost MultiFab */
        double e_lo, e_hi;
        if (s0_edge_cart) {
            e_lo = A(ec, i, j, k, 0);
            e_hi = g->dim == 2 ? A(ec, i, j + 1, k, 0) : A(ec, i, j, k + 1, 0);
        } else {
            int rp = r + 1;
            if (r == g->dom_hi[last]) rp = bcf_hi == MX_BC_INT_DIR ? g->dom_lo[last] : g->dom_hi[last];
            e_lo = S0(r);
            e_hi = S0(rp);
        }
        const double c = A(s0c, i, j, k, 0);
        if (g->dim == 2) {
            double divu = (A(um, i + 1, j, k, 0) - A(um, i, j, k, 0)) / dx[0] +
                          (A(vm, i, j + 1, k, 0) - A(vm, i, j, k, 0)) / dx[1];
            divu += (A(w0c, i, j + 1, k, 1) - A(w0c, i, j, k, 1)) / dx[1];
            if (fullform == 1) {
                A(f, i, j, k, comp) -= A(sc, i, j, k, comp) * divu;
            } else {
                double divs0u = c * (A(um, i + 1, j, k, 0) - A(um, i, j, k, 0)) / dx[0] +
                                (A(vm, i, j + 1, k, 0) * e_hi - A(vm, i, j, k, 0) * e_lo) / dx[1];
                A(f, i, j, k, comp) += -(A(sc, i, j, k, comp) - c) * divu - divs0u;
            }
        } else {
            double divu = (A(um, i + 1, j, k, 0) - A(um, i, j, k, 0)) / dx[0] +
                          (A(vm, i, j + 1, k, 0) - A(vm, i, j, k, 0)) / dx[1] +
                          (A(wm, i, j, k + 1, 0) - A(wm, i, j, k, 0)) / dx[2];
            divu += (A(w0c, i, j, k + 1, 2) - A(w0c, i, j, k, 2)) / dx[2];
            if (fullform == 1) {
                A(f, i, j, k, comp) -= A(sc, i, j, k, comp) * divu;
            } else {
                double divs0u = c * ((A(um, i + 1, j, k, 0) - A(um, i, j, k, 0)) / dx[0] +
                                     (A(vm, i, j + 1, k, 0) - A(vm, i, j, k, 0)) / dx[1]) +
                                (A(wm, i, j, k + 1, 0) * e_hi - A(wm, i, j, k, 0) * e_lo) / dx[2];
                A(f, i, j, k, comp) += -(A(sc, i, j, k, comp) - c) * divu - divs0u;
            }
        }
    }
    return 0;
}

int orc_addw0(const MxGeom* g, const MxFab* uedge, const MxFab* w0mac, const MxBase* w0, int lev, double mult,
              int spherical) {
    if (!spherical) {
        const int last = g->dim - 1;
        B3 b = nodal_box(g, last, 0);
        for (int d = 0; d < last; ++d) {
            b.lo[d] -= 1;
            b.hi[d] += 1;
        }
        V e = view(&uedge[last]);
        const int st = w0->nvar;
        const double* s0 = w0->p + (long)lev * w0->len * w0->nvar;
        FOR_BOX(b) {
            const int r = g->dim == 2 ? j : k;
            A(e, i, j, k, 0) += mult * S0(r);
        }
        return 0;
    }
    for (int d = 0; d < 3; ++d) {
        V e = view(&uedge[d]), w = view(&w0mac[d]);
        B3 b = nodal_box(g, d, 0);
        FOR_BOX(b) A(e, i, j, k, 0) += mult * A(w, i, j, k, 0);
    }
    return 0;
}

static double w0_nodal(const MxSphr* S, const MxGeom* g, const double* w0, int d, int i, int j, int k) {
    const double x = pos(S, g, 0, i, 0.0), y = pos(S, g, 1, j, 0.0), z = pos(S, g, 2, k, 0.0);
    const double radius = sqrt(x * x + y * y + z * z);
    const double v = interp_edge(S, w0, 1, 2, radius);
    if (radius == 0.0) return v;
    return v * (d == 0 ? x : (d == 1 ? y : z)) / radius;
}

int orc_make_w0mac(const MxGeom* g, const MxFab* w0mac, const MxFab* w0_cart, const double* w0,
                   const MxSphr* S) {
    const int type = S->w0mac_interp_type;
    for (int d = 0; d < 3; ++d) {
        V m = view(&w0mac[d]);
        B3 b = nodal_box(g, d, 1);
        if (type == 1) {
            V c = view(w0_cart);
            FOR_BOX(b)
            A(m, i, j, k, 0) = 0.5 * (A(c, i - (d == 0), j - (d == 1), k - (d == 2), d) + A(c, i, j, k, d));
        } else if (type == 2 || type == 3) {
            FOR_BOX(b) {
                const double x = pos(S, g, 0, i, d == 0 ? 0.0 : 0.5);
                const double y = pos(S, g, 1, j, d == 1 ? 0.0 : 0.5);
                const double z = pos(S, g, 2, k, d == 2 ? 0.0 : 0.5);
                const double radius = sqrt(x * x + y * y + z * z);
                const double v = interp_edge(S, w0, 1, type, radius);
                A(m, i, j, k, 0) = v * (d == 0 ? x : (d == 1 ? y : z)) / radius;
            }
        } else {
            FOR_BOX(b) {
                if (d == 0)
                    A(m, i, j, k, 0) = 0.25 * (w0_nodal(S, g, w0, 0, i, j, k) + w0_nodal(S, g, w0, 0, i, j + 1, k) +
                                               w0_nodal(S, g, w0, 0, i, j, k + 1) +
                                               w0_nodal(S, g, w0, 0, i, j + 1, k + 1));
                else if (d == 1)
                    A(m, i, j, k, 0) = 0.25 * (w0_nodal(S, g, w0, 1, i, j, k) + w0_nodal(S, g, w0, 1, i + 1, j, k) +
                                               w0_nodal(S, g, w0, 1, i, j, k + 1) +
                                               w0_nodal(S, g, w0, 1, i + 1, j, k + 1));
                else
                    A(m, i, j, k, 0) = 0.25 * (w0_nodal(S, g, w0, 2, i, j, k) + w0_nodal(S, g, w0, 2, i + 1, j, k) +
                                               w0_nodal(S, g, w0, 2, i, j + 1, k) +
                                               w0_nodal(S, g, w0, 2, i + 1, j + 1, k));
            }
        }
    }
    return 0;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static double s0_cart_at(const MxSphr* S, const MxGeom* g, const double* s0, int st, int i, int j, int k) {
    /* Put1dArrayOnCart(s0, ., false, false, bcs_f) + FillPatch: first-order extrapolation
     * outside the domain (spherical problems have no periodic faces) */
    i = clampi(i, g->dom_lo[0], g->dom_hi[0]);
    j = clampi(j, g->dom_lo[1], g->dom_hi[1]);
    k = clampi(k, g->dom_lo[2], g->dom_hi[2]);
    const double x = pos(S, g, 0, i, 0.5), y = pos(S, g, 1, j, 0.5), z = pos(S, g, 2, k, 0.5);
    return interp_cc(S, s0, st, S->s0_interp_type, sqrt(x * x + y * y + z * z));
}

int orc_make_s0mac(const MxGeom* g, const MxBase* b0, const MxFab* s0mac, const MxFab* s0_cart,
                   const MxSphr* S) {
    const int type = S->s0mac_interp_type;
    const int st = b0->nvar;
    const double* s0 = b0->p;
    for (int d = 0; d < 3; ++d) {
        V m = view(&s0mac[d]);
        B3 b = nodal_box(g, d, 1);
        if (type == 1 && s0_cart) {
            V c = view(s0_cart);
            FOR_BOX(b)
            A(m, i, j, k, 0) = 0.5 * (A(c, i - (d == 0), j - (d == 1), k - (d == 2), 0) + A(c, i, j, k, 0));
        } else if (type == 1) {
            FOR_BOX(b)
            A(m, i, j, k, 0) = 0.5 * (s0_cart_at(S, g, s0, st, i - (d == 0), j - (d == 1), k - (d == 2)) +
                                      s0_cart_at(S, g, s0, st, i, j, k));
        } else {
            FOR_BOX(b) {
                const double x = pos(S, g, 0, i, d == 0 ? 0.0 : 0.5);
                const double y = pos(S, g, 1, j, d == 1 ? 0.0 : 0.5);
                const double z = pos(S, g, 2, k, d == 2 ? 0.0 : 0.5);
                A(m, i, j, k, 0) = interp_cc(S, s0, st, type, sqrt(x * x + y * y + z * z));
            }
        }
    }
    return 0;
}

int orc_update_vel(const MxGeom* g, const MxFab* uold, const MxFab* unew, const MxFab* umac, const MxFab* uedge,
                   const MxFab* force, const MxFab* sponge, const MxFab* w0_cart, const MxFab* w0mac, double dt,
                   int do_sponge, int spherical) {
    const int dim = g->dim;
    V uo = view(uold), un = view(unew), fo = view(force);
    V um = view(&umac[0]), vm = view(&umac[1]), wm = view(&umac[dim == 3 ? 2 : 1]);
    V ex = view(&uedge[0]), ey = view(&uedge[1]), ez = view(&uedge[dim == 3 ? 2 : 1]);
    V sp = uo, w0c = uo, wx = uo, wy = uo, wz = uo;
    if (do_sponge) sp = view(sponge);
    if (spherical) {
        wx = view(&w0mac[0]);
        wy = view(&w0mac[1]);
        wz = view(&w0mac[2]);
    } else {
        w0c = view(w0_cart);
    }
    B3 b = valid_box(g);
    const double* dx = g->dx;
    FOR_BOX(b) {
        const double ubar = 0.5 * (A(um, i, j, k, 0) + A(um, i + 1, j, k, 0));
        const double vbar = 0.5 * (A(vm, i, j, k, 0) + A(vm, i, j + 1, k, 0));
        const double wbar = dim == 3 ? 0.5 * (A(wm, i, j, k, 0) + A(wm, i, j, k + 1, 0)) : 0.0;
        for (int n = 0; n < dim; ++n) {
            double ugrad;
            if (dim == 3)
                ugrad = ubar * (A(ex, i + 1, j, k, n) - A(ex, i, j, k, n)) / dx[0] +
                        vbar * (A(ey, i, j + 1, k, n) - A(ey, i, j, k, n)) / dx[1] +
                        wbar * (A(ez, i, j, k + 1, n) - A(ez, i, j, k, n)) / dx[2];
            else
                ugrad = ubar * (A(ex, i + 1, j, k, n) - A(ex, i, j, k, n)) / dx[0] +
                        vbar * (A(ey, i, j + 1, k, n) - A(ey, i, j, k, n)) / dx[1];
            A(un, i, j, k, n) = A(uo, i, j, k, n) - dt * ugrad + dt * A(fo, i, j, k, n);
        }
        if (!spherical) {
            const double w0bar = dim == 2 ? 0.5 * (A(w0c, i, j, k, 1) + A(w0c, i, j + 1, k, 1))
                                          : 0.5 * (A(w0c, i, j, k, 2) + A(w0c, i, j, k + 1, 2));
            for (int n = 0; n < dim; ++n) {
                if (dim == 2)
                    A(un, i, j, k, n) -= dt * w0bar * (A(ey, i, j + 1, k, n) - A(ey, i, j, k, n)) / dx[1];
                else
                    A(un, i, j, k, n) -= dt * w0bar * (A(ez, i, j, k + 1, n) - A(ez, i, j, k, n)) / dx[2];
                if (do_sponge) A(un, i, j, k, n) *= A(sp, i, j, k, 0);
            }
        } else {
            for (int n = 0; n < 3; ++n) {
                const double gx = (A(ex, i + 1, j, k, n) - A(ex, i, j, k, n)) / dx[0];
                const double gy = (A(ey, i, j + 1, k, n) - A(ey, i, j, k, n)) / dx[1];
                const double gz = (A(ez, i, j, k + 1, n) - A(ez, i, j, k, n)) / dx[2];
                const double w0g = gx * 0.5 * (A(wx, i, j, k, 0) + A(wx, i + 1, j, k, 0)) +
                                   gy * 0.5 * (A(wy, i, j, k, 0) + A(wy, i, j + 1, k, 0)) +
                                   gz * 0.5 * (A(wz, i, j, k, 0) + A(wz, i, j, k + 1, 0));
                A(un, i, j, k, n) -= dt * w0g;
            }
            if (do_sponge)
                for (int n = 0; n < 3; ++n) A(un, i, j, k, n) *= A(sp, i, j, k, 0);
        }
    }
    return 0;
}
