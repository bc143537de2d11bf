// mx_cart.cu — the passes either side of the edge-state / flux path (SURVEY.md section 8f,
// ranks 2-4): conversions, perturbational form, the rho' forcing, the 1-D -> 3-D base-state
// mappings and the velocity update.  All of them are one-pass, HBM-bound zone loops: one thread
// per zone (or face), i fastest so that a warp reads a contiguous row, no shared memory.
//
// Parity contract as in mx_device.cuh: the reference expression trees are kept (association
// included), the file is compiled -fmad=false, divisions and square roots are the IEEE ones.
//   Source/MaestroConvert.cpp:6-114        ConvertRhoXToX / ConvertRhoHToH / PutInPertForm
//   Source/MaestroFill3dData.cpp:9-1188    Put1dArrayOnCart / Addw0 / MakeW0mac / MakeS0mac
//   Source/MaestroForce.cpp:261-467        ModifyScalForce
//   Source/MaestroAdvection.cpp:299-526    UpdateVel
#include "mx_cart.h"

namespace {

struct Box3 {
    int lo[3], hi[3];
};

template <class F>
__global__ void __launch_bounds__(128) k_box(Box3 b, F f) {
    const int i = b.lo[0] + blockIdx.x * 128 + threadIdx.x;
    const int j = b.lo[1] + blockIdx.y;
    const int k = b.lo[2] + blockIdx.z;
    if (i <= b.hi[0]) f(i, j, k);
}

template <class F>
void for_box(const Box3& b, cudaStream_t st, F f) {
    const int nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1, nz = b.hi[2] - b.lo[2] + 1;
    if (nx <= 0 || ny <= 0 || nz <= 0) return;
    // gridDim.y / .z are limited to 65535: boxes are far smaller in every direction
    dim3 grid((nx + 127) / 128, ny, nz);
    k_box<<<grid, 128, 0, st>>>(b, f);
    mx_count_launch();
}

Box3 valid_box(const KGeom& g) {
    Box3 b;
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = g.vlo[d];
        b.hi[d] = g.vhi[d];
    }
    return b;
}

// mfi.nodaltilebox(dir) of the whole valid box, grown by `ng` (grownnodaltilebox(dir, ng))
Box3 nodal_box(const KGeom& g, int dir, int ng) {
    Box3 b = valid_box(g);
    b.hi[dir] += 1;
    for (int d = 0; d < g.dim; ++d) {
        b.lo[d] -= ng;
        b.hi[d] += ng;
    }
    return b;
}

__device__ __forceinline__ double& at(const DFab& f, int i, int j, int k, int n = 0) {
    return f.p[f.off(i, j, k) + (long)n * f.sn];
}

// std::min / std::max semantics (amrex::min/max)
__device__ __forceinline__ double dmin_(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax_(double a, double b) { return (a < b) ? b : a; }

// QuadInterp, Source/MaestroFill3dData.cpp:339-355, with its default limit = true
// (Source/Maestro.H:65-68; every call site on this path leaves the argument out)
__device__ __forceinline__ double quad_interp(double x, double x0, double x1, double x2, double y0,
                                              double y1, double y2) {
    double y = y0 + (y1 - y0) / (x1 - x0) * (x - x0) +
               ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0)) / (x2 - x0) * (x - x0) * (x - x1);
    const double hi = dmax_(y0, dmax_(y1, y2)), lo = dmin_(y0, dmin_(y1, y2));
    if (y > hi) y = hi;
    if (y < lo) y = lo;
    return y;
}

// edge-centred 1-D array -> value at `radius` (MaestroFill3dData.cpp:218-250, :534-575)
__device__ __forceinline__ double interp_edge(const KSphr& S, const double* s0, int st, int type,
                                              double radius) {
    int index = int(radius / S.drf);
    const int nr = S.nr;
    if (type == 1) {
        const double rfac = (radius - double(index) * S.drf) / S.drf;
        return rfac > 0.5 ? s0[(index + 1) * st] : s0[index * st];
    } else if (type == 2) {
        const double rfac = (radius - double(index) * S.drf) / S.drf;
        if (index < nr) return rfac * s0[(index + 1) * st] + (1.0 - rfac) * s0[index * st];
        return s0[nr * st];
    }
    if (index <= 0) {
        index = 0;
    } else if (index >= nr - 1) {
        index = nr - 2;
    } else if (radius - S.redge[index] < S.redge[index + 1]) {  // (sic)
        index--;
    }
    return quad_interp(radius, S.redge[index], S.redge[index + 1], S.redge[index + 2], s0[index * st],
                       s0[(index + 1) * st], s0[(index + 2) * st]);
}

// cell-centred 1-D array -> value at `radius` (MaestroFill3dData.cpp:262-325, :1010-1180)
__device__ __forceinline__ double interp_cc(const KSphr& S, const double* s0, int st, int type,
                                            double radius) {
    int index = int(radius / S.drf);
    const int nr = S.nr;
    const double drf = S.drf;
    if (type == 1) return s0[index * st];
    if (type == 2) {
        if (radius >= S.rcc[index]) {
            if (index >= nr - 1) return s0[(nr - 1) * st];
            return s0[(index + 1) * st] * (radius - S.rcc[index]) / drf +
                   s0[index * st] * (S.rcc[index + 1] - radius) / drf;
        }
        if (index == 0) return s0[index * st];
        if (index > nr - 1) return s0[(nr - 1) * st];
        return s0[index * st] * (radius - S.rcc[index - 1]) / drf +
               s0[(index - 1) * st] * (S.rcc[index] - radius) / drf;
    }
    if (index == 0) {
        index = 1;
    } else if (index >= nr - 1) {
        index = nr - 2;
    }
    return quad_interp(radius, S.rcc[index - 1], S.rcc[index], S.rcc[index + 1], s0[(index - 1) * st],
                       s0[index * st], s0[(index + 1) * st]);
}

// position of index q along direction d: cell centre (half = 0.5) or face (half = 0)
__device__ __forceinline__ double pos(const KSphr& S, const double* dx, int d, int q, double half) {
    return S.plo[d] + (double(q) + half) * dx[d] - S.cen[d];
}

struct DX3 {
    double v[3];
};

}  // namespace

// ---- ConvertRhoXToX / ConvertRhoHToH -------------------------------------------------------
void cart_convert_by_rho(const KGeom& g, const DFab& s, int rho, int c0, int nc, int flag,
                         cudaStream_t st) {
    for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
        const long o = s.off(i, j, k);
        const double r = s.p[o + (long)rho * s.sn];
        for (int n = c0; n < c0 + nc; ++n) {
            double* q = s.p + o + (long)n * s.sn;
            *q = flag ? *q / r : *q * r;
        }
    });
}

// copy components [c0, c0+nc) of the valid box (MultiFab::Copy with 0 ghost cells)
void cart_copy_valid(const KGeom& g, const DFab& src, const DFab& dst, int c0, int nc, cudaStream_t st) {
    for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
        for (int n = c0; n < c0 + nc; ++n) at(dst, i, j, k, n) = at(src, i, j, k, n);
    });
}

// ---- Put1dArrayOnCart / PutInPertForm ------------------------------------------------------
// mode 0: store; 1: scal -= value; 2: scal += value (PutInPertForm, component `comp`)
void cart_put_1d(const KGeom& g, const double* s0, int st0, const DFab& out, int edge, int vec,
                 int mode, int comp, const KSphr& S, cudaStream_t st) {
    DX3 dx = {{g.dx[0], g.dx[1], g.dx[2]}};
    const int dim = g.dim;
    if (!S.spherical) {
        const int oc = mode == 0 ? (vec ? dim - 1 : 0) : comp;
        for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
            const int r = dim == 2 ? j : k;
            const double v = edge ? 0.5 * (s0[r * st0] + s0[(r + 1) * st0]) : s0[r * st0];
            double& q = at(out, i, j, k, oc);
            q = mode == 0 ? v : (mode == 1 ? q - v : q + v);
        });
        return;
    }
    for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
        const double x = pos(S, dx.v, 0, i, 0.5), y = pos(S, dx.v, 1, j, 0.5), z = pos(S, dx.v, 2, k, 0.5);
        const double radius = sqrt(x * x + y * y + z * z);
        const double v = edge ? interp_edge(S, s0, st0, S.w0t, radius) : interp_cc(S, s0, st0, S.s0t, radius);
        if (mode == 0) {
            if (vec) {
                at(out, i, j, k, 0) = v * x / radius;
                at(out, i, j, k, 1) = v * y / radius;
                at(out, i, j, k, 2) = v * z / radius;
            } else {
                at(out, i, j, k, 0) = v;
            }
        } else {
            double& q = at(out, i, j, k, comp);
            q = mode == 1 ? q - v : q + v;
        }
    });
}

// ---- ModifyScalForce -----------------------------------------------------------------------
__global__ void k_divw0(const double* w0, const double* rcc, const double* redge, int nr, double* out) {
    // MaestroForce.cpp:290-301 (!use_exact_base_state); the last entry stays zero
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    double v = 0.0;
    if (r < nr - 1) {
        const double dr_loc = redge[r + 1] - redge[r];
        v = (redge[r + 1] * redge[r + 1] * w0[r + 1] - redge[r] * redge[r] * w0[r]) /
            (dr_loc * rcc[r] * rcc[r]);
    }
    out[r] = v;
}

void cart_modify_scal_force(const KGeom& g, const DFab& force, const DFab& state, const DFab* umac,
                            const double* s0e, int st_e, const DFab& s0c, const DFab& w0c,
                            const double* w0, double* divw0_ws, int comp, int fullform, int bcf_hi,
                            const DFab* s0ec, const KSphr& S, cudaStream_t st) {
    DX3 dx = {{g.dx[0], g.dx[1], g.dx[2]}};
    const int dim = g.dim;
    const DFab um = umac[0], vm = umac[1], wm = umac[dim == 3 ? 2 : 1];
    if (S.spherical) {
        k_divw0<<<(S.nr + 127) / 128, 128, 0, st>>>(w0, S.rcc, S.redge, S.nr, divw0_ws);
        mx_count_launch();
        const int dlo0 = g.dlo[0], dlo1 = g.dlo[1], dlo2 = g.dlo[2];
        const int dhi0 = g.dhi[0], dhi1 = g.dhi[1], dhi2 = g.dhi[2];
        for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
            const double divumac = (at(um, i + 1, j, k) - at(um, i, j, k)) / dx.v[0] +
                                   (at(vm, i, j + 1, k) - at(vm, i, j, k)) / dx.v[1] +
                                   (at(wm, i, j, k + 1) - at(wm, i, j, k)) / dx.v[2];
            // divu_cart = Put1dArrayOnCart(divw0, ., false, false) at this zone
            const double x = pos(S, dx.v, 0, i, 0.5), y = pos(S, dx.v, 1, j, 0.5), z = pos(S, dx.v, 2, k, 0.5);
            const double radius = sqrt(x * x + y * y + z * z);
            const double divw = interp_cc(S, divw0_ws, 1, S.s0t, radius);
            const double sc = at(state, i, j, k, comp);
            double& f = at(force, i, j, k, comp);
            if (fullform == 1) {
                f -= sc * (divumac + divw);
            } else {
                const double s0 = at(s0c, i, j, k);
                const double s0_xhi = i < dhi0 ? 0.5 * (s0 + at(s0c, i + 1, j, k)) : s0;
                const double s0_xlo = i > dlo0 ? 0.5 * (s0 + at(s0c, i - 1, j, k)) : s0;
                const double s0_yhi = j < dhi1 ? 0.5 * (s0 + at(s0c, i, j + 1, k)) : s0;
                const double s0_ylo = j > dlo1 ? 0.5 * (s0 + at(s0c, i, j - 1, k)) : s0;
                const double s0_zhi = k < dhi2 ? 0.5 * (s0 + at(s0c, i, j, k + 1)) : s0;
                const double s0_zlo = k > dlo2 ? 0.5 * (s0 + at(s0c, i, j, k - 1)) : s0;
                const double divs0u =
                    (at(um, i + 1, j, k) * s0_xhi - at(um, i, j, k) * s0_xlo) / dx.v[0] +
                    (at(vm, i, j + 1, k) * s0_yhi - at(vm, i, j, k) * s0_ylo) / dx.v[1] +
                    (at(wm, i, j, k + 1) * s0_zhi - at(wm, i, j, k) * s0_zlo) / dx.v[2];
                f += -divs0u - (sc - s0) * (divumac + divw);
            }
        });
        return;
    }
    // planar.  s0_edge_cart(., r) = s0_edge(r) inside the box; the ghost cell above the top
    // zone is the neighbouring box's s0_edge(r+1) inside the domain and the bcs_f fill at the
    // top domain face (MaestroForce.cpp:275-277 with Put1dArrayOnCart's FillPatch)
    const int last = dim - 1;
    const int dhi = g.dhi[last], dlo = g.dlo[last];
    const bool have_ec = s0ec != nullptr;
    DFab ec;
    if (have_ec) ec = *s0ec;
    for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
        const int r = dim == 2 ? j : k;
        double e_lo, e_hi;
        if (have_ec) {
            e_lo = at(ec, i, j, k);
            e_hi = dim == 2 ? at(ec, i, j + 1, k) : at(ec, i, j, k + 1);
        } else {
            int rp = r + 1;
            if (r == dhi) rp = (bcf_hi == MX_BC_INT_DIR) ? dlo : dhi;
            e_lo = s0e[r * st_e];
            e_hi = s0e[rp * st_e];
        }
        const double sc = at(state, i, j, k, comp);
        const double s0 = at(s0c, i, j, k);
        double& f = at(force, i, j, k, comp);
        if (dim == 2) {
            double divu = (at(um, i + 1, j, k) - at(um, i, j, k)) / dx.v[0] +
                          (at(vm, i, j + 1, k) - at(vm, i, j, k)) / dx.v[1];
            divu += (at(w0c, i, j + 1, k, 1) - at(w0c, i, j, k, 1)) / dx.v[1];
            if (fullform == 1) {
                f -= sc * divu;
            } else {
                const double divs0u = s0 * (at(um, i + 1, j, k) - at(um, i, j, k)) / dx.v[0] +
                                      (at(vm, i, j + 1, k) * e_hi - at(vm, i, j, k) * e_lo) / dx.v[1];
                f += -(sc - s0) * divu - divs0u;
            }
        } else {
            double divu = (at(um, i + 1, j, k) - at(um, i, j, k)) / dx.v[0] +
                          (at(vm, i, j + 1, k) - at(vm, i, j, k)) / dx.v[1] +
                          (at(wm, i, j, k + 1) - at(wm, i, j, k)) / dx.v[2];
            divu += (at(w0c, i, j, k + 1, 2) - at(w0c, i, j, k, 2)) / dx.v[2];
            if (fullform == 1) {
                f -= sc * divu;
            } else {
                const double divs0u = s0 * ((at(um, i + 1, j, k) - at(um, i, j, k)) / dx.v[0] +
                                            (at(vm, i, j + 1, k) - at(vm, i, j, k)) / dx.v[1]) +
                                      (at(wm, i, j, k + 1) * e_hi - at(wm, i, j, k) * e_lo) / dx.v[2];
                f += -(sc - s0) * divu - divs0u;
            }
        }
    });
}

// ---- Addw0 ---------------------------------------------------------------------------------
void cart_addw0(const KGeom& g, const DFab* ue, const DFab* w0mac, const double* w0, int st_w, double mult,
                int spherical, cudaStream_t st) {
    const int dim = g.dim;
    if (!spherical) {
        const int last = dim - 1;
        Box3 b = nodal_box(g, last, 0);
        for (int d = 0; d < last; ++d) {  // grow(nodaltilebox(last), IntVect(1, [1,] 0))
            b.lo[d] -= 1;
            b.hi[d] += 1;
        }
        const DFab e = ue[last];
        for_box(b, st, [=] __device__(int i, int j, int k) {
            const int r = dim == 2 ? j : k;
            at(e, i, j, k) += mult * w0[r * st_w];
        });
        return;
    }
    for (int d = 0; d < 3; ++d) {
        const DFab e = ue[d], w = w0mac[d];
        for_box(nodal_box(g, d, 0), st,
                [=] __device__(int i, int j, int k) { at(e, i, j, k) += mult * at(w, i, j, k); });
    }
}

// ---- MakeW0mac -----------------------------------------------------------------------------
void cart_make_w0mac(const KGeom& g, const DFab* w0mac, const DFab& w0c, const double* w0, const KSphr& S,
                     cudaStream_t st) {
    DX3 dx = {{g.dx[0], g.dx[1], g.dx[2]}};
    const int type = S.w0mt;
    for (int d = 0; d < 3; ++d) {
        const DFab m = w0mac[d];
        const Box3 b = nodal_box(g, d, 1);
        if (type == 1) {
            for_box(b, st, [=] __device__(int i, int j, int k) {
                at(m, i, j, k) = 0.5 * (at(w0c, i - (d == 0), j - (d == 1), k - (d == 2), d) + at(w0c, i, j, k, d));
            });
        } else if (type == 2 || type == 3) {
            for_box(b, st, [=] __device__(int i, int j, int k) {
                const double x = pos(S, dx.v, 0, i, d == 0 ? 0.0 : 0.5);
                const double y = pos(S, dx.v, 1, j, d == 1 ? 0.0 : 0.5);
                const double z = pos(S, dx.v, 2, k, d == 2 ? 0.0 : 0.5);
                const double radius = sqrt(x * x + y * y + z * z);
                const double v = interp_edge(S, w0, 1, type, radius);
                const double c = d == 0 ? x : (d == 1 ? y : z);
                at(m, i, j, k) = v * c / radius;
            });
        } else {
            // type 4: average of the four nodal values of the face; a nodal value is the linear
            // interpolation of w0 at the node times the unit normal (:499-537)
            for_box(b, st, [=] __device__(int i, int j, int k) {
                auto nodal = [&](int ii, int jj, int kk) {
                    const double x = pos(S, dx.v, 0, ii, 0.0), y = pos(S, dx.v, 1, jj, 0.0),
                                 z = pos(S, dx.v, 2, kk, 0.0);
                    const double radius = sqrt(x * x + y * y + z * z);
                    const double v = interp_edge(S, w0, 1, 2, radius);
                    if (radius == 0.0) return v;
                    const double c = d == 0 ? x : (d == 1 ? y : z);
                    return v * c / radius;
                };
                // the two transverse directions in the reference's order (:683-702)
                const int a = d == 0 ? 1 : 0, bdir = d == 2 ? 1 : 2;
                const int ai = a == 0, aj = a == 1, bj = bdir == 1, bk = bdir == 2;
                at(m, i, j, k) = 0.25 * (nodal(i, j, k) + nodal(i + ai, j + aj, k) +
                                         nodal(i, j + bj, k + bk) + nodal(i + ai, j + aj + bj, k + bk));
            });
        }
    }
}

// ---- MakeS0mac -----------------------------------------------------------------------------
void cart_make_s0mac(const KGeom& g, const double* s0, int st0, const DFab* s0mac, const DFab* s0c,
                     const KSphr& S, cudaStream_t st) {
    DX3 dx = {{g.dx[0], g.dx[1], g.dx[2]}};
    const int type = S.s0mt;
    const int dlo0 = g.dlo[0], dlo1 = g.dlo[1], dlo2 = g.dlo[2];
    const int dhi0 = g.dhi[0], dhi1 = g.dhi[1], dhi2 = g.dhi[2];
    for (int d = 0; d < 3; ++d) {
        const DFab m = s0mac[d];
        const Box3 b = nodal_box(g, d, 1);
        if (type == 1 && s0c != nullptr) {
            const DFab c = *s0c;
            for_box(b, st, [=] __device__(int i, int j, int k) {
                at(m, i, j, k) = 0.5 * (at(c, i - (d == 0), j - (d == 1), k - (d == 2)) + at(c, i, j, k));
            });
        } else if (type == 1) {
            for_box(b, st, [=] __device__(int i, int j, int k) {
                // s0_cart = Put1dArrayOnCart(s0, ., false, false, bcs_f) with 2 ghost cells:
                // outside the domain the first-order extrapolation = nearest cell inside
                auto cart = [&](int ii, int jj, int kk) {
                    ii = min(max(ii, dlo0), dhi0);
                    jj = min(max(jj, dlo1), dhi1);
                    kk = min(max(kk, dlo2), dhi2);
                    const double x = pos(S, dx.v, 0, ii, 0.5), y = pos(S, dx.v, 1, jj, 0.5),
                                 z = pos(S, dx.v, 2, kk, 0.5);
                    return interp_cc(S, s0, st0, S.s0t, sqrt(x * x + y * y + z * z));
                };
                at(m, i, j, k) = 0.5 * (cart(i - (d == 0), j - (d == 1), k - (d == 2)) + cart(i, j, k));
            });
        } else {
            for_box(b, st, [=] __device__(int i, int j, int k) {
                const double x = pos(S, dx.v, 0, i, d == 0 ? 0.0 : 0.5);
                const double y = pos(S, dx.v, 1, j, d == 1 ? 0.0 : 0.5);
                const double z = pos(S, dx.v, 2, k, d == 2 ? 0.0 : 0.5);
                at(m, i, j, k) = interp_cc(S, s0, st0, type, sqrt(x * x + y * y + z * z));
            });
        }
    }
}

// ---- UpdateVel -----------------------------------------------------------------------------
void cart_update_vel(const KGeom& g, const DFab& uold, const DFab& unew, const DFab* umac, const DFab* uedge,
                     const DFab& force, const DFab& sponge, const DFab& w0c, const DFab* w0mac, double dt,
                     int do_sponge, int spherical, cudaStream_t st) {
    DX3 dx = {{g.dx[0], g.dx[1], g.dx[2]}};
    const int dim = g.dim;
    const DFab um = umac[0], vm = umac[1], wm = umac[dim == 3 ? 2 : 1];
    const DFab ex = uedge[0], ey = uedge[1], ez = uedge[dim == 3 ? 2 : 1];
    if (spherical) {
        const DFab wx = w0mac[0], wy = w0mac[1], wz = w0mac[2];
        for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
            const double ubar = 0.5 * (at(um, i, j, k) + at(um, i + 1, j, k));
            const double vbar = 0.5 * (at(vm, i, j, k) + at(vm, i, j + 1, k));
            const double wbar = 0.5 * (at(wm, i, j, k) + at(wm, i, j, k + 1));
            for (int n = 0; n < 3; ++n) {
                const double dxe = at(ex, i + 1, j, k, n) - at(ex, i, j, k, n);
                const double dye = at(ey, i, j + 1, k, n) - at(ey, i, j, k, n);
                const double dze = at(ez, i, j, k + 1, n) - at(ez, i, j, k, n);
                const double ugrad = ubar * dxe / dx.v[0] + vbar * dye / dx.v[1] + wbar * dze / dx.v[2];
                double v = at(uold, i, j, k, n) - dt * ugrad + dt * at(force, i, j, k, n);
                const double gx = dxe / dx.v[0], gy = dye / dx.v[1], gz = dze / dx.v[2];
                // gradux * 0.5 * (w0macx(i) + w0macx(i+1)): (gradux * 0.5) * sum, :487-500
                const double w0g = gx * 0.5 * (at(wx, i, j, k) + at(wx, i + 1, j, k)) +
                                   gy * 0.5 * (at(wy, i, j, k) + at(wy, i, j + 1, k)) +
                                   gz * 0.5 * (at(wz, i, j, k) + at(wz, i, j, k + 1));
                v -= dt * w0g;
                if (do_sponge) v *= at(sponge, i, j, k);
                at(unew, i, j, k, n) = v;
            }
        });
        return;
    }
    for_box(valid_box(g), st, [=] __device__(int i, int j, int k) {
        const double ubar = 0.5 * (at(um, i, j, k) + at(um, i + 1, j, k));
        const double vbar = 0.5 * (at(vm, i, j, k) + at(vm, i, j + 1, k));
        const double wbar = dim == 3 ? 0.5 * (at(wm, i, j, k) + at(wm, i, j, k + 1)) : 0.0;
        const double w0bar = dim == 2 ? 0.5 * (at(w0c, i, j, k, 1) + at(w0c, i, j + 1, k, 1))
                                      : 0.5 * (at(w0c, i, j, k, 2) + at(w0c, i, j, k + 1, 2));
        for (int n = 0; n < dim; ++n) {
            const double dxe = at(ex, i + 1, j, k, n) - at(ex, i, j, k, n);
            const double dye = at(ey, i, j + 1, k, n) - at(ey, i, j, k, n);
            double ugrad = ubar * dxe / dx.v[0] + vbar * dye / dx.v[1];
            double dle = dye;  // difference along the last direction
            if (dim == 3) {
                dle = at(ez, i, j, k + 1, n) - at(ez, i, j, k, n);
                ugrad = ugrad + wbar * dle / dx.v[2];
            }
            double v = at(uold, i, j, k, n) - dt * ugrad + dt * at(force, i, j, k, n);
            v -= dt * w0bar * dle / dx.v[dim - 1];
            if (do_sponge) v *= at(sponge, i, j, k);
            at(unew, i, j, k, n) = v;
        }
    });
}
