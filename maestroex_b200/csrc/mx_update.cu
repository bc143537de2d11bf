// mx_update.cu — the zone loop of Maestro::UpdateScal (Source/MaestroAdvection.cpp:86-251):
// conservative update of (rho X)_k and rho, or of (rho h), from the face fluxes this library
// produced.  A pure streaming kernel: per zone-component it reads sold, force and the two
// faces per direction (each face is shared with the neighbouring zone: L1/L2) and writes snew,
// i.e. 8 + 8 + 24 + 8 = 48 B per zone-species of compulsory HBM traffic.  Threads run along
// x (coalesced in every array); the species loop runs inside the thread so that the
// density sum (:186-200) keeps the reference's summation order; in 3-D a thread walks UZ
// zones of a z column so that the shared z face is read once.  Expression trees follow the
// reference (-fmad=false); x / dx uses a correctly rounded FMA sequence (div_dx).
#include "mx_update.h"

#include <cmath>

namespace {

struct UArgs {
    KGeom g;
    DFab so, sn, fo, fl[3];
    double dt, bcd;
    double rdx[3];  // RN(1 / dx)
    int pow2[3];    // dx is a power of two: x / dx == x * rdx exactly
    int do_eos, rho, rhoh, first_spec, num_spec, start_comp;
    unsigned char* mask;
    unsigned long long* count;
};

constexpr int UZ = 4;  // zones per thread along z (3-D): the upper z face of one zone is the
                       // lower face of the next, carried in a register

// x / h, correctly rounded (== the reference's IEEE division by dx): a power of two multiplies
// exactly, otherwise q0 = RN(x * RN(1/h)) refined twice with exact FMA residuals (Markstein;
// same sequence as div_h in mx_stream.cuh, checked by tests/test_divconst.py).  The residuals
// are exact only away from underflow, so tiny operands take the plain division.
__device__ __forceinline__ double div_dx(double x, double h, double y, int pow2) {
    double q = x * y;
    if (!pow2) {
        if (fabs(x) < 1.0e-250 && x != 0.0) return x / h;
        double r = __fma_rn(-h, q, x);
        q = __fma_rn(r, y, q);
        r = __fma_rn(-h, q, x);
        q = __fma_rn(r, y, q);
    }
    return q;
}

template <int DIM>
__global__ void __launch_bounds__(128) k_update_scal(const UArgs A) {
    constexpr int NZ = DIM == 3 ? UZ : 1;
    const KGeom& g = A.g;
    const int nx = g.vhi[0] - g.vlo[0] + 1, ny = g.vhi[1] - g.vlo[1] + 1;
    const int i = g.vlo[0] + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = g.vlo[1] + blockIdx.y;
    const int k0 = DIM == 3 ? g.vlo[2] + blockIdx.z * UZ : 0;
    if (i > g.vhi[0]) return;
    const int nk = DIM == 3 ? min(UZ, g.vhi[2] + 1 - k0) : 1;  // zones of this thread
    const long io = A.so.off(i, j, k0), in = A.sn.off(i, j, k0), ifo = A.fo.off(i, j, k0);
    const long x0 = A.fl[0].off(i, j, k0), y0 = A.fl[1].off(i, j, k0);
    const long z0 = DIM == 3 ? A.fl[2].off(i, j, k0) : 0;
    const long so_sz = A.so.sz, sn_sz = A.sn.sz, fo_sz = A.fo.sz;
    const long fx_sz = A.fl[0].sz, fy_sz = A.fl[1].sz, fy_sy = A.fl[1].sy, fz_sz = A.fl[2].sz;
    const double dt = A.dt;
    const double h0 = g.dx[0], h1 = g.dx[1], h2 = g.dx[2];
    const double r0 = A.rdx[0], r1 = A.rdx[1], r2 = A.rdx[2];
    const int p0 = A.pow2[0], p1 = A.pow2[1], p2 = A.pow2[2];

    if (A.start_comp == A.rhoh) {  // :117-135
        const double* fx = A.fl[0].comp(A.rhoh) + x0;
        const double* fy = A.fl[1].comp(A.rhoh) + y0;
        const double* fz = A.fl[2].comp(A.rhoh) + z0;
        const double* so = A.so.comp(A.rhoh) + io;
        const double* fo = A.fo.comp(A.rhoh) + ifo;
        const double* rn = A.sn.comp(A.rho) + in;
        double* sn = A.sn.comp(A.rhoh) + in;
        unsigned nflag = 0;
        const bool full = (nk == NZ);
        double vxl[NZ], vxh[NZ], vyl[NZ], vyh[NZ], vz[NZ + 1], vso[NZ], vfo[NZ], vrn[NZ];
        vz[0] = DIM == 3 ? fz[0] : 0.0;
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
            const bool on = full || q < nk;
            vxl[q] = on ? fx[0] : 0.0;
            vxh[q] = on ? fx[1] : 0.0;
            vyl[q] = on ? fy[0] : 0.0;
            vyh[q] = on ? fy[fy_sy] : 0.0;
            vz[q + 1] = (DIM == 3 && on) ? fz[fz_sz] : 0.0;
            vso[q] = on ? so[0] : 0.0;
            vfo[q] = on ? fo[0] : 0.0;
            vrn[q] = on ? rn[0] : 0.0;
            fx += fx_sz;
            fy += fy_sz;
            fz += fz_sz;
            so += so_sz;
            fo += fo_sz;
            rn += sn_sz;
        }
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
            double divterm = div_dx(vxh[q] - vxl[q], h0, r0, p0) + div_dx(vyh[q] - vyl[q], h1, r1, p1);
            if (DIM == 3) divterm = divterm + div_dx(vz[q + 1] - vz[q], h2, r2, p2);
            if (full || q < nk) {
                sn[0] = vso[q] + dt * (-divterm + vfo[q]);
                // :137-159: the EOS is the caller's; flag the zone
                const bool flag = A.do_eos && vrn[q] <= A.bcd;
                if (A.mask) {
                    A.mask[(long)(i - g.vlo[0]) +
                           (long)nx * ((long)(j - g.vlo[1]) +
                                       (long)ny * (DIM == 3 ? k0 + q - g.vlo[2] : 0))] = flag ? 1 : 0;
                }
                nflag += flag ? 1u : 0u;
            }
            sn += sn_sz;
        }
        if (nflag) atomicAdd(A.count, (unsigned long long)nflag);
        return;
    }

    // species, :161-183, and the density as the sum of the species updates, :186-200.  The
    // species loop is outermost so that every zone accumulates in the reference's order.
    double rho[NZ];
    bool neg[NZ];
    {
        const double* ro = A.so.comp(A.rho) + io;
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
            rho[q] = q < nk ? ro[q * so_sz] : 0.0;
            neg[q] = false;
        }
    }
    const bool full = (nk == NZ);  // block-uniform except in the last z chunk
    for (int n = 0; n < A.num_spec; ++n) {
        const int c = A.first_spec + n;
        const double* fx = A.fl[0].comp(c) + x0;
        const double* fy = A.fl[1].comp(c) + y0;
        const double* fz = A.fl[2].comp(c) + z0;
        const double* so = A.so.comp(c) + io;
        const double* fo = A.fo.comp(c) + ifo;
        double* sn = A.sn.comp(c) + in;
        // all loads of the column first (memory-level parallelism), then the arithmetic
        double vxl[NZ], vxh[NZ], vyl[NZ], vyh[NZ], vz[NZ + 1], vso[NZ], vfo[NZ];
        vz[0] = DIM == 3 ? fz[0] : 0.0;
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
            const bool on = full || q < nk;
            vxl[q] = on ? fx[0] : 0.0;
            vxh[q] = on ? fx[1] : 0.0;
            vyl[q] = on ? fy[0] : 0.0;
            vyh[q] = on ? fy[fy_sy] : 0.0;
            vz[q + 1] = (DIM == 3 && on) ? fz[fz_sz] : 0.0;
            vso[q] = on ? so[0] : 0.0;
            vfo[q] = on ? fo[0] : 0.0;
            fx += fx_sz;
            fy += fy_sz;
            fz += fz_sz;
            so += so_sz;
            fo += fo_sz;
        }
#pragma unroll
        for (int q = 0; q < NZ; ++q) {
            double divterm = div_dx(vxh[q] - vxl[q], h0, r0, p0);
            divterm += div_dx(vyh[q] - vyl[q], h1, r1, p1);
            if (DIM == 3) divterm += div_dx(vz[q + 1] - vz[q], h2, r2, p2);
            const double snew = vso[q] + dt * (-divterm + vfo[q]);
            if (full || q < nk) sn[0] = snew;
            sn += sn_sz;
            rho[q] += snew - vso[q];
            if (snew < 0.0) neg[q] = true;
        }
    }
#pragma unroll
    for (int q = 0; q < NZ; ++q) {
        if (q >= nk) continue;
        const long zq = in + q * sn_sz;
        double r = rho[q];
        // density floor, :214-222
        if (r < 0.5 * A.bcd) {
            for (int n = 0; n < A.num_spec; ++n) {
                double* s = A.sn.comp(A.first_spec + n) + zq;
                *s *= 0.5 * A.bcd / r;
            }
            r = 0.5 * A.bcd;
        }
        A.sn.comp(A.rho)[zq] = r;
        // negative species are zeroed, their deficit taken from the others in proportion, :225-250
        if (neg[q]) {
            for (int n = 0; n < A.num_spec; ++n) {
                double* sc = A.sn.comp(A.first_spec + n) + zq;
                if (*sc < 0.0) {
                    const double delta = -*sc;
                    double sumX = 0.0;
                    for (int m = 0; m < A.num_spec; ++m) {
                        const double v = A.sn.comp(A.first_spec + m)[zq];
                        if (m != n && v >= 0.0) sumX += v;
                    }
                    for (int m = 0; m < A.num_spec; ++m) {
                        double* s2 = A.sn.comp(A.first_spec + m) + zq;
                        if (m != n && *s2 >= 0.0) {
                            const double frac = *s2 / sumX;
                            *s2 -= frac * delta;
                        }
                    }
                    *sc = 0.0;
                }
            }
        }
    }
}

}  // namespace

void update_scal(const KGeom& g, const DFab& sold, const DFab& snew, const DFab* sflux,
                 const DFab& force, const MxUpdateParams& p, unsigned char* mask,
                 unsigned long long* count, cudaStream_t stream) {
    UArgs A;
    A.g = g;
    A.so = sold;
    A.sn = snew;
    A.fo = force;
    for (int d = 0; d < 3; ++d) A.fl[d] = sflux[d < g.dim ? d : 0];
    A.dt = p.dt;
    A.bcd = p.base_cutoff_density;
    A.do_eos = p.do_eos_h_above_cutoff;
    A.rho = p.rho_comp;
    A.rhoh = p.rhoh_comp;
    A.first_spec = p.first_spec;
    A.num_spec = p.num_spec;
    A.start_comp = p.start_comp;
    A.mask = mask;
    A.count = count;
    const int nx = g.vhi[0] - g.vlo[0] + 1, ny = g.vhi[1] - g.vlo[1] + 1;
    const int nz = g.dim == 3 ? g.vhi[2] - g.vlo[2] + 1 : 1;
    for (int d = 0; d < 3; ++d) {
        int e;
        A.rdx[d] = 1.0 / g.dx[d];
        A.pow2[d] = std::frexp(g.dx[d], &e) == 0.5 ? 1 : 0;
    }
    const int tpb = nx >= 96 ? 128 : 64;
    dim3 grid((nx + tpb - 1) / tpb, ny, g.dim == 3 ? (nz + UZ - 1) / UZ : 1);
    if (g.dim == 3)
        k_update_scal<3><<<grid, tpb, 0, stream>>>(A);
    else
        k_update_scal<2><<<grid, tpb, 0, stream>>>(A);
    mx_count_launch();
}
