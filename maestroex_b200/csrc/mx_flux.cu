// mx_flux.cu — Maestro::MakeRhoXFlux / MakeRhoHFlux / CelltoEdge and the ghost-slab
// pack/unpack kernels for sm_100a.  Pure streaming kernels: one thread per face, x fastest
// (coalesced 8-byte accesses), species loop inside the thread exactly like the reference
// (Source/MaestroMakeFlux.cpp:194-197) so that sflux(Rho) and etarhoflux are accumulated in
// the reference's order.
#include "mx_common.h"
#include "mx_device.cuh"
#include "mx_flux.h"
#include "mx_fluxdev.cuh"

namespace {

constexpr int TPB = 256;

struct FaceIdx {
    int i, j, k;
};

__device__ __forceinline__ bool face_of(const KGeom& g, int d, long t, FaceIdx& c) {
    int n[3];
    for (int q = 0; q < 3; ++q) n[q] = g.vhi[q] - g.vlo[q] + 1 + (q == d ? 1 : 0);
    const long tot = (long)n[0] * n[1] * n[2];
    if (t >= tot) return false;
    c.i = g.vlo[0] + (int)(t % n[0]);
    const long r = t / n[0];
    c.j = g.vlo[1] + (int)(r % n[1]);
    c.k = g.vlo[2] + (int)(r / n[1]);
    return true;
}

// Source/MaestroMakeFlux.cpp:189-407 (3-D) and :109-185 (2-D); direction d per launch
__global__ void k_rhoX(KGeom g, KFlux f, int d, DFab sflux, DFab eta, DFab sedge, DFab umac,
                       DFab r0o, DFab r0n) {
    FaceIdx c;
    if (!face_of(g, d, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    const int vert = g.dim - 1;
    const int r = g.dim == 3 ? c.k : c.j;
    const long so = sedge.off(c.i, c.j, c.k);
    const long fo = sflux.off(c.i, c.j, c.k);
    const double u = umac.p[umac.off(c.i, c.j, c.k)];
    const double ser = sedge.comp(f.rho)[so];
    double rho0_edge;
    const bool sph = g.dim == 3 && f.spherical;
    if (sph) {
        rho0_edge = 0.5 * r0o.p[r0o.off(c.i, c.j, c.k)] + 0.5 * r0n.p[r0n.off(c.i, c.j, c.k)];
    } else if (d == vert) {
        rho0_edge = 0.5 * (f.rho0_edge_old[r * f.st_rho0_edge_old] +
                           f.rho0_edge_new[r * f.st_rho0_edge_new]);
    } else {
        rho0_edge = 0.5 * (f.rho0_old[r * f.st_rho0_old] + f.rho0_new[r * f.st_rho0_new]);
    }
    const bool do_eta = (d == vert) && !sph && f.evolve && !f.exact;
    double frho = 0.0;  // :47-51
    double etav = do_eta ? eta.p[eta.off(c.i, c.j, c.k)] : 0.0;
    for (int comp = f.start_comp; comp < f.start_comp + f.num_comp; ++comp) {
        const double se = sedge.comp(comp)[so];
        double F;
        if (f.spt == MX_PRED_RHOPRIME_AND_X) {
            F = u * (rho0_edge + ser) * se;
        } else if (f.spt == MX_PRED_RHOX) {
            F = u * se;
        } else if (f.spt == MX_PRED_RHO_AND_X) {
            F = u * ser * se;
        } else {
            F = sflux.comp(comp)[fo];  // reference leaves the entry untouched
        }
        sflux.comp(comp)[fo] = F;
        if (do_eta) {
            // :288-301
            if (comp >= f.first_spec && comp <= f.first_spec + f.num_spec - 1) etav += F;
            if (comp == f.first_spec + f.num_spec - 1) {
                etav -= f.w0[r * f.st_w0] * f.rho0_pred_edge[r * f.st_rho0_pred_edge];
            }
        }
        frho += F;  // :223
    }
    sflux.comp(f.rho)[fo] = frho;
    if (do_eta) eta.p[eta.off(c.i, c.j, c.k)] = etav;
}

// Source/MaestroMakeFlux.cpp:600-1021; direction d per launch
__global__ void k_rhoh(KGeom g, KFlux f, int d, DFab sflux, DFab sedge, DFab umac, DFab r0o,
                       DFab r0n, DFab rh0o, DFab rh0n, DFab h0o, DFab h0n) {
    FaceIdx c;
    if (!face_of(g, d, blockIdx.x * (long)blockDim.x + threadIdx.x, c)) return;
    const int vert = g.dim - 1;
    const int r = g.dim == 3 ? c.k : c.j;
    const long so = sedge.off(c.i, c.j, c.k);
    const long fo = sflux.off(c.i, c.j, c.k);
    const double u = umac.p[umac.off(c.i, c.j, c.k)];
    const double seh = sedge.comp(f.rhoh)[so];
    const int ept = f.ept, spt = f.spt;
    const bool have_h = ept == MX_PRED_H || ept == MX_PRED_T_THEN_H || ept == MX_PRED_TPRIME_THEN_H;
    const bool have_hprime = ept == MX_PRED_HPRIME;
    const bool have_rhoh = ept == MX_PRED_RHOH;
    double* F = sflux.comp(f.rhoh) + fo;
    const bool sph = g.dim == 3 && f.spherical;
    if (!sph) {
        const double ser = mxd::needs_sedge_rho_for_rhoh(f) ? sedge.comp(f.rho)[so] : 0.0;
        double Fv;
        if (mxd::rhoh_flux_planar(f, d == vert, r, u, ser, seh, &Fv)) *F = Fv;
    } else if (f.exact) {
        // QUIRK: x-face tests `have_h || have_rhoh` (:797); y/z `have_h || have_hprime`
        const bool skip = (d == 0) ? (have_h || have_rhoh) : (have_h || have_hprime);
        if (skip) {
        } else if (have_rhoh) {
            *F = u * seh;
        } else {
            const double rhoh0_edge =
                0.5 * rh0o.p[rh0o.off(c.i, c.j, c.k)] + 0.5 * rh0n.p[rh0n.off(c.i, c.j, c.k)];
            *F = u * (rhoh0_edge + seh);
        }
    } else {
        const double ser = sedge.comp(f.rho)[so];
        const double rho0_edge =
            0.5 * r0o.p[r0o.off(c.i, c.j, c.k)] + 0.5 * r0n.p[r0n.off(c.i, c.j, c.k)];
        const double h0_edge =
            0.5 * h0o.p[h0o.off(c.i, c.j, c.k)] + 0.5 * h0n.p[h0n.off(c.i, c.j, c.k)];
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

// Source/MaestroCelltoEdge.cpp:22-48
__global__ void k_cell_to_edge(const double* __restrict__ cell, double* __restrict__ edge, int nr,
                               int lo, int hi) {
    const int r = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (r > hi) return;
    double v;
    if (r == 0) {
        v = cell[r];
    } else if (r == 1 || r == lo) {
        v = 0.5 * (cell[r - 1] + cell[r]);
    } else if (r == nr) {
        v = cell[r - 1];
    } else if (r == nr - 1 || r == hi) {
        v = 0.5 * (cell[r] + cell[r - 1]);
    } else {
        const double tmp =
            7.0 / 12.0 * (cell[r] + cell[r - 1]) - 1.0 / 12.0 * (cell[r + 1] + cell[r - 2]);
        const double s0min = mxd::dmin_(cell[r], cell[r - 1]);
        const double s0max = mxd::dmax_(cell[r], cell[r - 1]);
        v = mxd::dmin_(mxd::dmax_(tmp, s0min), s0max);
    }
    edge[r] = v;
}

// ghost-slab pack / unpack: slab of depth ng on `side` of direction `dir`; the other two
// directions span the full fab extent (so edge and corner ghosts travel with the faces when
// the exchange is done direction by direction)
__global__ void k_halo(DFab fab, int dir, int start, int ng, int comp0, int ncomp, double* buf,
                       int unpack) {
    int n[3] = {fab.n[0], fab.n[1], fab.n[2]};
    n[dir] = ng;
    const long per = (long)n[0] * n[1] * n[2];
    const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (t >= per * ncomp) return;
    const int c = (int)(t / per);
    long r = t % per;
    int idx[3];
    idx[0] = (int)(r % n[0]);
    r /= n[0];
    idx[1] = (int)(r % n[1]);
    idx[2] = (int)(r / n[1]);
    idx[dir] += start;
    double* p = fab.comp(comp0 + c) + idx[0] + (long)idx[1] * fab.sy + (long)idx[2] * fab.sz;
    if (unpack) *p = buf[t];
    else buf[t] = *p;
}

}  // namespace

void flux_rhoX(const KGeom& g, const KFlux& f, const DFab* sflux, const DFab& etarho,
               const DFab* sedge, const DFab* umac, const KMac& mac, cudaStream_t stream) {
    for (int d = 0; d < g.dim; ++d) {
        long tot = 1;
        for (int q = 0; q < 3; ++q) tot *= g.vhi[q] - g.vlo[q] + 1 + (q == d ? 1 : 0);
        const int nblk = (int)((tot + TPB - 1) / TPB);
        k_rhoX<<<nblk, TPB, 0, stream>>>(g, f, d, sflux[d], etarho, sedge[d], umac[d],
                                        mac.r0_old[d], mac.r0_new[d]);
        mx_count_launch();
    }
}

void flux_rhoh(const KGeom& g, const KFlux& f, const DFab* sflux, const DFab* sedge,
               const DFab* umac, const KMac& mac, cudaStream_t stream) {
    for (int d = 0; d < g.dim; ++d) {
        long tot = 1;
        for (int q = 0; q < 3; ++q) tot *= g.vhi[q] - g.vlo[q] + 1 + (q == d ? 1 : 0);
        const int nblk = (int)((tot + TPB - 1) / TPB);
        k_rhoh<<<nblk, TPB, 0, stream>>>(g, f, d, sflux[d], sedge[d], umac[d], mac.r0_old[d],
                                        mac.r0_new[d], mac.rh0_old[d], mac.rh0_new[d],
                                        mac.h0_old[d], mac.h0_new[d]);
        mx_count_launch();
    }
}

namespace {
// time-centred spherical base density on the faces of direction d, 0.5*old + 0.5*new exactly as
// Source/MaestroMakeFlux.cpp:316-318 forms it per face and species, written once per call in
// the layout of one sedge[d] component so that the fused kernel reads it with the offsets it
// already has
__global__ void k_mac_average(KGeom g, int d, DFab r0o, DFab r0n, DFab like, double* out) {
    const int nx = g.vhi[0] - g.vlo[0] + 1 + (d == 0), ny = g.vhi[1] - g.vlo[1] + 1 + (d == 1);
    const int nz = g.vhi[2] - g.vlo[2] + 1 + (d == 2);
    const long tot = (long)nx * ny * nz;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot;
         t += (long)gridDim.x * blockDim.x) {
        const int i = g.vlo[0] + (int)(t % nx);
        const int j = g.vlo[1] + (int)((t / nx) % ny);
        const int k = g.vlo[2] + (int)(t / ((long)nx * ny));
        out[like.off(i, j, k)] = 0.5 * r0o.p[r0o.off(i, j, k)] + 0.5 * r0n.p[r0n.off(i, j, k)];
    }
}
}  // namespace

void mac_average(const KGeom& g, int d, const DFab& r0o, const DFab& r0n, const DFab& like,
                 double* out, cudaStream_t stream) {
    k_mac_average<<<148 * 8, 256, 0, stream>>>(g, d, r0o, r0n, like, out);
    mx_count_launch();
}

void cell_to_edge(const double* cell, double* edge, int nr, int r_lo, int r_hi,
                  cudaStream_t stream) {
    const int lo = r_lo, hi = r_hi + 1;
    const int n = hi - lo + 1;
    k_cell_to_edge<<<(n + 127) / 128, 128, 0, stream>>>(cell, edge, nr, lo, hi);
    mx_count_launch();
}

void halo_pack(const KGeom& g, const DFab& fab, int dir, int side, int ng, int comp0, int ncomp,
               double* buf, bool unpack, cudaStream_t stream) {
    // index (relative to the fab's low corner) of the first plane of the slab.  A fab that is
    // nodal in `dir` (a MAC velocity / edge array: n+1 faces over n cells) shares its outermost
    // valid face with the neighbour, so the interior slab starts one face further in.
    const int nghost = g.vlo[dir] - fab.lo[dir];
    const int nodal = (fab.n[dir] - (g.vhi[dir] - g.vlo[dir] + 1)) & 1;
    int start;
    if (!unpack) {
        // interior cells adjacent to the face: what the neighbour's ghost region needs
        start = side == 0 ? nghost + nodal : fab.n[dir] - nghost - ng - nodal;
    } else {
        start = side == 0 ? nghost - ng : fab.n[dir] - nghost;
    }
    int n[3] = {fab.n[0], fab.n[1], fab.n[2]};
    n[dir] = ng;
    const long tot = (long)n[0] * n[1] * n[2] * ncomp;
    k_halo<<<(int)((tot + TPB - 1) / TPB), TPB, 0, stream>>>(fab, dir, start, ng, comp0, ncomp,
                                                           buf, unpack ? 1 : 0);
    mx_count_launch();
}
