// mx_fused.cu — fused single-pass edge-state (+ flux) kernel for sm_100a: the performance
// path for the 3-D, non-conservative scalar sweeps of MakeEdgeScal (ppm_type 0/1/2, every BC
// code except ext_dir, which MAESTROeX never applies to scalars — Source/MaestroBCFill.cpp:67).
//
// One launch replaces, per component, the reference's ~15 ParallelFor launches and 24
// intermediate MultiFab stores (Source/MaestroMakeEdgeScalars.cpp:187-289): cell data goes
// in, edge states (and, in density / enthalpy mode, fluxes) come out; every intermediate
// (parabola integrals, 1-D states, the six corner-coupled states, transverse differences)
// lives in shared memory or registers.
//
// Decomposition: a thread block owns an (x,y) tile of BX x BY valid cells (GX x GY threads,
// one per cell of the tile grown by one) and a chunk of z planes, and marches upward in z as
// a software pipeline ("2.5-D streaming"): iteration `a` loads plane a+LA of the scalar into
// a shared-memory ring, reconstructs the parabolas of plane a in all three directions (the z
// stencil lives in per-thread registers), forms 1-D states and the z-independent corner
// states of plane a, the z-coupled corner states of plane a-1, and emits the x/y edge states
// of plane a-1 and the z edge state of face a.  Components are looped inside the block so
// that velocity planes and base-state entries are re-read from L1/L2, and rho' is swept
// first so that the species fluxes can use its edge state.
//
// Index-space bookkeeping follows the reference exactly, with the tile playing the role of
// the reference's tile box ("index-space extents"): everything is
// recomputed on the tile's one-cell halo instead of being exchanged between blocks.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <functional>
#include <map>
#include <mutex>
#include <vector>
#include <cstdlib>
#include <cstring>

#include <cuda.h>
#include <cstdint>

#include "mx_device.cuh"
#include "mx_fluxdev.cuh"
#include "mx_fused.h"

using namespace mxd;

namespace {

constexpr int MAXC = 40;

struct FusedArgs {
    KGeom g;
    double dt, eps, dt2;
    double c6h[3], c4h[3];  // dt/6/h_d, dt/4/h_d
    double h[3], inv_h[3];
    int inv_exact[3];  // h_d is a power of two: x / h_d == x * inv_h_d exactly
    int slope_order;
    int is_vel;  // velocity sweep: outflow clamp of the normal component at physical faces
    DFab state, force, sedge[3], umac[3], sflux[3], eta;
    DFab r0o[3], r0n[3];
    double* r0e[3];  // spherical: 0.5*(r0mac_old + r0mac_new) laid out like one sedge[d] component
    double* h0e[3];  // spherical enthalpy pass: 0.5*(h0mac_old + h0mac_new), same layout
    int ncomp;
    int comp[MAXC];
    int bclo[MAXC][3], bchi[MAXC][3];
    int mode;  // 0: edge states only, 1: density pass (comp[0] = Rho, then species + fluxes),
               // 2: enthalpy pass (one component + (rho h) flux)
    KFlux f;
    int do_eta;
    int zc;        // z planes per chunk
    int zs_lo, zs_hi;  // planes in the short first / last chunk (0: none), see set_zchunks
    int nzc;           // number of z chunks
    int ntx, nty;  // number of tiles
    int bcdir[3];  // some swept component has a non-interior BC in this direction
    // plane strides in BYTES of umac[0..2], force, sedge[0..2], sflux[0..2], eta: the streaming
    // kernel advances its running pointers by these, which the compiler can then use straight
    // from the constant bank (no LDC + shift per pointer and plane)
    long bs[11];
    alignas(64) CUtensorMap tmap_state;  // (x, y, z, comp) view of `state` for the TMA ring fill
};

// z range [k0, k1) of chunk c.  Where a swept component has a physical BC on a vertical domain
// face the box touches, the four planes next to that face form their own short chunk, so that
// only those blocks run the BC instantiation of the streaming kernels (which is ~20 % slower)
// and every other chunk of the column runs the BC-free one.
__device__ __forceinline__ void chunk_range(const FusedArgs& A, int c, int& k0, int& k1) {
    const int vlo = A.g.vlo[2], vend = A.g.vhi[2] + 1;
    const int mlo = vlo + A.zs_lo, mhi = vend - A.zs_hi;
    if (A.zs_lo && c == 0) {
        k0 = vlo;
        k1 = mlo;
    } else if (A.zs_hi && c == A.nzc - 1) {
        k0 = mhi;
        k1 = vend;
    } else {
        k0 = mlo + (c - (A.zs_lo ? 1 : 0)) * A.zc;
        k1 = min(k0 + A.zc, mhi);
    }
}



#include "mx_stream.cuh"
#include "mx_stream4.cuh"

// chunking of the z range (see chunk_range): short boundary chunks only for k_stream4/5.
// Unless MXADV_ZCHUNK fixes it, the chunk length is chosen by simulating the block schedule
// (one resident block per SM): long chunks amortise the 2.5 planes of per-chunk overhead
// (512^3: 118.7 ms at 63 planes vs 123.0 at 32), short ones balance the last wave (256^3: 15.5 ms
// at 43 vs 16.8 at 32); profiles/r1c_experiments.md.
int sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    return n;
}

// makespan (in plane-iterations) of `tiles` x chunks blocks handed out in launch order
// (blockIdx.x fastest) to `sms` SMs, one resident block each; a chunk of len planes costs
// len + 2.5 (two extra planes + prologue, fitted to the 512^3 z-chunk scan)
double chunk_makespan(long tiles, int sms, const std::vector<int>& lens) {
    std::vector<double> free_at(sms, 0.0);
    std::make_heap(free_at.begin(), free_at.end(), std::greater<double>());
    double end = 0.0;
    for (int len : lens) {
        const double d = len + 2.5;
        for (long t = 0; t < tiles; ++t) {
            std::pop_heap(free_at.begin(), free_at.end(), std::greater<double>());
            const double f = free_at.back() + d;
            free_at.back() = f;
            end = f > end ? f : end;
            std::push_heap(free_at.begin(), free_at.end(), std::greater<double>());
        }
    }
    return end;
}

void set_zchunks(FusedArgs& A, bool allow_short) {
    const int nz = A.g.vhi[2] - A.g.vlo[2] + 1;
    A.zs_lo = A.zs_hi = 0;
    if (allow_short && A.bcdir[2] && nz > 8 + A.zc / 2 && !getenv("MXADV_NO_SHORT_CHUNKS")) {
        if (A.g.vlo[2] <= A.g.dlo[2] + 3) A.zs_lo = 4;
        if (A.g.vhi[2] >= A.g.dhi[2] - 3) A.zs_hi = 4;
    }
    const int mid = nz - A.zs_lo - A.zs_hi;
    const int nshort = (A.zs_lo ? 1 : 0) + (A.zs_hi ? 1 : 0);
    if (allow_short && !getenv("MXADV_ZCHUNK") && mid > 24) {
        // pick the chunk length with the smallest simulated makespan; cached per shape
        const long tiles = (long)A.ntx * A.nty;
        static std::map<std::array<long, 4>, int> cache;
        static std::mutex mtx;
        const std::array<long, 4> key = {tiles, mid, A.zs_lo, A.zs_hi};
        std::lock_guard<std::mutex> lock(mtx);
        auto it = cache.find(key);
        if (it == cache.end()) {
            const int sms = sm_count();
            double best = -1.0;
            int best_zc = A.zc;
            for (int n = (mid + ST_MAXZC - 3) / (ST_MAXZC - 2); n <= (mid + 23) / 24; ++n) {
                const int zc = (mid + n - 1) / n;
                if (zc > ST_MAXZC - 2) continue;
                std::vector<int> lens;
                if (A.zs_lo) lens.push_back(A.zs_lo);
                for (int c = 0, left = mid; c < n && left > 0; ++c, left -= zc)
                    lens.push_back(left < zc ? left : zc);
                if (A.zs_hi) lens.push_back(A.zs_hi);
                const double m = chunk_makespan(tiles, sms, lens);
                if (best < 0.0 || m < best) {
                    best = m;
                    best_zc = zc;
                }
            }
            it = cache.emplace(key, best_zc).first;
        }
        A.zc = it->second;
    }
    A.nzc = (mid + A.zc - 1) / A.zc + nshort;
}

int dispatch(FusedArgs& A, int ppm_type, cudaStream_t stream) {
    const int nx = A.g.vhi[0] - A.g.vlo[0] + 1, ny = A.g.vhi[1] - A.g.vlo[1] + 1;
    A.ntx = (nx + 29) / 30;
    A.nty = (ny + (S4_GY - 2) - 1) / (S4_GY - 2);
    {
        const DFab* f[11] = {&A.umac[0], &A.umac[1], &A.umac[2], &A.force, &A.sedge[0], &A.sedge[1],
                             &A.sedge[2], &A.sflux[0], &A.sflux[1], &A.sflux[2], &A.eta};
        for (int q = 0; q < 11; ++q) A.bs[q] = f[q]->sz * (long)sizeof(double);
    }
    if (A.zc > ST_MAXZC) A.zc = ST_MAXZC;
    set_zchunks(A, true);
    if (ppm_type == 0) return launch_stream4_mode<0>(A, stream);
    if (ppm_type == 1) return launch_stream4_mode<1>(A, stream);
    return launch_stream4_mode<2>(A, stream);
}

bool is_pow2(double h) {
    int e;
    return std::frexp(h, &e) == 0.5;
}

void fill_common(FusedArgs& A, const KGeom& g, const KEdge& e, const DFab& state,
                 const DFab* sedge, const DFab* umac, const DFab& force) {
    memset(&A, 0, sizeof(A));
    A.g = g;
    A.is_vel = e.is_vel;
    A.dt = e.dt;
    A.eps = e.eps;
    A.dt2 = 0.5 * e.dt;
    const double dt4 = 0.25 * e.dt, dt6 = e.dt / 6.0;
    for (int d = 0; d < 3; ++d) {
        A.h[d] = g.dx[d];
        A.inv_h[d] = 1.0 / g.dx[d];
        A.inv_exact[d] = is_pow2(g.dx[d]) ? 1 : 0;
        A.c6h[d] = dt6 / g.dx[d];
        A.c4h[d] = dt4 / g.dx[d];
        A.sedge[d] = sedge[d];
        A.umac[d] = umac[d];
    }
    A.slope_order = e.slope_order;
    A.state = state;
    A.force = force;
    // z-chunk: enough blocks to fill the machine a few times over, small pipeline overhead
    const int nz = g.vhi[2] - g.vlo[2] + 1;
    int zc = 32;
    const char* env = getenv("MXADV_ZCHUNK");
    if (env) zc = atoi(env);
    if (zc < 1) zc = 1;
    A.zc = zc < nz ? zc : nz;
}

void set_bcdir(FusedArgs& A) {
    for (int d = 0; d < 3; ++d) {
        A.bcdir[d] = 0;
        for (int n = 0; n < A.ncomp; ++n)
            if (A.bclo[n][d] != 0 || A.bchi[n][d] != 0) A.bcdir[d] = 1;
    }
}

bool comps_ok(const MxEdgeParams* par, int first_bccomp, int ncomp, bool* any_bc) {
    for (int n = 0; n < ncomp; ++n) {
        const int* bc = par->bc + 6 * (first_bccomp + n);
        for (int q = 0; q < 6; ++q) {
            if (bc[q] == BC_EXT_DIR) return false;
            if (bc[q] != 0) *any_bc = true;
        }
    }
    return true;
}

}  // namespace

bool fused_edge_supported(const KGeom& g, const KEdge& e) {
    if (getenv("MXADV_FORCE_STAGED")) return false;
    return g.dim == 3 && !e.is_cons && !e.trace;
}

int fused_edge_scal(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, const DFab& state,
                    const DFab* sedge, const DFab* umac, const DFab& force, cudaStream_t stream) {
    bool any_bc = false;
    if (!comps_ok(par, par->start_bccomp, par->num_comp, &any_bc)) return 1;  // ext_dir: staged
    for (int c0 = 0; c0 < par->num_comp; c0 += MAXC) {
        FusedArgs A;
        fill_common(A, g, e0, state, sedge, umac, force);
        A.mode = 0;
        A.ncomp = std::min(MAXC, par->num_comp - c0);
        for (int n = 0; n < A.ncomp; ++n) {
            A.comp[n] = par->start_scomp + c0 + n;
            const int* bc = par->bc + 6 * (par->start_bccomp + c0 + n);
            for (int d = 0; d < 3; ++d) {
                A.bclo[n][d] = bc[d];
                A.bchi[n][d] = bc[3 + d];
            }
        }
        set_bcdir(A);
        if (dispatch(A, e0.ppm_type, stream) != 0) return -1;
    }
    return 0;
}

bool fused_density_supported(const KGeom& g, const KEdge& e, const KFlux& f) {
    if (getenv("MXADV_FORCE_STAGED")) return false;
    return g.dim == 3 && !e.is_vel && !e.is_cons && !e.trace && f.num_comp + 1 <= MAXC &&
           (f.spt == MX_PRED_RHOPRIME_AND_X || f.spt == MX_PRED_RHO_AND_X);
}

int fused_density(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, int rho_bccomp,
                  const KFlux& f, const DFab& state, const DFab* sedge, const DFab* sflux,
                  const DFab& etarho, const DFab* umac, const DFab& force, const KMac& mac,
                  double* const* r0e, cudaStream_t stream) {
    bool any_bc = false;
    if (!comps_ok(par, par->start_bccomp, par->num_comp, &any_bc)) return 1;
    if (!comps_ok(par, rho_bccomp, 1, &any_bc)) return 1;
    FusedArgs A;
    fill_common(A, g, e0, state, sedge, umac, force);
    A.mode = 1;
    A.f = f;
    const bool sph = f.spherical != 0;
    A.do_eta = (!sph && f.evolve && !f.exact) ? 1 : 0;
    A.eta = etarho;
    for (int d = 0; d < 3; ++d) {
        A.sflux[d] = sflux[d];
        A.r0o[d] = mac.r0_old[d];
        A.r0n[d] = mac.r0_new[d];
        A.r0e[d] = r0e ? r0e[d] : nullptr;
    }
    A.ncomp = par->num_comp + 1;
    for (int n = 0; n < A.ncomp; ++n) {
        const int bcc = n == 0 ? rho_bccomp : par->start_bccomp + n - 1;
        A.comp[n] = n == 0 ? f.rho : par->start_scomp + n - 1;
        const int* bc = par->bc + 6 * bcc;
        for (int d = 0; d < 3; ++d) {
            A.bclo[n][d] = bc[d];
            A.bchi[n][d] = bc[3 + d];
        }
    }
    set_bcdir(A);
    return dispatch(A, e0.ppm_type, stream);
}

bool fused_enthalpy_supported(const KGeom& g, const KEdge& e, const KFlux& f) {
    if (getenv("MXADV_FORCE_STAGED")) return false;
    // spherical: the use_exact_base_state forms (rhoh0mac, with the reference's x-face quirk)
    // stay on the separate flux kernel
    return g.dim == 3 && !e.is_vel && !e.is_cons && !e.trace && !(f.spherical && f.exact);
}

int fused_enthalpy(const KGeom& g, const KEdge& e0, const MxEdgeParams* par, const KFlux& f,
                   const DFab& state, const DFab* sedge, const DFab* sflux, const DFab* umac,
                   const DFab& force, double* const* r0e, double* const* h0e, cudaStream_t stream) {
    bool any_bc = false;
    if (!comps_ok(par, par->start_bccomp, 1, &any_bc)) return 1;
    FusedArgs A;
    fill_common(A, g, e0, state, sedge, umac, force);
    A.mode = 2;
    A.f = f;
    if (f.spherical && (!r0e || !h0e)) return 1;
    for (int d = 0; d < 3; ++d) {
        A.sflux[d] = sflux[d];
        A.r0e[d] = f.spherical ? r0e[d] : nullptr;
        A.h0e[d] = f.spherical ? h0e[d] : nullptr;
    }
    A.ncomp = 1;
    A.comp[0] = par->start_scomp;
    const int* bc = par->bc + 6 * par->start_bccomp;
    for (int d = 0; d < 3; ++d) {
        A.bclo[0][d] = bc[d];
        A.bchi[0][d] = bc[3 + d];
    }
    set_bcdir(A);
    return dispatch(A, e0.ppm_type, stream);
}
