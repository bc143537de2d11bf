// halo2.cpp — a C++ host (no Python, no torch) driving TWO GPUs through the C ABI of
// libmxadv.so: one host thread per GPU, one brick per GPU of a 2x1x1 periodic-in-x domain.
//   1. every thread fills its ghost-grown state / force / MAC-velocity fabs from ONE analytic
//      global field, poisons the ghost layers that face the other brick, restores them with
//      mx_halo_exchange (transport argv[1]: 0 = NCCL, 1 = NVLink P2P) and checks them against
//      the analytic field (periodic wrap included);
//   2. both bricks run mx_make_edge_scal (the fused streaming kernel); the x faces on the brick
//      interface are computed by BOTH bricks and must agree bit for bit.
// Build: g++ -std=c++17 -pthread halo2.cpp -I../../include -L../../maestroex_b200/csrc -lmxadv
//        -L/usr/local/cuda/lib64 -lcudart   (done by __graft_entry__.build()).
#include <cuda_runtime.h>
#include <mxadv.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static const int NB[3] = {32, 28, 30};  // zones per brick
static const int NG = 3, NCOMP = 5;
static const int GRID[3] = {2, 1, 1};
static const int DOM[3] = {64, 28, 30};

static double field(int what, int c, int i, int j, int k) {
    // smooth + "hash" noise, periodic in x and y over the global domain
    const int ii = ((i % DOM[0]) + DOM[0]) % DOM[0], jj = ((j % DOM[1]) + DOM[1]) % DOM[1];
    unsigned h = 2166136261u;
    for (int v : {what, c, ii, jj, k}) h = (h ^ (unsigned)(v + 977)) * 16777619u;
    const double noise = (h >> 8) * (1.0 / 16777216.0);
    return 1.0 + 0.3 * std::sin(0.2 * ii + 0.1 * c) * std::cos(0.3 * jj) + 0.05 * k + 0.02 * noise +
           (what == 2 ? -1.2 : 0.0);
}

struct HostFab {
    std::vector<double> h;
    double* d = nullptr;
    MxFab f;
    long n[3];
    long at(int c, int i, int j, int k) const {
        return ((c * n[2] + (k - f.lo[2])) * n[1] + (j - f.lo[1])) * n[0] + (i - f.lo[0]);
    }
};

static HostFab make(const int lo[3], const int hi[3], int ncomp, int what) {
    HostFab F;
    for (int q = 0; q < 3; ++q) {
        F.f.lo[q] = lo[q];
        F.f.hi[q] = hi[q];
        F.n[q] = hi[q] - lo[q] + 1;
    }
    F.f.ncomp = ncomp;
    F.h.resize((size_t)ncomp * F.n[0] * F.n[1] * F.n[2]);
    for (int c = 0; c < ncomp; ++c)
        for (int k = lo[2]; k <= hi[2]; ++k)
            for (int j = lo[1]; j <= hi[1]; ++j)
                for (int i = lo[0]; i <= hi[0]; ++i) F.h[F.at(c, i, j, k)] = field(what, c, i, j, k);
    return F;
}

static std::atomic<int> g_fail{0};
static unsigned char g_id[MX_COMM_ID_BYTES];
static std::vector<double> g_iface[2];  // interface x faces of sedge as computed by each brick
#define CHECK(x)                                                                             \
    do {                                                                                     \
        if (!(x)) {                                                                          \
            std::fprintf(stderr, "rank %d: CHECK failed: %s (%s)\n", rank, #x, mx_last_error()); \
            g_fail++;                                                                        \
            return;                                                                          \
        }                                                                                    \
    } while (0)

static void worker(int rank, int transport) {
    CHECK(cudaSetDevice(rank) == cudaSuccess);
    MxComm* comm = nullptr;
    CHECK(mx_comm_create(&comm, g_id, rank, 2, GRID, transport) == MX_OK);
    MxGeom g;
    g.dim = 3;
    for (int q = 0; q < 3; ++q) {
        g.valid_lo[q] = q == 0 ? rank * NB[0] : 0;
        g.valid_hi[q] = g.valid_lo[q] + NB[q] - 1;
        g.dom_lo[q] = 0;
        g.dom_hi[q] = DOM[q] - 1;
        g.dx[q] = 1.0 / DOM[q];
    }
    auto grown = [&](int ng, int nodal_dir, int lo[3], int hi[3]) {
        for (int q = 0; q < 3; ++q) {
            lo[q] = g.valid_lo[q] - ng;
            hi[q] = g.valid_hi[q] + ng + (q == nodal_dir ? 1 : 0);
        }
    };
    int lo[3], hi[3];
    grown(NG, -1, lo, hi);
    HostFab state = make(lo, hi, NCOMP, 0);
    grown(1, -1, lo, hi);
    HostFab force = make(lo, hi, NCOMP, 1);
    HostFab umac[3], sedge[3];
    for (int d = 0; d < 3; ++d) {
        grown(1, d, lo, hi);
        umac[d] = make(lo, hi, 1, 2 + d);
        grown(0, d, lo, hi);
        sedge[d] = make(lo, hi, NCOMP, 9);
    }
    auto upload = [&](HostFab& F, int poison_ng) -> bool {
        std::vector<double> tmp = F.h;
        if (poison_ng > 0)  // ghost layers in x (both sides: periodic 2x1x1)
            for (int c = 0; c < F.f.ncomp; ++c)
                for (int k = F.f.lo[2]; k <= F.f.hi[2]; ++k)
                    for (int j = F.f.lo[1]; j <= F.f.hi[1]; ++j)
                        for (int q = 0; q < poison_ng; ++q) {
                            tmp[F.at(c, F.f.lo[0] + q, j, k)] = std::nan("");
                            tmp[F.at(c, F.f.hi[0] - q, j, k)] = std::nan("");
                        }
        if (cudaMalloc(&F.d, tmp.size() * 8) != cudaSuccess) return false;
        F.f.p = F.d;
        return cudaMemcpy(F.d, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    CHECK(upload(state, NG));
    CHECK(upload(force, 1));
    for (int d = 0; d < 3; ++d) {
        CHECK(upload(umac[d], 1));
        CHECK(upload(sedge[d], 0));
    }
    const int periodic[3] = {1, 1, 0};
    for (int rep = 0; rep < 2; ++rep) {  // twice: buffer / epoch re-use
        CHECK(mx_halo_exchange(comm, &g, &state.f, NG, 0, NCOMP, periodic, nullptr) == MX_OK);
        CHECK(mx_halo_exchange(comm, &g, &force.f, 1, 0, NCOMP, periodic, nullptr) == MX_OK);
        for (int d = 0; d < 3; ++d)
            CHECK(mx_halo_exchange(comm, &g, &umac[d].f, 1, 0, 1, periodic, nullptr) == MX_OK);
    }
    CHECK(cudaDeviceSynchronize() == cudaSuccess);
    auto same = [&](HostFab& F) -> bool {
        std::vector<double> got(F.h.size());
        if (cudaMemcpy(got.data(), F.d, got.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        return std::memcmp(got.data(), F.h.data(), got.size() * 8) == 0;
    };
    CHECK(same(state));
    CHECK(same(force));
    for (int d = 0; d < 3; ++d) CHECK(same(umac[d]));

    // fused edge-state sweep on the brick
    std::vector<int> bc(NCOMP * 6, 0);
    for (int c = 0; c < NCOMP; ++c) {
        bc[c * 6 + 2] = MX_BC_HOEXTRAP;  // lo z: slip wall
        bc[c * 6 + 5] = MX_BC_FOEXTRAP;  // hi z: outflow
    }
    MxEdgeParams par;
    std::memset(&par, 0, sizeof(par));
    par.dt = 0.2 * g.dx[0];
    par.rel_eps = 1e-8;
    par.ppm_type = 1;
    par.slope_order = 4;
    par.start_scomp = 0;
    par.start_bccomp = 0;
    par.num_comp = NCOMP;
    par.nbccomp = NCOMP;
    par.bc = bc.data();
    MxFab se[3] = {sedge[0].f, sedge[1].f, sedge[2].f}, um[3] = {umac[0].f, umac[1].f, umac[2].f};
    mx_reset_launch_count();
    CHECK(mx_make_edge_scal(&g, &state.f, se, um, &force.f, &par, nullptr) == MX_OK);
    CHECK(cudaDeviceSynchronize() == cudaSuccess);
    std::vector<double> ex(sedge[0].h.size());
    CHECK(cudaMemcpy(ex.data(), sedge[0].d, ex.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess);
    // x faces both bricks compute: global face 32 (my hi+1 on rank 0, my lo on rank 1) and, by
    // periodicity, face 0 of rank 0 == face 64 of rank 1
    const int faces[2][2] = {{NB[0], 0}, {NB[0], 2 * NB[0]}};
    for (int w = 0; w < 2; ++w)
        for (int c = 0; c < NCOMP; ++c)
            for (int k = 0; k < NB[2]; ++k)
                for (int j = 0; j < NB[1]; ++j)
                    g_iface[rank].push_back(ex[sedge[0].at(c, faces[rank][w], j, k)]);
    for (double v : g_iface[rank])
        if (!std::isfinite(v)) {
            std::fprintf(stderr, "rank %d: non-finite edge state\n", rank);
            g_fail++;
            break;
        }
    mx_comm_destroy(comm);
}

int main(int argc, char** argv) {
    const int transport = argc > 1 ? std::atoi(argv[1]) : MX_HALO_NCCL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) {
        std::printf("HALO2 SKIP: %d CUDA devices\n", ndev);
        return 77;
    }
    if (mx_comm_unique_id(g_id) != MX_OK) {
        std::fprintf(stderr, "mx_comm_unique_id: %s\n", mx_last_error());
        return 1;
    }
    std::thread t0(worker, 0, transport), t1(worker, 1, transport);
    t0.join();
    t1.join();
    if (g_fail) return 1;
    if (g_iface[0].size() != g_iface[1].size() ||
        std::memcmp(g_iface[0].data(), g_iface[1].data(), g_iface[0].size() * 8) != 0) {
        std::fprintf(stderr, "interface faces differ between the two bricks\n");
        return 1;
    }
    std::printf("HALO2 OK: transport %s, %zu interface face values identical on both bricks\n",
                transport ? "P2P" : "NCCL", g_iface[0].size());
    return 0;
}
