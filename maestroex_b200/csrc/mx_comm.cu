// mx_comm.cu — ghost-cell exchange between the bricks of one level, one brick per GPU: what
// FabArray::FillBoundary does for this path (external/amrex/Src/Base/AMReX_FabArrayCommI.H:8-150,
// called through FillPatch from Source/MaestroDensityAdvance.cpp:98-104 and
// MaestroEnthalpyAdvance.cpp:120-136 before MakeEdgeScal), behind the C ABI so that a C++ host
// (the overlay) has the same multi-GPU path bench.py uses.
//
// One rank (process or host thread) per GPU.  Bricks form a px*py*pz grid, x fastest (AMReX
// box order).  The exchange goes direction by direction (x, then y including the x ghosts,
// then z including both), which also fills the edge and corner ghosts the corner-coupled
// stencil reads.  Two transports:
//   MX_HALO_NCCL  device-packed slabs moved by ncclSend/ncclRecv inside one group per direction;
//   MX_HALO_P2P   the pack kernel of the sender stores straight into the receiver's slab buffer
//                 through a CUDA-IPC mapping (NVLink peer stores), completion is signalled by a
//                 device-side epoch flag the receiver's unpack kernel waits on; no host
//                 synchronisation, no NCCL call on the data path.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host application already
// loaded — e.g. PyTorch's — is reused, so two NCCL builds never meet in one process).
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "mx_common.h"
#include "mx_flux.h"

namespace {

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const Nccl* nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already loaded?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define SYM(name) *reinterpret_cast<void**>(&n.name) = dlsym(h, "nccl" #name)
        SYM(GetUniqueId);
        SYM(CommInitRank);
        SYM(CommDestroy);
        SYM(Send);
        SYM(Recv);
        SYM(AllGather);
        SYM(GroupStart);
        SYM(GroupEnd);
        SYM(GetErrorString);
#undef SYM
        if (n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.Send && n.Recv && n.AllGather &&
            n.GroupStart && n.GroupEnd && n.GetErrorString)
            n.h = h;
    });
    return n.h ? &n : nullptr;
}

#define NCCL_OK(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return mx_fail(MX_ERR_CUDA, "%s failed: %s", #call, nccl()->GetErrorString(r__));  \
    } while (0)
#define CU_OK(call)                                                                            \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return mx_fail(MX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));      \
    } while (0)

// ---- P2P transport: device-side flags -------------------------------------------------------
// Per (direction, side) a receiver owns a slab buffer and an epoch flag `ready`; the sender owns
// an `ack` flag the receiver bumps when it has unpacked.  Flags are 64-bit counters in device
// memory of their owner, written by the peer through its IPC mapping.
__global__ void k_halo_push(DFab fab, int dir, int start, int ng, int comp0, int ncomp,
                            double* __restrict__ peer_buf, volatile unsigned long long* my_ack,
                            unsigned long long* peer_ready, unsigned long long epoch,
                            unsigned int* done_blocks) {
    // the receiver must have unpacked what we pushed last time (it acknowledges epoch - 1)
    if (threadIdx.x == 0) {
        while (*my_ack + 1 < epoch) __nanosleep(64);
    }
    __syncthreads();
    int n[3] = {fab.n[0], fab.n[1], fab.n[2]};
    n[dir] = ng;
    const long per = (long)n[0] * n[1] * n[2];
    const long tot = per * ncomp;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot;
         t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t / per);
        long r = t % per;
        int idx[3];
        idx[0] = (int)(r % n[0]);
        r /= n[0];
        idx[1] = (int)(r % n[1]);
        idx[2] = (int)(r / n[1]);
        idx[dir] += start;
        peer_buf[t] = fab.comp(comp0 + c)[idx[0] + (long)idx[1] * fab.sy + (long)idx[2] * fab.sz];
    }
    // last block out publishes the epoch after every block's stores are visible system-wide
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        if (prev + 1 == gridDim.x) {
            *done_blocks = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(peer_ready) = epoch;
            __threadfence_system();
        }
    }
}

__global__ void k_halo_pull(DFab fab, int dir, int start, int ng, int comp0, int ncomp,
                            const double* __restrict__ my_buf, volatile unsigned long long* my_ready,
                            unsigned long long* peer_ack, unsigned long long epoch,
                            unsigned int* done_blocks) {
    if (threadIdx.x == 0) {
        while (*my_ready < epoch) __nanosleep(64);
        __threadfence_system();
    }
    __syncthreads();
    int n[3] = {fab.n[0], fab.n[1], fab.n[2]};
    n[dir] = ng;
    const long per = (long)n[0] * n[1] * n[2];
    const long tot = per * ncomp;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tot;
         t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t / per);
        long r = t % per;
        int idx[3];
        idx[0] = (int)(r % n[0]);
        r /= n[0];
        idx[1] = (int)(r % n[1]);
        idx[2] = (int)(r / n[1]);
        idx[dir] += start;
        fab.comp(comp0 + c)[idx[0] + (long)idx[1] * fab.sy + (long)idx[2] * fab.sz] =
            __ldcv(my_buf + t);  // written by the peer: never from a stale L1 line
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        if (prev + 1 == gridDim.x) {
            *done_blocks = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(peer_ack) = epoch;
            __threadfence_system();
        }
    }
}

struct P2PSlot {  // what one rank publishes about one device allocation
    cudaIpcMemHandle_t h;
    void* raw;               // the pointer itself: valid for ranks that live in the same process
    long long pid;
    int dev;
    unsigned long long cap;  // doubles (slab buffers)
};

}  // namespace

struct MxComm {
    int rank = 0, nranks = 1, grid[3] = {1, 1, 1}, transport = MX_HALO_NCCL, dev = 0;
    ncclComm_t comm = nullptr;
    double* sbuf[3][2] = {};
    double* rbuf[3][2] = {};
    size_t cap[3] = {};  // doubles per slab buffer of direction d
    long long bytes_sent = 0;
    // P2P: my flags / counters, and the peers' mapped buffers and flags
    unsigned long long* flags_all = nullptr;     // mine: [d][s][0] ready, [d][s][1] ack
    unsigned long long* flags[3][2] = {};        // = flags_all + 2 * (2 * d + s)
    std::vector<unsigned long long*> peer_flags_all;  // per rank: its flags_all as I can address it
    unsigned int* counters = nullptr;            // 12 block counters
    double* peer_rbuf[3][2] = {};                // receive buffer of the neighbour I send side s to
    unsigned long long* peer_flags_r[3][2] = {};  // flags of the neighbour that receives my side s
    unsigned long long* peer_flags_s[3][2] = {};  // flags of the neighbour whose data fills my ghost side s
    unsigned long long epoch[3] = {};
    std::vector<void*> mapped;

    int coord(int d) const {
        return d == 0 ? rank % grid[0] : d == 1 ? (rank / grid[0]) % grid[1] : rank / (grid[0] * grid[1]);
    }
    int neighbour(int d, int side, bool periodic) const {
        int c[3] = {coord(0), coord(1), coord(2)};
        c[d] += side == 0 ? -1 : 1;
        if (c[d] < 0 || c[d] >= grid[d]) {
            if (!periodic) return -1;
            c[d] = (c[d] + grid[d]) % grid[d];
        }
        return c[0] + grid[0] * (c[1] + grid[1] * c[2]);
    }
};

namespace {

int free_dir(MxComm* c, int d) {
    for (int s = 0; s < 2; ++s) {
        if (c->sbuf[d][s]) cudaFree(c->sbuf[d][s]);
        if (c->rbuf[d][s]) cudaFree(c->rbuf[d][s]);
        c->sbuf[d][s] = c->rbuf[d][s] = nullptr;
    }
    c->cap[d] = 0;
    return MX_OK;
}

// address a peer's allocation: the raw pointer when the peer lives in this process (threads,
// one GPU each: peer access is enabled), its CUDA-IPC mapping otherwise
int map_peer(MxComm* c, const P2PSlot& s, void** out) {
    if (s.pid == (long long)getpid()) {
        if (s.dev != c->dev) {
            cudaError_t e = cudaDeviceEnablePeerAccess(s.dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return mx_fail(MX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", s.dev,
                               cudaGetErrorString(e));
            cudaGetLastError();
        }
        *out = s.raw;
        return MX_OK;
    }
    CU_OK(cudaIpcOpenMemHandle(out, s.h, cudaIpcMemLazyEnablePeerAccess));
    c->mapped.push_back(*out);
    return MX_OK;
}

int describe(MxComm* c, void* p, size_t cap, P2PSlot* s) {
    memset(s, 0, sizeof(*s));
    CU_OK(cudaIpcGetMemHandle(&s->h, p));
    s->raw = p;
    s->pid = (long long)getpid();
    s->dev = c->dev;
    s->cap = cap;
    return MX_OK;
}

// every rank contributes `cnt` slots; all[r * cnt + q] = slot q of rank r
int allgather_slots(MxComm* c, const std::vector<P2PSlot>& mine, std::vector<P2PSlot>& all,
                    cudaStream_t stream) {
    const size_t cnt = mine.size();
    all.resize(cnt * c->nranks);
    P2PSlot *d_mine = nullptr, *d_all = nullptr;
    CU_OK(cudaMalloc(&d_mine, cnt * sizeof(P2PSlot)));
    CU_OK(cudaMalloc(&d_all, all.size() * sizeof(P2PSlot)));
    CU_OK(cudaMemcpy(d_mine, mine.data(), cnt * sizeof(P2PSlot), cudaMemcpyHostToDevice));
    NCCL_OK(nccl()->AllGather(d_mine, d_all, cnt * sizeof(P2PSlot), ncclChar, c->comm, stream));
    CU_OK(cudaStreamSynchronize(stream));
    CU_OK(cudaMemcpy(all.data(), d_all, all.size() * sizeof(P2PSlot), cudaMemcpyDeviceToHost));
    cudaFree(d_mine);
    cudaFree(d_all);
    return MX_OK;
}

// (re)allocate the slab buffers of direction d; P2P: exchange their handles with every rank
int ensure_buffers(MxComm* c, int d, size_t n, const int nbr[2], cudaStream_t stream) {
    if (c->cap[d] >= n) return MX_OK;
    if (c->transport == MX_HALO_P2P && c->cap[d] != 0)
        return mx_fail(MX_ERR_UNSUPPORTED,
                       "P2P halo exchange: slab size of direction %d grew after the first exchange "
                       "(%zu -> %zu doubles); exchange the largest fab first",
                       d, c->cap[d], n);
    CU_OK(cudaStreamSynchronize(stream));
    free_dir(c, d);
    for (int s = 0; s < 2; ++s) {
        if (c->transport == MX_HALO_NCCL) CU_OK(cudaMalloc(&c->sbuf[d][s], n * sizeof(double)));
        CU_OK(cudaMalloc(&c->rbuf[d][s], n * sizeof(double)));
    }
    c->cap[d] = n;
    if (c->transport != MX_HALO_P2P) return MX_OK;
    std::vector<P2PSlot> mine(2), all;
    for (int s = 0; s < 2; ++s) {
        int rc = describe(c, c->rbuf[d][s], n, &mine[s]);
        if (rc != MX_OK) return rc;
    }
    int rc = allgather_slots(c, mine, all, stream);
    if (rc != MX_OK) return rc;
    for (int s = 0; s < 2; ++s) {
        if (nbr[s] < 0) continue;
        // my side-s interior fills the neighbour's ghost side 1-s
        const P2PSlot& to = all[(size_t)2 * nbr[s] + (1 - s)];
        if (to.cap < n) return mx_fail(MX_ERR_INVALID, "P2P halo exchange: neighbour slab is smaller");
        void* p = nullptr;
        rc = map_peer(c, to, &p);
        if (rc != MX_OK) return rc;
        c->peer_rbuf[d][s] = static_cast<double*>(p);
        c->peer_flags_r[d][s] = c->peer_flags_all[nbr[s]] + 2 * (2 * d + (1 - s));
    }
    return MX_OK;
}

}  // namespace

extern "C" {

int mx_comm_unique_id(unsigned char id[MX_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == MX_COMM_ID_BYTES, "ncclUniqueId size");
    if (!nccl()) return mx_fail(MX_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    ncclUniqueId u;
    NCCL_OK(nccl()->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return MX_OK;
}

int mx_comm_create(MxComm** out, const unsigned char id[MX_COMM_ID_BYTES], int rank, int nranks,
                   const int grid[3], int transport) {
    if (!out || !id || !grid) return mx_fail(MX_ERR_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks || grid[0] * grid[1] * grid[2] != nranks)
        return mx_fail(MX_ERR_INVALID, "bad rank grid %dx%dx%d for %d ranks", grid[0], grid[1], grid[2], nranks);
    if (transport != MX_HALO_NCCL && transport != MX_HALO_P2P)
        return mx_fail(MX_ERR_INVALID, "unknown halo transport %d", transport);
    if (!nccl()) return mx_fail(MX_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    MxComm* c = new MxComm;
    c->rank = rank;
    c->nranks = nranks;
    c->transport = transport;
    for (int d = 0; d < 3; ++d) c->grid[d] = grid[d];
    cudaGetDevice(&c->dev);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = nccl()->CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return mx_fail(MX_ERR_CUDA, "ncclCommInitRank failed: %s", nccl()->GetErrorString(r));
    }
    if (transport == MX_HALO_P2P) {
        if (cudaMalloc(&c->flags_all, 12 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMemset(c->flags_all, 0, 12 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&c->counters, 12 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemset(c->counters, 0, 12 * sizeof(unsigned int)) != cudaSuccess) {
            mx_comm_destroy(c);
            return mx_fail(MX_ERR_CUDA, "P2P flag allocation failed");
        }
        for (int d = 0; d < 3; ++d)
            for (int s = 0; s < 2; ++s) c->flags[d][s] = c->flags_all + 2 * (2 * d + s);
        cudaDeviceSynchronize();
        // everybody's flag block, addressed once per peer
        std::vector<P2PSlot> mine(1), all;
        int rc = describe(c, c->flags_all, 12, &mine[0]);
        if (rc == MX_OK) rc = allgather_slots(c, mine, all, nullptr);
        c->peer_flags_all.assign(nranks, nullptr);
        for (int r = 0; rc == MX_OK && r < nranks; ++r) {
            if (r == rank) {
                c->peer_flags_all[r] = c->flags_all;
                continue;
            }
            bool is_nbr = false;
            for (int d = 0; d < 3; ++d)
                for (int s = 0; s < 2; ++s) is_nbr |= (c->neighbour(d, s, true) == r);
            if (!is_nbr) continue;
            void* p = nullptr;
            rc = map_peer(c, all[r], &p);
            c->peer_flags_all[r] = static_cast<unsigned long long*>(p);
        }
        if (rc != MX_OK) {
            mx_comm_destroy(c);
            return rc;
        }
    }
    *out = c;
    return MX_OK;
}

int mx_comm_destroy(MxComm* c) {
    if (!c) return MX_OK;
    cudaDeviceSynchronize();
    for (void* p : c->mapped) cudaIpcCloseMemHandle(p);
    for (int d = 0; d < 3; ++d) free_dir(c, d);
    if (c->flags_all) cudaFree(c->flags_all);
    if (c->counters) cudaFree(c->counters);
    if (c->comm && nccl()) nccl()->CommDestroy(c->comm);
    delete c;
    return MX_OK;
}

long long mx_comm_bytes_sent(const MxComm* c) { return c ? c->bytes_sent : 0; }

int mx_halo_exchange(MxComm* c, const MxGeom* geom, const MxFab* fab, int ng, int comp0, int ncomp,
                     const int periodic[3], void* stream_) {
    if (!c || !geom || !fab || !periodic) return mx_fail(MX_ERR_INVALID, "null argument");
    if (comp0 < 0 || ncomp < 1 || comp0 + ncomp > fab->ncomp)
        return mx_fail(MX_ERR_INVALID, "bad components");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const KGeom kg = make_kgeom(geom);
    const DFab f = make_dfab(fab, fab->p);
    c->bytes_sent = 0;
    for (int d = 0; d < geom->dim; ++d) {
        if (c->grid[d] == 1) continue;  // periodic image of myself or physical BC: caller's fill
        const int nghost = geom->valid_lo[d] - fab->lo[d];
        if (ng < 1 || ng > nghost)
            return mx_fail(MX_ERR_INVALID, "fab has %d ghost cells in direction %d, %d asked", nghost, d, ng);
        int n3[3] = {f.n[0], f.n[1], f.n[2]};
        n3[d] = ng;
        const size_t n = (size_t)n3[0] * n3[1] * n3[2] * ncomp;
        const int nbr[2] = {c->neighbour(d, 0, periodic[d] != 0), c->neighbour(d, 1, periodic[d] != 0)};
        // P2P slab buffers are registered with the peers once: size them for ALL components of
        // the fab, so that later exchanges of other component ranges of it fit
        const size_t n_cap = c->transport == MX_HALO_P2P ? (size_t)n3[0] * n3[1] * n3[2] * fab->ncomp : n;
        int rc = ensure_buffers(c, d, n_cap, nbr, stream);
        if (rc != MX_OK) return rc;
        if (c->transport == MX_HALO_NCCL) {
            for (int s = 0; s < 2; ++s)
                if (nbr[s] >= 0) halo_pack(kg, f, d, s, ng, comp0, ncomp, c->sbuf[d][s], false, stream);
            NCCL_OK(nccl()->GroupStart());
            for (int s = 0; s < 2; ++s)
                if (nbr[s] >= 0) {
                    NCCL_OK(nccl()->Send(c->sbuf[d][s], n, ncclDouble, nbr[s], c->comm, stream));
                    c->bytes_sent += (long long)n * 8;
                }
            // the k-th send to a peer meets that peer's k-th recv: the peer sends its LOW
            // interior first, which belongs in my HIGH ghosts
            for (int s = 1; s >= 0; --s)
                if (nbr[s] >= 0) NCCL_OK(nccl()->Recv(c->rbuf[d][s], n, ncclDouble, nbr[s], c->comm, stream));
            NCCL_OK(nccl()->GroupEnd());
            for (int s = 0; s < 2; ++s)
                if (nbr[s] >= 0) halo_pack(kg, f, d, s, ng, comp0, ncomp, c->rbuf[d][s], true, stream);
        } else {
            const unsigned long long ep = ++c->epoch[d];
            const int nodal = (f.n[d] - (geom->valid_hi[d] - geom->valid_lo[d] + 1)) & 1;
            const int start_i[2] = {nghost + nodal, f.n[d] - nghost - ng - nodal};  // interior slab
            const int start_g[2] = {nghost - ng, f.n[d] - nghost};      // ghost slab of side s
            const int blocks = (int)std::min<size_t>((n + 255) / 256, 4 * 148);
            for (int s = 0; s < 2; ++s) {
                if (nbr[s] < 0) continue;
                // my side-s interior -> ghost side 1-s of neighbour nbr[s]; its ready flag is
                // flags[d][1-s][0] over there, and it acknowledges into MY flags[d][s][1]
                k_halo_push<<<blocks, 256, 0, stream>>>(f, d, start_i[s], ng, comp0, ncomp,
                                                       c->peer_rbuf[d][s], c->flags[d][s] + 1,
                                                       c->peer_flags_r[d][s], ep, c->counters + 2 * d + s);
                mx_count_launch();
                c->bytes_sent += (long long)n * 8;
            }
            for (int s = 0; s < 2; ++s) {
                if (nbr[s] < 0) continue;
                // ghost side s is filled by neighbour nbr[s] (its side 1-s interior); I bump the
                // ack flag of ITS sending side: that is its flags[d][1-s][1], which I reach
                // through the same mapping I push my side-s data through (peer_flags_r[d][s])
                k_halo_pull<<<blocks, 256, 0, stream>>>(f, d, start_g[s], ng, comp0, ncomp, c->rbuf[d][s],
                                                       c->flags[d][s], c->peer_flags_r[d][s] + 1, ep,
                                                       c->counters + 6 + 2 * d + s);
                mx_count_launch();
            }
        }
        CU_OK(cudaGetLastError());
    }
    return MX_OK;
}

}  // extern "C"
