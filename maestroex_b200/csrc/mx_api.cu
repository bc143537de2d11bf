// mx_api.cu — the C ABI of libmxadv.so (include/mxadv.h): argument validation, pointer
// residency (device / managed / host), stream-ordered device workspace, host staging, and
// dispatch to the fused (mx_fused.cu) or staged (mx_staged.cu) kernels.  No CPU compute
// path exists: without a CUDA device every compute entry point returns MX_ERR_NO_DEVICE.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "mx_common.h"
#include "mx_flux.h"
#include "mx_w0.h"
#include "mx_update.h"
#include "mx_cart.h"
#include "mx_velpred.h"
#include "mx_fused.h"
#include "mx_staged.h"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<long long> g_tma_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_OK(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(MX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));     \
    } while (0)

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MX_ERR_NO_DEVICE,
                    "no CUDA device available; libmxadv has no CPU fallback (%s)",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    // per device (a single process may drive several GPUs): bit d set = pool of device d tuned
    static std::atomic<unsigned long long> pool_init{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(pool_init.load(std::memory_order_relaxed) & bit)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;  // keep freed workspace cached in the pool
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        pool_init.fetch_or(bit, std::memory_order_relaxed);
    }
    return MX_OK;
}

bool is_device_ptr(const void* p) {
    if (p == nullptr) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

#define MX_TRY_(expr)                   \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != MX_OK) return rc__; \
    } while (0)

// One API call's worth of device mirrors for host-resident arguments.
class Session {
   public:
    explicit Session(void* stream) : st_(static_cast<cudaStream_t>(stream)) {}
    ~Session() {
        for (void* p : owned_) cudaFreeAsync(p, st_);
    }
    cudaStream_t stream() const { return st_; }
    bool any_host() const { return any_host_; }

    cudaError_t alloc(void** p, size_t bytes) {
        cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 8, st_);
        if (e == cudaSuccess) owned_.push_back(*p);
        return e;
    }

    // Device view of a fab.  Host fabs get a full-size device mirror; components
    // [c0, c0+nc) are uploaded when `in`, and registered for download when `out`.
    int fab(const MxFab* f, DFab* out, int c0, int nc, bool in, bool outp) {
        if (f == nullptr || f->p == nullptr) return fail(MX_ERR_INVALID, "null fab");
        if (c0 < 0 || nc < 0 || c0 + nc > f->ncomp)
            return fail(MX_ERR_INVALID, "component range [%d,%d) outside fab with %d comps", c0,
                        c0 + nc, f->ncomp);
        for (int d = 0; d < 3; ++d)
            if (f->hi[d] < f->lo[d]) return fail(MX_ERR_INVALID, "empty fab box");
        if (is_device_ptr(f->p)) {
            *out = make_dfab(f, f->p);
            return MX_OK;
        }
        any_host_ = true;
        DFab d = make_dfab(f, nullptr);
        void* dev = nullptr;
        CUDA_OK(alloc(&dev, sizeof(double) * (size_t)d.sn * f->ncomp));
        d.p = static_cast<double*>(dev);
        if (pipe_) {  // copies are issued slab by slab in run_slabs()
            if (nc > 0 && (in || outp)) items_.push_back({f->p, d, c0, nc, in, outp, d.lo[2]});
            *out = d;
            return MX_OK;
        }
        if (in && nc > 0) {
            CUDA_OK(cudaMemcpyAsync(d.p + (size_t)c0 * d.sn, f->p + (size_t)c0 * d.sn,
                                    sizeof(double) * (size_t)d.sn * nc, cudaMemcpyHostToDevice,
                                    st_));
            h2d_ += sizeof(double) * (size_t)d.sn * nc;
        }
        if (outp && nc > 0) {
            backs_.push_back({f->p + (size_t)c0 * d.sn, d.p + (size_t)c0 * d.sn,
                              sizeof(double) * (size_t)d.sn * nc});
        }
        *out = d;
        return MX_OK;
    }

    // extra upload of another component range into an existing host mirror
    int upload(const MxFab* f, const DFab& d, int c0, int nc) {
        if (is_device_ptr(f->p) || nc <= 0) return MX_OK;
        if (pipe_) {
            items_.push_back({f->p, d, c0, nc, true, false, d.lo[2]});
            return MX_OK;
        }
        CUDA_OK(cudaMemcpyAsync(d.p + (size_t)c0 * d.sn, f->p + (size_t)c0 * d.sn,
                                sizeof(double) * (size_t)d.sn * nc, cudaMemcpyHostToDevice,
                                st_));
        return MX_OK;
    }
    int download_later(const MxFab* f, const DFab& d, int c0, int nc) {
        if (is_device_ptr(f->p) || nc <= 0) return MX_OK;
        if (pipe_) {
            // an earlier upload of the same components makes the item read-modify-write
            for (auto& it : items_) {
                if (it.h == f->p && it.c0 == c0 && it.nc == nc) {
                    it.out = true;
                    return MX_OK;
                }
            }
            items_.push_back({f->p, d, c0, nc, false, true, d.lo[2]});
            return MX_OK;
        }
        backs_.push_back({f->p + (size_t)c0 * d.sn, d.p + (size_t)c0 * d.sn,
                          sizeof(double) * (size_t)d.sn * nc});
        return MX_OK;
    }

    // download only the box [lo, hi] of components [c0, c0+nc) at finish(): for outputs that are
    // written on the valid zones only, so that the host fab's ghost zones neither have to be
    // uploaded first nor travel back (one strided 3-D copy per component)
    int download_box(const MxFab* f, const DFab& d, int c0, int nc, const int lo[3], const int hi[3]) {
        if (is_device_ptr(f->p) || nc <= 0) return MX_OK;
        Back3 b;
        b.host = f->p;
        b.g = d;
        b.c0 = c0;
        b.nc = nc;
        for (int q = 0; q < 3; ++q) {
            b.lo[q] = lo[q];
            b.hi[q] = hi[q];
        }
        backs3_.push_back(b);
        return MX_OK;
    }

    // 1-D base-state array at level `lev` -> device pointer + stride
    int base(const MxBase* b, int lev, const double** p, int* stride) {
        *p = nullptr;
        *stride = 1;
        if (b == nullptr || b->p == nullptr) return MX_OK;
        if (lev < 0 || lev >= b->nlev) return fail(MX_ERR_INVALID, "base-state level %d", lev);
        const size_t n = (size_t)b->nlev * b->len * b->nvar;
        const double* base = b->p;
        if (!is_device_ptr(b->p)) {
            void* dev = nullptr;
            CUDA_OK(alloc(&dev, sizeof(double) * n));
            CUDA_OK(cudaMemcpyAsync(dev, b->p, sizeof(double) * n, cudaMemcpyHostToDevice, st_));
            base = static_cast<const double*>(dev);
        }
        *p = base + (size_t)lev * b->len * b->nvar;
        *stride = b->nvar;
        return MX_OK;
    }

    // copy results back to host fabs and wait (host callers get synchronous semantics)
    int finish() {
        CUDA_OK(cudaGetLastError());
        if (pipe_) {  // deferred items that never went through run_slabs: whole components
            for (auto& it : items_)
                if (it.out)
                    backs_.push_back({it.h + (size_t)it.c0 * it.g.sn, it.g.p + (size_t)it.c0 * it.g.sn,
                                      sizeof(double) * (size_t)it.g.sn * it.nc});
            items_.clear();
            pipe_ = false;
        }
        for (auto& b : backs_) {
            CUDA_OK(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, st_));
        }
        for (auto& b : backs3_) {
            const size_t off = (size_t)(b.lo[0] - b.g.lo[0]) + (size_t)(b.lo[1] - b.g.lo[1]) * b.g.sy +
                               (size_t)(b.lo[2] - b.g.lo[2]) * b.g.sz;
            for (int c = b.c0; c < b.c0 + b.nc; ++c) {
                cudaMemcpy3DParms p3 = {};
                const size_t pitch = sizeof(double) * (size_t)b.g.n[0];
                p3.srcPtr = make_cudaPitchedPtr(b.g.p + (size_t)c * b.g.sn + off, pitch, b.g.n[0], b.g.n[1]);
                p3.dstPtr = make_cudaPitchedPtr(b.host + (size_t)c * b.g.sn + off, pitch, b.g.n[0], b.g.n[1]);
                p3.extent = make_cudaExtent(sizeof(double) * (size_t)(b.hi[0] - b.lo[0] + 1),
                                            (size_t)(b.hi[1] - b.lo[1] + 1), (size_t)(b.hi[2] - b.lo[2] + 1));
                p3.kind = cudaMemcpyDeviceToHost;
                CUDA_OK(cudaMemcpy3DAsync(&p3, st_));
            }
        }
        if (any_host_) CUDA_OK(cudaStreamSynchronize(st_));
        return MX_OK;
    }

    // ---- slab pipeline for host-resident fabs -------------------------------------------
    // A host caller's step is PCIe bound (928 B/zone-species of copies against ~60 B of HBM
    // traffic per zone-species of kernel work), and the link is full duplex: the valid z range
    // is cut into slabs and the uploads of slab s+1, the kernel of slab s and the downloads
    // of slab s-1 overlap (uploads and kernels in the caller's stream, downloads in a second one).  Results are unchanged: the kernels are box invariant
    // (tests/test_gpu_fullsize.py), every input plane is uploaded once, every written plane
    // downloaded once.
    void defer_copies() { pipe_ = true; }
    bool deferred() const { return pipe_; }

    // everything the deferred input items still owe, on the compute stream (leaves the
    // pipeline: the caller continues with the unpipelined composition and finish())
    int upload_rest() {
        for (auto& it : items_) {
            if (it.in) MX_TRY_(copy_planes(it, it.next, it.g.lo[2] + it.g.n[2] - 1, true, st_));
        }
        return MX_OK;
    }

    // launch(slab geometry) -> 0 ok, > 0 unsupported (only legal on the first slab), < 0 error
    int run_slabs(const KGeom& kg, int reach, const std::function<int(const KGeom&)>& launch,
                  int* rc_first) {
        *rc_first = 0;
        const int nz = kg.vhi[2] - kg.vlo[2] + 1;
        int thick = (nz + 7) / 8;  // measured on 128^3 boxes: 32-plane slabs beat 8 and 16
        if (thick < 32) thick = 32;
        if (const char* e = getenv("MXADV_SLAB")) thick = atoi(e) > 0 ? atoi(e) : nz;
        const int nslab = (nz + thick - 1) / thick;
        // one download stream per calling host thread (an OpenMP caller has a box in flight per
        // thread; sharing the stream would make every call wait for the others' copies)
        // ... and per device: a stream belongs to the device that was current at its creation
        constexpr int MAXDEV = 64;
        thread_local cudaStream_t s_outs[MAXDEV] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        cudaStream_t& s_out = s_outs[dev & (MAXDEV - 1)];
        if (!s_out) CUDA_OK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        // events live in a holder so that every early return below destroys them
        struct Events {
            std::vector<cudaEvent_t> v;
            ~Events() {
                for (auto e : v) cudaEventDestroy(e);
            }
        } evs, mks;
        for (int q = 0; q < 2 * nslab + 1; ++q) {
            cudaEvent_t e;
            CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            evs.v.push_back(e);
        }
        std::vector<cudaEvent_t>& ev = evs.v;
        // MXADV_TRACE=1: device timeline of the pipeline on stderr (debugging aid)
        const bool trace = getenv("MXADV_TRACE") != nullptr;
        std::vector<std::pair<const char*, cudaEvent_t> > marks;
        auto mark = [&](const char* what, cudaStream_t st) {
            if (!trace) return;
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            mks.v.push_back(e);
            cudaEventRecord(e, st);
            marks.push_back({what, e});
        };
        int rc = MX_OK;
        auto body = [&]() -> int {
            // the mirrors were allocated (and the base-state tables written) in st_ order
            mark("start", st_);
            CUDA_OK(cudaEventRecord(ev[2 * nslab], st_));
            CUDA_OK(cudaStreamWaitEvent(s_out, ev[2 * nslab], 0));
            for (int s = 0; s < nslab; ++s) {
                const bool first = s == 0, last = s == nslab - 1;
                KGeom g = kg;
                g.vlo[2] = kg.vlo[2] + s * thick;
                g.vhi[2] = last ? kg.vhi[2] : g.vlo[2] + thick - 1;
                g.skip_top_z = last ? 0 : 1;
                // inputs (read-only and read-modify-write alike: with skip_top_z every plane is
                // written by exactly one slab): each plane once, just ahead of its slab's kernel
                // and IN THE COMPUTE STREAM.  A separate upload stream was measured
                // (MXADV_TRACE=1) to buy nothing: the compute stream's wait on an event of the
                // upload stream only resolved once that stream's whole queue had drained.  The
                // downloads are the long pole (2.1x the upload bytes) and do overlap.
                for (auto& it : items_) {
                    if (!it.in) continue;
                    const int top = it.g.lo[2] + it.g.n[2] - 1;
                    const int hi = last ? top : std::min(top, g.vhi[2] + reach);
                    MX_TRY_(copy_planes(it, it.next, hi, true, st_));
                }
                mark("h2d done", st_);
                const int r = launch(g);
                if (r != 0) {
                    if (r > 0 && first) {
                        *rc_first = r;
                        return MX_OK;
                    }
                    return fail(MX_ERR_CUDA, "fused kernel launch failed in the slab pipeline");
                }
                mark("kernel done", st_);
                CUDA_OK(cudaEventRecord(ev[2 * s + 1], st_));
                CUDA_OK(cudaStreamWaitEvent(s_out, ev[2 * s + 1], 0));
                mark("d2h may start", s_out);
                for (auto& it : items_) {
                    if (!it.out) continue;
                    const int top = it.g.lo[2] + it.g.n[2] - 1;
                    const int lo = first ? std::max(it.g.lo[2], kg.vlo[2]) : g.vlo[2];
                    const int hi = last ? std::min(top, kg.vhi[2] + 1) : g.vhi[2];
                    MX_TRY_(copy_planes(it, lo, hi, false, s_out));
                }
                mark("d2h done", s_out);
            }
            CUDA_OK(cudaStreamSynchronize(s_out));
            CUDA_OK(cudaStreamSynchronize(st_));
            CUDA_OK(cudaGetLastError());
            return MX_OK;
        };
        rc = body();
        if (rc != MX_OK) {  // never free mirrors under copies still in flight
            cudaStreamSynchronize(s_out);
            cudaStreamSynchronize(st_);
        }
        if (trace && !marks.empty()) {
            for (auto& m : marks) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, marks[0].second, m.second);
                fprintf(stderr, "[mxadv pipeline] %8.3f ms  %s\n", ms, m.first);
            }
        }
        if (rc == MX_OK && *rc_first == 0) {
            items_.clear();
            pipe_ = false;
        }
        return rc;
    }

   private:
    struct Item {
        double* h;   // host fab
        DFab g;      // its device mirror
        int c0, nc;  // component range
        bool in, out;
        int next;    // first z plane not uploaded yet
    };
    // planes [k_lo, k_hi] of every component of the item, one contiguous copy per component
    int copy_planes(Item& it, int k_lo, int k_hi, bool to_device, cudaStream_t st) {
        const int top = it.g.lo[2] + it.g.n[2] - 1;
        k_lo = std::max(k_lo, it.g.lo[2]);
        k_hi = std::min(k_hi, top);
        if (k_hi < k_lo) return MX_OK;
        const size_t bytes = sizeof(double) * (size_t)it.g.sz * (k_hi - k_lo + 1);
        for (int c = it.c0; c < it.c0 + it.nc; ++c) {
            const size_t off = (size_t)c * it.g.sn + (size_t)(k_lo - it.g.lo[2]) * it.g.sz;
            if (to_device)
                CUDA_OK(cudaMemcpyAsync(it.g.p + off, it.h + off, bytes, cudaMemcpyHostToDevice, st));
            else
                CUDA_OK(cudaMemcpyAsync(it.h + off, it.g.p + off, bytes, cudaMemcpyDeviceToHost, st));
        }
        if (to_device) it.next = std::max(it.next, k_hi + 1);
        return MX_OK;
    }
    std::vector<Item> items_;
    bool pipe_ = false;
    struct Back {
        double* host;
        const double* dev;
        size_t bytes;
    };
    cudaStream_t st_;
    std::vector<void*> owned_;
    std::vector<Back> backs_;
    struct Back3 {
        double* host;
        DFab g;
        int c0, nc, lo[3], hi[3];
    };
    std::vector<Back3> backs3_;
    bool any_host_ = false;
    size_t h2d_ = 0;
};

#define MX_TRY(expr)              \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != MX_OK) return rc__; \
    } while (0)

int check_geom(const MxGeom* g) {
    if (g == nullptr) return fail(MX_ERR_INVALID, "null geometry");
    if (g->dim != 2 && g->dim != 3) return fail(MX_ERR_INVALID, "dim must be 2 or 3");
    for (int d = 0; d < g->dim; ++d) {
        if (g->valid_hi[d] < g->valid_lo[d]) return fail(MX_ERR_INVALID, "empty valid box");
        if (!(g->dx[d] > 0.0)) return fail(MX_ERR_INVALID, "dx must be positive");
    }
    return MX_OK;
}

// fab must cover grow(valid, ng) (+1 face in `nodal_dir`)
int check_cover(const MxGeom* g, const MxFab* f, int ng, int nodal_dir, const char* name) {
    for (int d = 0; d < g->dim; ++d) {
        const int hi = g->valid_hi[d] + (d == nodal_dir ? 1 : 0);
        if (f->lo[d] > g->valid_lo[d] - ng || f->hi[d] < hi + ng)
            return fail(MX_ERR_INVALID, "%s needs %d ghost cell(s) in direction %d", name, ng, d);
    }
    return MX_OK;
}

// true when the fab box is exactly the valid box made nodal in `nodal_dir` (no ghosts): every
// entry of the swept components is then overwritten and a host mirror needs no upload
bool exact_face_box(const MxGeom* g, const MxFab* f, int nodal_dir) {
    for (int d = 0; d < g->dim; ++d) {
        const int hi = g->valid_hi[d] + (d == nodal_dir ? 1 : 0);
        if (f->lo[d] != g->valid_lo[d] || f->hi[d] != hi) return false;
    }
    return true;
}

int check_edge_params(const MxGeom* g, const MxEdgeParams* p) {
    if (p == nullptr || p->bc == nullptr) return fail(MX_ERR_INVALID, "null edge parameters");
    if (p->ppm_type < 0 || p->ppm_type > 2)
        return fail(MX_ERR_UNSUPPORTED, "ppm_type %d (reference supports 0, 1, 2)", p->ppm_type);
    if (p->ppm_type == 0 && p->slope_order != 0 && p->slope_order != 2 && p->slope_order != 4)
        return fail(MX_ERR_UNSUPPORTED, "slope_order %d (reference supports 0, 2, 4)",
                    p->slope_order);
    if (p->num_comp < 0 || p->start_scomp < 0) return fail(MX_ERR_INVALID, "bad component range");
    if (p->start_bccomp < 0 || p->start_bccomp + p->num_comp > p->nbccomp)
        return fail(MX_ERR_INVALID, "bc table too short");
    (void)g;
    return MX_OK;
}

KEdge make_kedge(const MxEdgeParams* p) {
    KEdge e;
    memset(&e, 0, sizeof(e));
    e.dt = p->dt;
    e.eps = p->rel_eps;
    e.ppm_type = p->ppm_type;
    e.slope_order = p->slope_order;
    e.trace = p->ppm_trace_forces;
    e.is_vel = p->is_vel;
    e.is_cons = p->is_conservative;
    return e;
}

struct EdgeArgs {
    DFab state, force, sedge[3], umac[3];
};

int stage_edge_args(Session& S, const MxGeom* g, const MxFab* state, const MxFab* sedge,
                    const MxFab* umac, const MxFab* force, const MxEdgeParams* p, EdgeArgs* a) {
    const int ng_s = p->ppm_type == 2 ? 4 : 3;
    MX_TRY(check_cover(g, state, ng_s, -1, "state"));
    MX_TRY(check_cover(g, force, (p->ppm_trace_forces && p->ppm_type != 0) ? ng_s : 1, -1,
                       "force"));
    MX_TRY(S.fab(state, &a->state, p->start_scomp, p->num_comp, true, false));
    MX_TRY(S.fab(force, &a->force, p->start_scomp, p->num_comp, true, false));
    for (int d = 0; d < g->dim; ++d) {
        MX_TRY(check_cover(g, &umac[d], 1, d, "umac"));
        MX_TRY(check_cover(g, &sedge[d], 0, d, "sedge"));
        if (umac[d].ncomp < 1) return fail(MX_ERR_INVALID, "umac has no components");
        MX_TRY(S.fab(&umac[d], &a->umac[d], 0, 1, true, false));
        MX_TRY(S.fab(&sedge[d], &a->sedge[d], p->start_scomp, p->num_comp,
                     !exact_face_box(g, &sedge[d], d), true));
    }
    if (g->dim == 2) {
        a->umac[2] = a->umac[0];
        a->sedge[2] = a->sedge[0];
    }
    return MX_OK;
}

int run_edge(Session& S, const KGeom& kg, const MxEdgeParams* p, const EdgeArgs& a) {
    KEdge e = make_kedge(p);
    if (p->num_comp == 0) return MX_OK;
    if (fused_edge_supported(kg, e)) {
        // > 0: a BC code the fused kernel does not carry (ext_dir): use the staged kernels
        const int rc = fused_edge_scal(kg, e, p, a.state, a.sedge, a.umac, a.force, S.stream());
        if (rc == 0) return MX_OK;
        if (rc < 0) return fail(MX_ERR_CUDA, "fused edge kernel launch failed");
    }
    void* scratch = nullptr;
    CUDA_OK(S.alloc(&scratch, sizeof(double) * staged_scratch_doubles(kg)));
    staged_make_edge_scal(kg, e, p, a.state, a.sedge, a.umac, a.force,
                          static_cast<double*>(scratch), S.stream());
    return MX_OK;
}

int fill_kflux(Session& S, const MxFluxParams* p, const MxBase* rho0_old,
               const MxBase* rho0_edge_old, const MxBase* rho0_new, const MxBase* rho0_edge_new,
               const MxBase* rho0_pred, const MxBase* w0, const MxBase* rhoh0_old,
               const MxBase* rhoh0_edge_old, const MxBase* rhoh0_new,
               const MxBase* rhoh0_edge_new, KFlux* f) {
    memset(f, 0, sizeof(*f));
    f->spt = p->species_pred_type;
    f->ept = p->enthalpy_pred_type;
    f->spherical = p->spherical;
    f->exact = p->use_exact_base_state;
    f->evolve = p->evolve_base_state;
    f->rho = p->rho_comp;
    f->rhoh = p->rhoh_comp;
    f->first_spec = p->first_spec;
    f->num_spec = p->num_spec;
    f->start_comp = p->start_comp;
    f->num_comp = p->num_comp;
    // spherical: the 1-D arrays have ONE radial level whatever the AMR level and the flux forms
    // read the base state from the face arrays (r0mac / h0mac), never from these
    // (Source/MaestroMakeFlux.cpp:311-407, :788-1021): not staged at all
    if (p->spherical) return MX_OK;
    MX_TRY(S.base(rho0_old, p->lev, &f->rho0_old, &f->st_rho0_old));
    MX_TRY(S.base(rho0_new, p->lev, &f->rho0_new, &f->st_rho0_new));
    MX_TRY(S.base(rho0_edge_old, p->lev, &f->rho0_edge_old, &f->st_rho0_edge_old));
    MX_TRY(S.base(rho0_edge_new, p->lev, &f->rho0_edge_new, &f->st_rho0_edge_new));
    MX_TRY(S.base(rho0_pred, p->lev, &f->rho0_pred_edge, &f->st_rho0_pred_edge));
    MX_TRY(S.base(w0, p->lev, &f->w0, &f->st_w0));
    MX_TRY(S.base(rhoh0_old, p->lev, &f->rhoh0_old, &f->st_rhoh0_old));
    MX_TRY(S.base(rhoh0_new, p->lev, &f->rhoh0_new, &f->st_rhoh0_new));
    MX_TRY(S.base(rhoh0_edge_old, p->lev, &f->rhoh0_edge_old, &f->st_rhoh0_edge_old));
    MX_TRY(S.base(rhoh0_edge_new, p->lev, &f->rhoh0_edge_new, &f->st_rhoh0_edge_new));
    return MX_OK;
}

int stage_mac(Session& S, const MxGeom* g, const MxFab* src, DFab* dst, const char* name) {
    if (src == nullptr) return fail(MX_ERR_INVALID, "%s is required when spherical", name);
    for (int d = 0; d < 3; ++d) {
        MX_TRY(check_cover(g, &src[d], 0, d, name));
        MX_TRY(S.fab(&src[d], &dst[d], 0, 1, true, false));
    }
    return MX_OK;
}

int need_base(const MxBase* b, const char* name) {
    if (b == nullptr || b->p == nullptr) return fail(MX_ERR_INVALID, "%s is required", name);
    return MX_OK;
}

}  // namespace

int mx_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
void mx_count_launch(int n) { g_launches += n; }
void mx_count_tma_launch() { g_tma_launches += 1; }

extern "C" {

int mx_abi_version(void) { return MXADV_ABI_VERSION; }
const char* mx_last_error(void) { return g_err.c_str(); }

int mx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

long long mx_launch_count(void) { return g_launches.load(); }
void mx_reset_launch_count(void) {
    g_launches = 0;
    g_tma_launches = 0;
}
long long mx_tma_launch_count(void) { return g_tma_launches.load(); }

int mx_free_workspace(void) {
    MX_TRY(require_device());
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemPoolTrimTo(pool, 0));
    return MX_OK;
}

int mx_make_edge_scal(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                      const MxFab* umac, const MxFab* force, const MxEdgeParams* par,
                      void* stream) {
    MX_TRY(check_geom(geom));
    if (!state || !sedge || !umac || !force) return fail(MX_ERR_INVALID, "null fab argument");
    MX_TRY(check_edge_params(geom, par));
    MX_TRY(require_device());
    Session S(stream);
    EdgeArgs a;
    MX_TRY(stage_edge_args(S, geom, state, sedge, umac, force, par, &a));
    const KGeom kg = make_kgeom(geom);
    MX_TRY(run_edge(S, kg, par, a));
    return S.finish();
}

int mx_make_rhoX_flux(const MxGeom* geom, const MxFab* sflux, const MxFab* etarhoflux,
                      const MxFab* sedge, const MxFab* umac, const MxBase* rho0_old,
                      const MxBase* rho0_edge_old, const MxFab* r0mac_old,
                      const MxBase* rho0_new, const MxBase* rho0_edge_new,
                      const MxFab* r0mac_new, const MxBase* rho0_predicted_edge,
                      const MxBase* w0, const MxFluxParams* par, void* stream) {
    MX_TRY(check_geom(geom));
    if (!sflux || !sedge || !umac || !par) return fail(MX_ERR_INVALID, "null argument");
    if (par->species_pred_type < 1 || par->species_pred_type > 3)
        return fail(MX_ERR_UNSUPPORTED, "species_pred_type %d", par->species_pred_type);
    MX_TRY(require_device());
    const bool sph = geom->dim == 3 && par->spherical;
    const bool do_eta = !sph && par->evolve_base_state && !par->use_exact_base_state;
    if (!sph && par->species_pred_type == MX_PRED_RHOPRIME_AND_X) {
        MX_TRY(need_base(rho0_old, "rho0_old"));
        MX_TRY(need_base(rho0_new, "rho0_new"));
        MX_TRY(need_base(rho0_edge_old, "rho0_edge_old"));
        MX_TRY(need_base(rho0_edge_new, "rho0_edge_new"));
    }
    if (do_eta) {
        if (!etarhoflux) return fail(MX_ERR_INVALID, "etarhoflux is required");
        MX_TRY(need_base(w0, "w0"));
        MX_TRY(need_base(rho0_predicted_edge, "rho0_predicted_edge"));
    }
    Session S(stream);
    KFlux f;
    MX_TRY(fill_kflux(S, par, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new,
                      rho0_predicted_edge, w0, nullptr, nullptr, nullptr, nullptr, &f));
    const KGeom kg = make_kgeom(geom);
    DFab sf[3], se[3], um[3], eta;
    KMac mac;
    memset(&mac, 0, sizeof(mac));
    memset(&eta, 0, sizeof(eta));
    for (int d = 0; d < geom->dim; ++d) {
        MX_TRY(check_cover(geom, &sflux[d], 0, d, "sflux"));
        MX_TRY(check_cover(geom, &sedge[d], 0, d, "sedge"));
        MX_TRY(check_cover(geom, &umac[d], 0, d, "umac"));
        // species components in/out; the Rho component of sedge is an input, of sflux an output
        const bool keep = !exact_face_box(geom, &sflux[d], d);
        MX_TRY(S.fab(&sflux[d], &sf[d], par->start_comp, par->num_comp, keep, true));
        if (keep) MX_TRY(S.upload(&sflux[d], sf[d], par->rho_comp, 1));
        MX_TRY(S.download_later(&sflux[d], sf[d], par->rho_comp, 1));
        MX_TRY(S.fab(&sedge[d], &se[d], par->start_comp, par->num_comp, true, false));
        MX_TRY(S.upload(&sedge[d], se[d], par->rho_comp, 1));
        MX_TRY(S.fab(&umac[d], &um[d], 0, 1, true, false));
    }
    if (do_eta) {
        MX_TRY(check_cover(geom, etarhoflux, 0, geom->dim - 1, "etarhoflux"));
        MX_TRY(S.fab(etarhoflux, &eta, 0, 1, true, true));
    }
    if (sph) {
        MX_TRY(stage_mac(S, geom, r0mac_old, mac.r0_old, "r0mac_old"));
        MX_TRY(stage_mac(S, geom, r0mac_new, mac.r0_new, "r0mac_new"));
    }
    flux_rhoX(kg, f, sf, eta, se, um, mac, S.stream());
    return S.finish();
}

int mx_make_rhoh_flux(const MxGeom* geom, const MxFab* sflux, const MxFab* sedge,
                      const MxFab* umac, const MxBase* rho0_old, const MxBase* rho0_edge_old,
                      const MxFab* r0mac_old, const MxBase* rho0_new,
                      const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                      const MxBase* rhoh0_old, const MxBase* rhoh0_edge_old,
                      const MxFab* rh0mac_old, const MxBase* rhoh0_new,
                      const MxBase* rhoh0_edge_new, const MxFab* rh0mac_new,
                      const MxFab* h0mac_old, const MxFab* h0mac_new, const MxFluxParams* par,
                      void* stream) {
    MX_TRY(check_geom(geom));
    if (!sflux || !sedge || !umac || !par) return fail(MX_ERR_INVALID, "null argument");
    if (par->enthalpy_pred_type < 0 || par->enthalpy_pred_type > 6)
        return fail(MX_ERR_UNSUPPORTED, "enthalpy_pred_type %d", par->enthalpy_pred_type);
    MX_TRY(require_device());
    const bool sph = geom->dim == 3 && par->spherical;
    const int ept = par->enthalpy_pred_type;
    const bool have_h =
        ept == MX_PRED_H || ept == MX_PRED_T_THEN_H || ept == MX_PRED_TPRIME_THEN_H;
    if (!sph) {
        if (have_h && par->species_pred_type == MX_PRED_RHOPRIME_AND_X) {
            MX_TRY(need_base(rho0_old, "rho0_old"));
            MX_TRY(need_base(rho0_new, "rho0_new"));
            MX_TRY(need_base(rho0_edge_old, "rho0_edge_old"));
            MX_TRY(need_base(rho0_edge_new, "rho0_edge_new"));
        }
        if (ept == MX_PRED_RHOHPRIME || ept == MX_PRED_T_THEN_RHOHPRIME) {
            MX_TRY(need_base(rhoh0_old, "rhoh0_old"));
            MX_TRY(need_base(rhoh0_new, "rhoh0_new"));
            MX_TRY(need_base(rhoh0_edge_old, "rhoh0_edge_old"));
            MX_TRY(need_base(rhoh0_edge_new, "rhoh0_edge_new"));
        }
    }
    Session S(stream);
    KFlux f;
    MX_TRY(fill_kflux(S, par, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new, nullptr, nullptr,
                      rhoh0_old, rhoh0_edge_old, rhoh0_new, rhoh0_edge_new, &f));
    const KGeom kg = make_kgeom(geom);
    DFab sf[3], se[3], um[3];
    KMac mac;
    memset(&mac, 0, sizeof(mac));
    for (int d = 0; d < geom->dim; ++d) {
        MX_TRY(check_cover(geom, &sflux[d], 0, d, "sflux"));
        MX_TRY(check_cover(geom, &sedge[d], 0, d, "sedge"));
        MX_TRY(check_cover(geom, &umac[d], 0, d, "umac"));
        // the reference leaves sflux(RhoH) untouched for some option combinations: in + out
        MX_TRY(S.fab(&sflux[d], &sf[d], par->rhoh_comp, 1, true, true));
        MX_TRY(S.fab(&sedge[d], &se[d], par->rhoh_comp, 1, true, false));
        MX_TRY(S.upload(&sedge[d], se[d], par->rho_comp, 1));
        MX_TRY(S.fab(&umac[d], &um[d], 0, 1, true, false));
    }
    if (sph) {
        if (par->use_exact_base_state) {
            MX_TRY(stage_mac(S, geom, rh0mac_old, mac.rh0_old, "rh0mac_old"));
            MX_TRY(stage_mac(S, geom, rh0mac_new, mac.rh0_new, "rh0mac_new"));
        } else {
            MX_TRY(stage_mac(S, geom, r0mac_old, mac.r0_old, "r0mac_old"));
            MX_TRY(stage_mac(S, geom, r0mac_new, mac.r0_new, "r0mac_new"));
            MX_TRY(stage_mac(S, geom, h0mac_old, mac.h0_old, "h0mac_old"));
            MX_TRY(stage_mac(S, geom, h0mac_new, mac.h0_new, "h0mac_new"));
        }
    }
    flux_rhoh(kg, f, sf, se, um, mac, S.stream());
    return S.finish();
}

int mx_density_edge_flux(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                         const MxFab* sflux, const MxFab* etarhoflux, const MxFab* umac,
                         const MxFab* force, const MxEdgeParams* epar, int rho_bccomp,
                         const MxBase* rho0_old, const MxBase* rho0_edge_old,
                         const MxFab* r0mac_old, const MxBase* rho0_new,
                         const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                         const MxBase* rho0_predicted_edge, const MxBase* w0,
                         const MxFluxParams* fpar, void* stream) {
    MX_TRY(check_geom(geom));
    if (!state || !sedge || !sflux || !umac || !force || !fpar)
        return fail(MX_ERR_INVALID, "null argument");
    MX_TRY(check_edge_params(geom, epar));
    if (geom->dim != 3) return fail(MX_ERR_UNSUPPORTED, "mx_density_edge_flux is 3-D only");
    if (epar->is_vel || epar->is_conservative)
        return fail(MX_ERR_UNSUPPORTED,
                    "mx_density_edge_flux: is_vel / is_conservative sweeps use mx_make_edge_scal");
    if (fpar->species_pred_type != MX_PRED_RHOPRIME_AND_X &&
        fpar->species_pred_type != MX_PRED_RHO_AND_X)
        return fail(MX_ERR_UNSUPPORTED,
                    "mx_density_edge_flux: species_pred_type %d predicts (rho X) conservatively; "
                    "use mx_make_edge_scal + mx_make_rhoX_flux",
                    fpar->species_pred_type);
    if (epar->start_scomp != fpar->start_comp || epar->num_comp != fpar->num_comp)
        return fail(MX_ERR_INVALID, "edge and flux component ranges differ");
    if (rho_bccomp < 0 || rho_bccomp >= epar->nbccomp)
        return fail(MX_ERR_INVALID, "rho_bccomp outside bc table");
    MX_TRY(require_device());
    const bool sph = fpar->spherical != 0;
    const bool do_eta = !sph && fpar->evolve_base_state && !fpar->use_exact_base_state;
    if (do_eta && !etarhoflux) return fail(MX_ERR_INVALID, "etarhoflux is required");

    Session S(stream);
    KFlux f;
    MX_TRY(fill_kflux(S, fpar, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new,
                      rho0_predicted_edge, w0, nullptr, nullptr, nullptr, nullptr, &f));
    const KGeom kg = make_kgeom(geom);
    const int ng_s = epar->ppm_type == 2 ? 4 : 3;
    const KEdge e = make_kedge(epar);
    // host fabs: copy and compute slab by slab (Session::run_slabs)
    if (!sph && fused_density_supported(kg, e, f) && !is_device_ptr(state->p) &&
        kg.vhi[2] - kg.vlo[2] + 1 >= 32 && !getenv("MXADV_NO_PIPELINE"))
        S.defer_copies();
    MX_TRY(check_cover(geom, state, ng_s, -1, "state"));
    MX_TRY(check_cover(geom, force, 1, -1, "force"));
    const int rho = fpar->rho_comp;
    EdgeArgs a;
    MX_TRY(S.fab(state, &a.state, epar->start_scomp, epar->num_comp, true, false));
    MX_TRY(S.upload(state, a.state, rho, 1));
    MX_TRY(S.fab(force, &a.force, epar->start_scomp, epar->num_comp, true, false));
    MX_TRY(S.upload(force, a.force, rho, 1));
    DFab sf[3], eta;
    memset(&eta, 0, sizeof(eta));
    KMac mac;
    memset(&mac, 0, sizeof(mac));
    for (int d = 0; d < 3; ++d) {
        MX_TRY(check_cover(geom, &umac[d], 1, d, "umac"));
        MX_TRY(check_cover(geom, &sedge[d], 0, d, "sedge"));
        MX_TRY(check_cover(geom, &sflux[d], 0, d, "sflux"));
        MX_TRY(S.fab(&umac[d], &a.umac[d], 0, 1, true, false));
        const bool keep_e = !exact_face_box(geom, &sedge[d], d);
        const bool keep_f = !exact_face_box(geom, &sflux[d], d);
        MX_TRY(S.fab(&sedge[d], &a.sedge[d], epar->start_scomp, epar->num_comp, keep_e, true));
        if (keep_e) MX_TRY(S.upload(&sedge[d], a.sedge[d], rho, 1));
        MX_TRY(S.download_later(&sedge[d], a.sedge[d], rho, 1));
        MX_TRY(S.fab(&sflux[d], &sf[d], epar->start_scomp, epar->num_comp, keep_f, true));
        if (keep_f) MX_TRY(S.upload(&sflux[d], sf[d], rho, 1));
        MX_TRY(S.download_later(&sflux[d], sf[d], rho, 1));
    }
    if (do_eta) {
        MX_TRY(check_cover(geom, etarhoflux, 0, 2, "etarhoflux"));
        MX_TRY(S.fab(etarhoflux, &eta, 0, 1, true, true));
    }
    if (sph) {
        MX_TRY(stage_mac(S, geom, r0mac_old, mac.r0_old, "r0mac_old"));
        MX_TRY(stage_mac(S, geom, r0mac_new, mac.r0_new, "r0mac_new"));
    }
    if (S.deferred()) {
        int rc1 = 0;
        MX_TRY(S.run_slabs(
            kg, ng_s + 1,
            [&](const KGeom& g) {
                return fused_density(g, e, epar, rho_bccomp, f, a.state, a.sedge, sf, eta, a.umac,
                                     a.force, mac, nullptr, S.stream());
            },
            &rc1));
        if (rc1 == 0) return S.finish();
        MX_TRY(S.upload_rest());  // component set the fused kernel does not take: compose below
    } else if (fused_density_supported(kg, e, f)) {
        // spherical: the time-centred rho0 on faces is species independent; form it once
        // (3 x 8 B/zone of workspace) instead of twice per face and species inside the kernel
        double* r0e[3] = {nullptr, nullptr, nullptr};
        if (sph) {
            for (int d = 0; d < 3; ++d) {
                void* p = nullptr;
                CUDA_OK(S.alloc(&p, sizeof(double) * (size_t)a.sedge[d].sn));
                r0e[d] = static_cast<double*>(p);
                mac_average(kg, d, mac.r0_old[d], mac.r0_new[d], a.sedge[d], r0e[d], S.stream());
            }
        }
        const int rc = fused_density(kg, e, epar, rho_bccomp, f, a.state, a.sedge, sf, eta, a.umac,
                                     a.force, mac, sph ? r0e : nullptr, S.stream());
        if (rc == 0) return S.finish();
        if (rc < 0) return fail(MX_ERR_CUDA, "fused density kernel launch failed");
    }
    // composition of the three reference calls (MaestroDensityAdvance.cpp:115,128,182)
    MX_TRY(run_edge(S, kg, epar, a));
    MxEdgeParams rpar = *epar;
    rpar.start_scomp = rho;
    rpar.start_bccomp = rho_bccomp;
    rpar.num_comp = 1;
    MX_TRY(run_edge(S, kg, &rpar, a));
    flux_rhoX(kg, f, sf, eta, a.sedge, a.umac, mac, S.stream());
    return S.finish();
}

static int enthalpy_edge_flux_impl(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                                   const MxFab* sflux, const MxFab* umac, const MxFab* force,
                                   const MxEdgeParams* epar, const MxBase* rho0_old,
                                   const MxBase* rho0_edge_old, const MxBase* rho0_new,
                                   const MxBase* rho0_edge_new, const MxBase* rhoh0_old,
                                   const MxBase* rhoh0_edge_old, const MxBase* rhoh0_new,
                                   const MxBase* rhoh0_edge_new, const MxFab* r0mac_old,
                                   const MxFab* r0mac_new, const MxFab* rh0mac_old,
                                   const MxFab* rh0mac_new, const MxFab* h0mac_old,
                                   const MxFab* h0mac_new, const MxFluxParams* fpar,
                                   void* stream) {
    MX_TRY(check_geom(geom));
    if (!state || !sedge || !sflux || !umac || !force || !fpar)
        return fail(MX_ERR_INVALID, "null argument");
    MX_TRY(check_edge_params(geom, epar));
    if (geom->dim != 3) return fail(MX_ERR_UNSUPPORTED, "the fused enthalpy pass is 3-D only");
    const bool sph = fpar->spherical != 0;
    if (sph && (fpar->use_exact_base_state ? (!rh0mac_old || !rh0mac_new)
                                           : (!r0mac_old || !r0mac_new || !h0mac_old || !h0mac_new)))
        return fail(MX_ERR_INVALID, "spherical enthalpy pass: base-state face arrays are required");
    if (epar->is_vel) return fail(MX_ERR_UNSUPPORTED, "mx_enthalpy_edge_flux: is_vel");
    if (epar->num_comp != 1 || epar->start_scomp != fpar->rhoh_comp)
        return fail(MX_ERR_INVALID, "mx_enthalpy_edge_flux sweeps the single component RhoH");
    const int ept = fpar->enthalpy_pred_type;
    if (ept == MX_PRED_T_THEN_RHOHPRIME || ept == MX_PRED_T_THEN_H ||
        ept == MX_PRED_TPRIME_THEN_H)
        return fail(MX_ERR_UNSUPPORTED,
                    "enthalpy_pred_type %d needs the EOS between edge prediction and flux "
                    "(HfromRhoTedge); call mx_make_edge_scal and mx_make_rhoh_flux separately",
                    ept);
    MX_TRY(require_device());
    Session S(stream);
    KFlux f;
    MX_TRY(fill_kflux(S, fpar, rho0_old, rho0_edge_old, rho0_new, rho0_edge_new, nullptr, nullptr,
                      rhoh0_old, rhoh0_edge_old, rhoh0_new, rhoh0_edge_new, &f));
    const KGeom kg = make_kgeom(geom);
    const KEdge e = make_kedge(epar);
    if (!sph && fused_enthalpy_supported(kg, e, f) && !is_device_ptr(state->p) &&
        kg.vhi[2] - kg.vlo[2] + 1 >= 32 && !getenv("MXADV_NO_PIPELINE"))
        S.defer_copies();
    EdgeArgs a;
    MX_TRY(stage_edge_args(S, geom, state, sedge, umac, force, epar, &a));
    DFab sf[3];
    KMac mac;
    memset(&mac, 0, sizeof(mac));
    for (int d = 0; d < 3; ++d) {
        MX_TRY(check_cover(geom, &sflux[d], 0, d, "sflux"));
        MX_TRY(S.fab(&sflux[d], &sf[d], fpar->rhoh_comp, 1, !exact_face_box(geom, &sflux[d], d),
                     true));
        MX_TRY(S.upload(&sedge[d], a.sedge[d], fpar->rho_comp, 1));
    }
    if (sph) {
        if (fpar->use_exact_base_state) {
            MX_TRY(stage_mac(S, geom, rh0mac_old, mac.rh0_old, "rh0mac_old"));
            MX_TRY(stage_mac(S, geom, rh0mac_new, mac.rh0_new, "rh0mac_new"));
        } else {
            MX_TRY(stage_mac(S, geom, r0mac_old, mac.r0_old, "r0mac_old"));
            MX_TRY(stage_mac(S, geom, r0mac_new, mac.r0_new, "r0mac_new"));
            MX_TRY(stage_mac(S, geom, h0mac_old, mac.h0_old, "h0mac_old"));
            MX_TRY(stage_mac(S, geom, h0mac_new, mac.h0_new, "h0mac_new"));
        }
    }
    if (S.deferred()) {
        int rc1 = 0;
        MX_TRY(S.run_slabs(
            kg, (epar->ppm_type == 2 ? 4 : 3) + 1,
            [&](const KGeom& g) {
                return fused_enthalpy(g, e, epar, f, a.state, a.sedge, sf, a.umac, a.force, nullptr,
                                      nullptr, S.stream());
            },
            &rc1));
        if (rc1 == 0) return S.finish();
        MX_TRY(S.upload_rest());
    } else if (fused_enthalpy_supported(kg, e, f)) {
        // spherical: the time-centred rho0 and h0 on faces, formed once in sedge's layout
        double *r0e[3] = {nullptr, nullptr, nullptr}, *h0e[3] = {nullptr, nullptr, nullptr};
        if (sph) {
            for (int d = 0; d < 3; ++d) {
                void* p = nullptr;
                CUDA_OK(S.alloc(&p, sizeof(double) * (size_t)a.sedge[d].sn));
                r0e[d] = static_cast<double*>(p);
                mac_average(kg, d, mac.r0_old[d], mac.r0_new[d], a.sedge[d], r0e[d], S.stream());
                CUDA_OK(S.alloc(&p, sizeof(double) * (size_t)a.sedge[d].sn));
                h0e[d] = static_cast<double*>(p);
                mac_average(kg, d, mac.h0_old[d], mac.h0_new[d], a.sedge[d], h0e[d], S.stream());
            }
        }
        const int rc = fused_enthalpy(kg, e, epar, f, a.state, a.sedge, sf, a.umac, a.force,
                                      sph ? r0e : nullptr, sph ? h0e : nullptr, S.stream());
        if (rc == 0) return S.finish();
        if (rc < 0) return fail(MX_ERR_CUDA, "fused enthalpy kernel launch failed");
    }
    MX_TRY(run_edge(S, kg, epar, a));
    flux_rhoh(kg, f, sf, a.sedge, a.umac, mac, S.stream());
    return S.finish();
}

int mx_enthalpy_edge_flux(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                          const MxFab* sflux, const MxFab* umac, const MxFab* force,
                          const MxEdgeParams* epar, const MxBase* rho0_old,
                          const MxBase* rho0_edge_old, const MxBase* rho0_new,
                          const MxBase* rho0_edge_new, const MxBase* rhoh0_old,
                          const MxBase* rhoh0_edge_old, const MxBase* rhoh0_new,
                          const MxBase* rhoh0_edge_new, const MxFluxParams* fpar,
                          void* stream) {
    if (fpar && fpar->spherical)
        return fail(MX_ERR_UNSUPPORTED,
                    "mx_enthalpy_edge_flux is plane-parallel; spherical: mx_enthalpy_edge_flux_sph");
    return enthalpy_edge_flux_impl(geom, state, sedge, sflux, umac, force, epar, rho0_old,
                                   rho0_edge_old, rho0_new, rho0_edge_new, rhoh0_old,
                                   rhoh0_edge_old, rhoh0_new, rhoh0_edge_new, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, nullptr, fpar, stream);
}

int mx_enthalpy_edge_flux_sph(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                              const MxFab* sflux, const MxFab* umac, const MxFab* force,
                              const MxEdgeParams* epar, const MxFab* r0mac_old,
                              const MxFab* r0mac_new, const MxFab* rh0mac_old,
                              const MxFab* rh0mac_new, const MxFab* h0mac_old,
                              const MxFab* h0mac_new, const MxFluxParams* fpar, void* stream) {
    if (fpar && !fpar->spherical)
        return fail(MX_ERR_INVALID, "mx_enthalpy_edge_flux_sph needs fpar->spherical");
    return enthalpy_edge_flux_impl(geom, state, sedge, sflux, umac, force, epar, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, r0mac_old,
                                   r0mac_new, rh0mac_old, rh0mac_new, h0mac_old, h0mac_new, fpar,
                                   stream);
}

int mx_cell_to_edge(const double* cell, double* edge, int nr, int r_lo, int r_hi, void* stream) {
    if (!cell || !edge || nr < 1 || r_lo < 0 || r_hi >= nr || r_hi < r_lo)
        return fail(MX_ERR_INVALID, "bad cell_to_edge arguments");
    MX_TRY(require_device());
    Session S(stream);
    const double* dc = cell;
    double* de = edge;
    const bool hc = !is_device_ptr(cell), he = !is_device_ptr(edge);
    if (hc) {
        void* p;
        CUDA_OK(S.alloc(&p, sizeof(double) * nr));
        CUDA_OK(cudaMemcpyAsync(p, cell, sizeof(double) * nr, cudaMemcpyHostToDevice, S.stream()));
        dc = static_cast<const double*>(p);
    }
    if (he) {
        void* p;
        CUDA_OK(S.alloc(&p, sizeof(double) * (nr + 1)));
        CUDA_OK(cudaMemcpyAsync(p, edge, sizeof(double) * (nr + 1), cudaMemcpyHostToDevice,
                                S.stream()));
        de = static_cast<double*>(p);
    }
    cell_to_edge(dc, de, nr, r_lo, r_hi, S.stream());
    CUDA_OK(cudaGetLastError());
    if (he) {
        CUDA_OK(cudaMemcpyAsync(edge, de, sizeof(double) * (nr + 1), cudaMemcpyDeviceToHost,
                                S.stream()));
    }
    if (hc || he) CUDA_OK(cudaStreamSynchronize(S.stream()));
    return MX_OK;
}

// ---- Maestro::Makew0 ------------------------------------------------------------------
namespace {
// device view of a 1-D array of n doubles: host arrays are mirrored (uploaded when `in`)
int w0_arr(Session& S, const double* src, size_t n, bool in, const double** dev, bool* host) {
    *host = false;
    if (src == nullptr) return fail(MX_ERR_INVALID, "null base-state array");
    if (is_device_ptr(src)) {
        *dev = src;
        return MX_OK;
    }
    void* p = nullptr;
    CUDA_OK(S.alloc(&p, sizeof(double) * n));
    if (in) CUDA_OK(cudaMemcpyAsync(p, src, sizeof(double) * n, cudaMemcpyHostToDevice, S.stream()));
    *dev = static_cast<const double*>(p);
    *host = true;
    return MX_OK;
}
int check_w0_params(const MxW0Params* p) {
    if (p == nullptr) return fail(MX_ERR_INVALID, "null MxW0Params");
    if (p->nr_fine < 2) return fail(MX_ERR_INVALID, "nr_fine = %d", p->nr_fine);
    if (!(p->dr > 0.0)) return fail(MX_ERR_INVALID, "dr must be positive");
    if (p->base_cutoff_density_coord < 0)
        return fail(MX_ERR_INVALID, "base_cutoff_density_coord = %d", p->base_cutoff_density_coord);
    return MX_OK;
}
}  // namespace

int mx_make_w0_planar(const MxW0Params* par, const double* w0_old, const double* Sbar,
                      const double* p0_old, const double* p0_new, const double* gamma1bar_old,
                      const double* gamma1bar_new, const double* p0_minus_peosbar,
                      const double* etarho_cc, double* w0, double* w0_force, double* p0bdot,
                      void* stream) {
    MX_TRY(check_w0_params(par));
    MX_TRY(require_device());
    Session S(stream);
    const size_t nr = (size_t)par->nr_fine;
    const double *d_w0o, *d_S, *d_p0o, *d_p0n, *d_g1o, *d_g1n, *d_pmp, *d_eta, *d_w0, *d_f;
    bool h, h_w0, h_f;
    MX_TRY(w0_arr(S, w0_old, nr + 1, true, &d_w0o, &h));
    MX_TRY(w0_arr(S, Sbar, nr, true, &d_S, &h));
    MX_TRY(w0_arr(S, p0_old, nr, true, &d_p0o, &h));
    MX_TRY(w0_arr(S, p0_new, nr, true, &d_p0n, &h));
    MX_TRY(w0_arr(S, gamma1bar_old, nr, true, &d_g1o, &h));
    MX_TRY(w0_arr(S, gamma1bar_new, nr, true, &d_g1n, &h));
    MX_TRY(w0_arr(S, p0_minus_peosbar, nr, true, &d_pmp, &h));
    MX_TRY(w0_arr(S, etarho_cc, nr, true, &d_eta, &h));
    MX_TRY(w0_arr(S, w0, nr + 1, false, &d_w0, &h_w0));
    MX_TRY(w0_arr(S, w0_force, nr, false, &d_f, &h_f));
    void *scratch = nullptr, *pb = nullptr;
    CUDA_OK(S.alloc(&scratch, sizeof(double) * w0_scratch_doubles(par->nr_fine)));
    CUDA_OK(S.alloc(&pb, sizeof(double)));
    w0_planar(*par, d_w0o, d_S, d_p0o, d_p0n, d_g1o, d_g1n, d_pmp, d_eta,
              const_cast<double*>(d_w0), const_cast<double*>(d_f), static_cast<double*>(scratch),
              static_cast<double*>(pb), S.stream());
    CUDA_OK(cudaGetLastError());
    if (h_w0) CUDA_OK(cudaMemcpyAsync(w0, d_w0, sizeof(double) * (nr + 1), cudaMemcpyDeviceToHost, S.stream()));
    if (h_f) CUDA_OK(cudaMemcpyAsync(w0_force, d_f, sizeof(double) * nr, cudaMemcpyDeviceToHost, S.stream()));
    if (par->add_pb && p0bdot)
        CUDA_OK(cudaMemcpyAsync(p0bdot, pb, sizeof(double), cudaMemcpyDeviceToHost, S.stream()));
    // Makew0 runs once per time step on a few thousand values: give every caller the
    // reference's synchronous semantics (p0bdot / host arrays are valid on return)
    CUDA_OK(cudaStreamSynchronize(S.stream()));
    return MX_OK;
}

static int make_w0_sphr_impl(bool irreg, const MxW0Params* par, const double* w0_old, const double* Sbar,
                    const double* rho0_old, const double* rho0_new, const double* p0_old,
                    const double* p0_new, const double* gamma1bar_old,
                    const double* gamma1bar_new, const double* p0_minus_peosbar,
                    const double* etarho_cc, const double* etarho_ec, const double* r_cc_loc,
                    const double* r_edge_loc, double* w0, double* w0_force, void* stream) {
    MX_TRY(check_w0_params(par));
    if (irreg && par->base_cutoff_density_coord + 1 > par->nr_fine)
        return fail(MX_ERR_INVALID, "Makew0SphrIrreg: base_cutoff_density_coord + 1 exceeds nr_fine");
    MX_TRY(require_device());
    Session S(stream);
    const size_t nr = (size_t)par->nr_fine;
    const double *d_w0o, *d_S, *d_r0o, *d_r0n, *d_p0o, *d_p0n, *d_g1o, *d_g1n, *d_pmp, *d_ecc,
        *d_eec, *d_rcc, *d_re, *d_w0, *d_f;
    bool h, h_w0, h_f;
    MX_TRY(w0_arr(S, w0_old, nr + 1, true, &d_w0o, &h));
    MX_TRY(w0_arr(S, Sbar, nr, true, &d_S, &h));
    MX_TRY(w0_arr(S, rho0_old, nr, true, &d_r0o, &h));
    MX_TRY(w0_arr(S, rho0_new, nr, true, &d_r0n, &h));
    MX_TRY(w0_arr(S, p0_old, nr, true, &d_p0o, &h));
    MX_TRY(w0_arr(S, p0_new, nr, true, &d_p0n, &h));
    MX_TRY(w0_arr(S, gamma1bar_old, nr, true, &d_g1o, &h));
    MX_TRY(w0_arr(S, gamma1bar_new, nr, true, &d_g1n, &h));
    MX_TRY(w0_arr(S, p0_minus_peosbar, nr, true, &d_pmp, &h));
    MX_TRY(w0_arr(S, etarho_cc, nr, true, &d_ecc, &h));
    MX_TRY(w0_arr(S, etarho_ec, nr + 1, true, &d_eec, &h));
    MX_TRY(w0_arr(S, r_cc_loc, nr, true, &d_rcc, &h));
    MX_TRY(w0_arr(S, r_edge_loc, nr + 1, true, &d_re, &h));
    MX_TRY(w0_arr(S, w0, nr + 1, false, &d_w0, &h_w0));
    MX_TRY(w0_arr(S, w0_force, nr, false, &d_f, &h_f));
    void *scratch = nullptr, *st = nullptr;
    CUDA_OK(S.alloc(&scratch, sizeof(double) * w0_scratch_doubles(par->nr_fine)));
    CUDA_OK(S.alloc(&st, sizeof(int)));
    w0_sphr(*par, d_w0o, d_S, d_r0o, d_r0n, d_p0o, d_p0n, d_g1o, d_g1n, d_pmp, d_ecc, d_eec,
            d_rcc, d_re, const_cast<double*>(d_w0), const_cast<double*>(d_f),
            static_cast<double*>(scratch), static_cast<int*>(st), irreg, S.stream());
    CUDA_OK(cudaGetLastError());
    int status = 0;
    CUDA_OK(cudaMemcpyAsync(&status, st, sizeof(int), cudaMemcpyDeviceToHost, S.stream()));
    if (h_w0) CUDA_OK(cudaMemcpyAsync(w0, d_w0, sizeof(double) * (nr + 1), cudaMemcpyDeviceToHost, S.stream()));
    if (h_f) CUDA_OK(cudaMemcpyAsync(w0_force, d_f, sizeof(double) * nr, cudaMemcpyDeviceToHost, S.stream()));
    CUDA_OK(cudaStreamSynchronize(S.stream()));
    // the reference aborts here (MaestroMakew0.cpp:950-962); the overlay turns this into Abort
    if (status == 1) return fail(MX_ERR_INVALID, "tridiag: CAN NOT HAVE B(0) = 0.0");
    if (status == 2) return fail(MX_ERR_INVALID, "tridiag: TRIDIAG FAILED");
    return MX_OK;
}

int mx_make_w0_sphr(const MxW0Params* par, const double* w0_old, const double* Sbar,
                    const double* rho0_old, const double* rho0_new, const double* p0_old,
                    const double* p0_new, const double* gamma1bar_old,
                    const double* gamma1bar_new, const double* p0_minus_peosbar,
                    const double* etarho_cc, const double* etarho_ec, const double* r_cc_loc,
                    const double* r_edge_loc, double* w0, double* w0_force, void* stream) {
    return make_w0_sphr_impl(false, par, w0_old, Sbar, rho0_old, rho0_new, p0_old, p0_new, gamma1bar_old,
                             gamma1bar_new, p0_minus_peosbar, etarho_cc, etarho_ec, r_cc_loc, r_edge_loc,
                             w0, w0_force, stream);
}

int mx_make_w0_sphr_irreg(const MxW0Params* par, const double* w0_old, const double* Sbar,
                          const double* rho0_old, const double* rho0_new, const double* p0_old,
                          const double* p0_new, const double* gamma1bar_old,
                          const double* gamma1bar_new, const double* p0_minus_peosbar,
                          const double* etarho_cc, const double* etarho_ec, const double* r_cc_loc,
                          const double* r_edge_loc, double* w0, double* w0_force, void* stream) {
    return make_w0_sphr_impl(true, par, w0_old, Sbar, rho0_old, rho0_new, p0_old, p0_new, gamma1bar_old,
                             gamma1bar_new, p0_minus_peosbar, etarho_cc, etarho_ec, r_cc_loc, r_edge_loc,
                             w0, w0_force, stream);
}

// ---- Maestro::UpdateScal --------------------------------------------------------------
int mx_update_scal(const MxGeom* geom, const MxFab* stateold, const MxFab* statenew,
                   const MxFab* sflux, const MxFab* force, const MxUpdateParams* par,
                   unsigned char* eos_mask, long long* n_eos, void* stream) {
    MX_TRY(check_geom(geom));
    if (!stateold || !statenew || !sflux || !force || !par)
        return fail(MX_ERR_INVALID, "null argument");
    if (par->start_comp != par->rhoh_comp && par->start_comp != par->first_spec)
        return fail(MX_ERR_INVALID, "Invalid scalar in UpdateScal().");  // :253
    if (par->num_spec < 1 || par->first_spec < 0 ||
        par->first_spec + par->num_spec > stateold->ncomp || statenew->ncomp != stateold->ncomp ||
        force->ncomp < stateold->ncomp)
        return fail(MX_ERR_INVALID, "inconsistent component counts");
    const bool enth = par->start_comp == par->rhoh_comp;
    if (enth && par->do_eos_h_above_cutoff && !eos_mask)
        return fail(MX_ERR_INVALID,
                    "do_eos_h_above_cutoff: eos_mask is required (the EOS stays with the caller)");
    MX_TRY(check_cover(geom, stateold, 0, -1, "stateold"));
    MX_TRY(check_cover(geom, statenew, 0, -1, "statenew"));
    MX_TRY(check_cover(geom, force, 0, -1, "force"));
    MX_TRY(require_device());
    Session S(stream);
    const KGeom kg = make_kgeom(geom);
    DFab so, sn, fo, fl[3];
    const int c0 = enth ? par->rhoh_comp : par->first_spec;
    const int nc = enth ? 1 : par->num_spec;
    MX_TRY(S.fab(stateold, &so, c0, nc, true, false));
    if (!enth) MX_TRY(S.upload(stateold, so, par->rho_comp, 1));
    // statenew: the swept components are fully overwritten on valid zones only, so a host
    // mirror must start from the caller's data (ghost zones survive the round trip)
    MX_TRY(S.fab(statenew, &sn, c0, nc, true, true));
    if (!enth) {
        MX_TRY(S.upload(statenew, sn, par->rho_comp, 1));
        MX_TRY(S.download_later(statenew, sn, par->rho_comp, 1));
    } else {
        MX_TRY(S.upload(statenew, sn, par->rho_comp, 1));  // read for the EOS flag
    }
    MX_TRY(S.fab(force, &fo, c0, nc, true, false));
    for (int d = 0; d < geom->dim; ++d) {
        MX_TRY(check_cover(geom, &sflux[d], 0, d, "sflux"));
        MX_TRY(S.fab(&sflux[d], &fl[d], c0, nc, true, false));
    }
    const size_t nzones = (size_t)(kg.vhi[0] - kg.vlo[0] + 1) * (kg.vhi[1] - kg.vlo[1] + 1) *
                          (kg.dim == 3 ? kg.vhi[2] - kg.vlo[2] + 1 : 1);
    unsigned char* dmask = eos_mask;
    const bool host_mask = eos_mask && !is_device_ptr(eos_mask);
    if (host_mask) {
        void* p = nullptr;
        CUDA_OK(S.alloc(&p, nzones));
        dmask = static_cast<unsigned char*>(p);
    }
    void* dcount = nullptr;
    CUDA_OK(S.alloc(&dcount, sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), S.stream()));
    update_scal(kg, so, sn, fl, fo, *par, enth ? dmask : nullptr,
                static_cast<unsigned long long*>(dcount), S.stream());
    CUDA_OK(cudaGetLastError());
    if (enth && host_mask)
        CUDA_OK(cudaMemcpyAsync(eos_mask, dmask, nzones, cudaMemcpyDeviceToHost, S.stream()));
    unsigned long long cnt = 0;
    if (n_eos) {
        CUDA_OK(cudaMemcpyAsync(&cnt, dcount, sizeof(cnt), cudaMemcpyDeviceToHost, S.stream()));
        CUDA_OK(cudaStreamSynchronize(S.stream()));
        *n_eos = (long long)cnt;
    }
    if (host_mask) CUDA_OK(cudaStreamSynchronize(S.stream()));
    return S.finish();
}

long long mx_halo_count(const MxGeom* geom, const MxFab* fab, int dir, int ng, int ncomp) {
    if (!geom || !fab || dir < 0 || dir >= geom->dim) return 0;
    long long n = ncomp;
    for (int d = 0; d < 3; ++d) n *= (d == dir) ? ng : (fab->hi[d] - fab->lo[d] + 1);
    return n;
}

static int halo_common(const MxGeom* geom, const MxFab* fab, int dir, int side, int ng,
                       int comp0, int ncomp, double* buf, bool unpack, void* stream) {
    MX_TRY(check_geom(geom));
    if (!fab || !buf) return fail(MX_ERR_INVALID, "null argument");
    if (dir < 0 || dir >= geom->dim || (side != 0 && side != 1) || ng < 1)
        return fail(MX_ERR_INVALID, "bad halo slab");
    if (comp0 < 0 || comp0 + ncomp > fab->ncomp) return fail(MX_ERR_INVALID, "bad components");
    const int nghost = geom->valid_lo[dir] - fab->lo[dir];
    if (ng > nghost) return fail(MX_ERR_INVALID, "fab has only %d ghost cells", nghost);
    MX_TRY(require_device());
    if (!is_device_ptr(fab->p) || !is_device_ptr(buf))
        return fail(MX_ERR_UNSUPPORTED, "halo pack/unpack operates on device memory only");
    const KGeom kg = make_kgeom(geom);
    DFab d = make_dfab(fab, fab->p);
    halo_pack(kg, d, dir, side, ng, comp0, ncomp, buf, unpack, static_cast<cudaStream_t>(stream));
    CUDA_OK(cudaGetLastError());
    return MX_OK;
}

int mx_halo_pack(const MxGeom* geom, const MxFab* fab, int dir, int side, int ng, int comp0,
                 int ncomp, double* buf, void* stream) {
    return halo_common(geom, fab, dir, side, ng, comp0, ncomp, buf, false, stream);
}
int mx_halo_unpack(const MxGeom* geom, const MxFab* fab, int dir, int side, int ng, int comp0,
                   int ncomp, const double* buf, void* stream) {
    return halo_common(geom, fab, dir, side, ng, comp0, ncomp, const_cast<double*>(buf), true,
                       stream);
}

}  // extern "C"

// ==== the passes either side of the path (mx_cart.cu) ======================================
namespace {
// device copy of the MxSphr tables
int make_ksphr(Session& S, const MxSphr* sp, KSphr* out) {
    memset(out, 0, sizeof(*out));
    if (sp == nullptr || !sp->spherical) return MX_OK;
    if (sp->nr_fine < 3 || !sp->r_cc_loc || !sp->r_edge_loc || !(sp->dr_fine > 0.0))
        return fail(MX_ERR_INVALID, "MxSphr: nr_fine / dr_fine / r_cc_loc / r_edge_loc");
    out->spherical = 1;
    out->nr = sp->nr_fine;
    out->drf = sp->dr_fine;
    for (int d = 0; d < 3; ++d) {
        out->plo[d] = sp->prob_lo[d];
        out->cen[d] = sp->center[d];
    }
    out->s0t = sp->s0_interp_type;
    out->w0t = sp->w0_interp_type;
    out->s0mt = sp->s0mac_interp_type;
    out->w0mt = sp->w0mac_interp_type;
    const MxBase cc = {sp->r_cc_loc, 1, sp->nr_fine, 1}, ec = {sp->r_edge_loc, 1, sp->nr_fine + 1, 1};
    int st;
    MX_TRY(S.base(&cc, 0, &out->rcc, &st));
    MX_TRY(S.base(&ec, 0, &out->redge, &st));
    return MX_OK;
}
int check_type(int t, int hi, const char* name) {
    if (t < 1 || t > hi) return fail(MX_ERR_INVALID, "%s = %d out of range 1..%d", name, t, hi);
    return MX_OK;
}
}  // namespace

extern "C" {

int mx_convert_by_rho(const MxGeom* geom, const MxFab* scal, int rho_comp, int start_comp,
                      int num_comp, int flag, void* stream) {
    MX_TRY(check_geom(geom));
    if (!scal) return fail(MX_ERR_INVALID, "null argument");
    if (rho_comp < 0 || rho_comp >= scal->ncomp || start_comp < 0 || num_comp < 0 ||
        start_comp + num_comp > scal->ncomp || (rho_comp >= start_comp && rho_comp < start_comp + num_comp))
        return fail(MX_ERR_INVALID, "bad component range");
    MX_TRY(check_cover(geom, scal, 0, -1, "scal"));
    MX_TRY(require_device());
    Session S(stream);
    DFab d;
    MX_TRY(S.fab(scal, &d, start_comp, num_comp, true, true));
    MX_TRY(S.upload(scal, d, rho_comp, 1));
    cart_convert_by_rho(make_kgeom(geom), d, rho_comp, start_comp, num_comp, flag, S.stream());
    return S.finish();
}

int mx_put_1d_array_on_cart(const MxGeom* geom, const MxBase* s0, int lev, const MxFab* s0_cart,
                            int is_input_edge_centered, int is_output_a_vector,
                            const MxSphr* sphr, void* stream) {
    MX_TRY(check_geom(geom));
    if (!s0 || !s0->p || !s0_cart) return fail(MX_ERR_INVALID, "null argument");
    const bool sph = sphr && sphr->spherical;
    if (sph && geom->dim != 3) return fail(MX_ERR_INVALID, "Spherical is not valid for DIM < 3");
    if (s0_cart->ncomp < (is_output_a_vector ? geom->dim : 1))
        return fail(MX_ERR_INVALID, "s0_cart has too few components");
    if (sph) {
        MX_TRY(check_type(is_input_edge_centered ? sphr->w0_interp_type : sphr->s0_interp_type, 3,
                          is_input_edge_centered ? "w0_interp_type" : "s0_interp_type"));
    } else {
        const int last = geom->dim - 1;
        if (geom->valid_lo[last] < 0 || geom->valid_hi[last] + (is_input_edge_centered ? 1 : 0) >= s0->len)
            return fail(MX_ERR_INVALID, "s0 does not cover the box");
    }
    MX_TRY(check_cover(geom, s0_cart, 0, -1, "s0_cart"));
    MX_TRY(require_device());
    Session S(stream);
    KSphr ks;
    MX_TRY(make_ksphr(S, sphr, &ks));
    const double* p;
    int st;
    MX_TRY(S.base(s0, sph ? 0 : lev, &p, &st));
    DFab d;
    const int oc = sph ? 0 : (is_output_a_vector ? geom->dim - 1 : 0);
    const int nc = (sph && is_output_a_vector) ? 3 : 1;
    // valid zones only are written: a host mirror starts from the caller's data
    MX_TRY(S.fab(s0_cart, &d, oc, nc, true, true));
    cart_put_1d(make_kgeom(geom), p, st, d, is_input_edge_centered, is_output_a_vector, 0, 0, ks, S.stream());
    return S.finish();
}

int mx_put_in_pert_form(const MxGeom* geom, const MxFab* scal, const MxBase* s0, int lev, int comp,
                        int flag, const MxSphr* sphr, void* stream) {
    MX_TRY(check_geom(geom));
    if (!s0 || !s0->p || !scal) return fail(MX_ERR_INVALID, "null argument");
    const bool sph = sphr && sphr->spherical;
    if (sph && geom->dim != 3) return fail(MX_ERR_INVALID, "Spherical is not valid for DIM < 3");
    if (comp < 0 || comp >= scal->ncomp) return fail(MX_ERR_INVALID, "bad component");
    if (sph) {
        MX_TRY(check_type(sphr->s0_interp_type, 3, "s0_interp_type"));
    } else {
        const int last = geom->dim - 1;
        if (geom->valid_lo[last] < 0 || geom->valid_hi[last] >= s0->len)
            return fail(MX_ERR_INVALID, "s0 does not cover the box");
    }
    MX_TRY(check_cover(geom, scal, 0, -1, "scal"));
    MX_TRY(require_device());
    Session S(stream);
    KSphr ks;
    MX_TRY(make_ksphr(S, sphr, &ks));
    const double* p;
    int st;
    MX_TRY(S.base(s0, sph ? 0 : lev, &p, &st));
    DFab d;
    MX_TRY(S.fab(scal, &d, comp, 1, true, true));
    cart_put_1d(make_kgeom(geom), p, st, d, 0, 0, flag ? 1 : 2, comp, ks, S.stream());
    return S.finish();
}

int mx_modify_scal_force(const MxGeom* geom, const MxFab* scal_force, const MxFab* state,
                         const MxFab* umac, const MxBase* s0_edge, int lev, const MxFab* s0_cart,
                         const MxFab* w0_cart, const double* w0, int comp, int fullform, int bcf_hi,
                         const MxFab* s0_edge_cart, const MxSphr* sphr, void* stream) {
    MX_TRY(check_geom(geom));
    if (!scal_force || !state || !umac || !s0_cart) return fail(MX_ERR_INVALID, "null argument");
    const bool sph = sphr && sphr->spherical;
    if (sph && geom->dim != 3) return fail(MX_ERR_INVALID, "ModifyScalForce: Spherical is not valid for DIM < 3");
    if (comp < 0 || comp >= state->ncomp || comp >= scal_force->ncomp)
        return fail(MX_ERR_INVALID, "bad component");
    if (sph) {
        if (!w0) return fail(MX_ERR_INVALID, "spherical: w0 is required");
        MX_TRY(check_type(sphr->s0_interp_type, 3, "s0_interp_type"));
        MX_TRY(check_cover(geom, s0_cart, fullform == 1 ? 0 : 1, -1, "s0_cart"));
    } else {
        if (!w0_cart) return fail(MX_ERR_INVALID, "planar: w0_cart is required");
        if (w0_cart->ncomp < geom->dim) return fail(MX_ERR_INVALID, "w0_cart needs dim components");
        const int last = geom->dim - 1;
        if (s0_edge_cart) {
            for (int d = 0; d < geom->dim; ++d)
                if (s0_edge_cart->lo[d] > geom->valid_lo[d] ||
                    s0_edge_cart->hi[d] < geom->valid_hi[d] + (d == last ? 1 : 0))
                    return fail(MX_ERR_INVALID, "s0_edge_cart needs one ghost cell in the last direction");
        } else {
            if (!s0_edge || !s0_edge->p) return fail(MX_ERR_INVALID, "planar: s0_edge (or s0_edge_cart) is required");
            if (bcf_hi != MX_BC_INT_DIR && bcf_hi != MX_BC_FOEXTRAP && bcf_hi != MX_BC_REFLECT_EVEN)
                return fail(MX_ERR_INVALID, "bcf_hi must be a bcs_f code (int_dir, foextrap, reflect_even)");
            if (geom->valid_lo[last] < 0 || geom->valid_hi[last] + 1 >= s0_edge->len)
                return fail(MX_ERR_INVALID, "s0_edge does not cover the box");
        }
        MX_TRY(check_cover(geom, s0_cart, 0, -1, "s0_cart"));
        for (int d = 0; d < geom->dim; ++d) {  // one ghost cell above in the last direction
            const int ng = d == last ? 1 : 0;
            if (w0_cart->lo[d] > geom->valid_lo[d] || w0_cart->hi[d] < geom->valid_hi[d] + ng)
                return fail(MX_ERR_INVALID, "w0_cart needs one ghost cell in the last direction");
        }
    }
    MX_TRY(check_cover(geom, scal_force, 0, -1, "scal_force"));
    MX_TRY(check_cover(geom, state, 0, -1, "state"));
    MX_TRY(require_device());
    Session S(stream);
    KSphr ks;
    MX_TRY(make_ksphr(S, sphr, &ks));
    DFab f, stt, um[3], s0c, w0c;
    memset(&w0c, 0, sizeof(w0c));
    MX_TRY(S.fab(scal_force, &f, comp, 1, true, true));
    MX_TRY(S.fab(state, &stt, comp, 1, true, false));
    for (int d = 0; d < geom->dim; ++d) {
        MX_TRY(check_cover(geom, &umac[d], 0, d, "umac"));
        MX_TRY(S.fab(&umac[d], &um[d], 0, 1, true, false));
    }
    MX_TRY(S.fab(s0_cart, &s0c, 0, 1, true, false));
    const double* pe = nullptr;
    int ste = 1;
    const double* dw0 = nullptr;
    double* ws = nullptr;
    if (sph) {
        const MxBase bw = {w0, 1, sphr->nr_fine + 1, 1};
        int st;
        MX_TRY(S.base(&bw, 0, &dw0, &st));
        void* q = nullptr;
        CUDA_OK(S.alloc(&q, sizeof(double) * (size_t)sphr->nr_fine));
        ws = static_cast<double*>(q);
    } else {
        MX_TRY(S.fab(w0_cart, &w0c, geom->dim - 1, 1, true, false));
        if (!s0_edge_cart) MX_TRY(S.base(s0_edge, lev, &pe, &ste));
    }
    DFab ecd;
    if (!sph && s0_edge_cart) MX_TRY(S.fab(s0_edge_cart, &ecd, 0, 1, true, false));
    cart_modify_scal_force(make_kgeom(geom), f, stt, um, pe, ste, s0c, w0c, dw0, ws, comp, fullform, bcf_hi,
                           (!sph && s0_edge_cart) ? &ecd : nullptr, ks, S.stream());
    return S.finish();
}

int mx_addw0(const MxGeom* geom, const MxFab* uedge, const MxFab* w0mac, const MxBase* w0, int lev,
             double mult, int spherical, void* stream) {
    MX_TRY(check_geom(geom));
    if (!uedge) return fail(MX_ERR_INVALID, "null argument");
    if (spherical && geom->dim != 3) return fail(MX_ERR_INVALID, "Addw0: Spherical is not valid for DIM < 3");
    if (spherical && !w0mac) return fail(MX_ERR_INVALID, "spherical: w0mac is required");
    if (!spherical && (!w0 || !w0->p)) return fail(MX_ERR_INVALID, "planar: w0 is required");
    MX_TRY(require_device());
    Session S(stream);
    const KGeom kg = make_kgeom(geom);
    DFab ue[3], wm[3];
    const double* pw = nullptr;
    int stw = 1;
    if (spherical) {
        for (int d = 0; d < 3; ++d) {
            MX_TRY(check_cover(geom, &uedge[d], 0, d, "uedge"));
            MX_TRY(check_cover(geom, &w0mac[d], 0, d, "w0mac"));
            MX_TRY(S.fab(&uedge[d], &ue[d], 0, 1, true, true));
            MX_TRY(S.fab(&w0mac[d], &wm[d], 0, 1, true, false));
        }
    } else {
        const int last = geom->dim - 1;
        if (geom->valid_lo[last] < 0 || geom->valid_hi[last] + 1 >= w0->len)
            return fail(MX_ERR_INVALID, "w0 does not cover the box");
        for (int d = 0; d < geom->dim; ++d) {  // one ghost cell in the transverse directions
            const int ng = d == last ? 0 : 1;
            const int hi = geom->valid_hi[d] + (d == last ? 1 : 0);
            if (uedge[last].lo[d] > geom->valid_lo[d] - ng || uedge[last].hi[d] < hi + ng)
                return fail(MX_ERR_INVALID, "uedge needs one transverse ghost cell");
        }
        MX_TRY(S.fab(&uedge[last], &ue[last], 0, 1, true, true));
        MX_TRY(S.base(w0, lev, &pw, &stw));
    }
    cart_addw0(kg, ue, wm, pw, stw, mult, spherical, S.stream());
    return S.finish();
}

int mx_make_w0mac(const MxGeom* geom, const MxFab* w0mac, const MxFab* w0_cart, const double* w0,
                  const MxSphr* sphr, void* stream) {
    MX_TRY(check_geom(geom));
    if (!w0mac || !sphr) return fail(MX_ERR_INVALID, "null argument");
    if (!sphr->spherical || geom->dim != 3) return fail(MX_ERR_INVALID, "Error: only call MakeW0mac for spherical");
    MX_TRY(check_type(sphr->w0mac_interp_type, 4, "w0mac_interp_type"));
    const int type = sphr->w0mac_interp_type;
    if (type == 1 && !w0_cart) return fail(MX_ERR_INVALID, "w0mac_interp_type 1: w0_cart is required");
    if (type != 1 && !w0) return fail(MX_ERR_INVALID, "w0 is required");
    for (int d = 0; d < 3; ++d) MX_TRY(check_cover(geom, &w0mac[d], 1, d, "w0mac"));
    if (type == 1) MX_TRY(check_cover(geom, w0_cart, 2, -1, "w0_cart"));
    MX_TRY(require_device());
    Session S(stream);
    KSphr ks;
    MX_TRY(make_ksphr(S, sphr, &ks));
    DFab wm[3], w0c;
    memset(&w0c, 0, sizeof(w0c));
    for (int d = 0; d < 3; ++d) MX_TRY(S.fab(&w0mac[d], &wm[d], 0, 1, true, true));
    const double* dw0 = nullptr;
    if (type == 1) {
        MX_TRY(S.fab(w0_cart, &w0c, 0, 3, true, false));
    } else {
        const MxBase bw = {w0, 1, sphr->nr_fine + 1, 1};
        int st;
        MX_TRY(S.base(&bw, 0, &dw0, &st));
    }
    cart_make_w0mac(make_kgeom(geom), wm, w0c, dw0, ks, S.stream());
    return S.finish();
}

int mx_make_s0mac(const MxGeom* geom, const MxBase* s0, const MxFab* s0mac, const MxFab* s0_cart,
                  const MxSphr* sphr, void* stream) {
    MX_TRY(check_geom(geom));
    if (!s0mac || !sphr || !s0 || !s0->p) return fail(MX_ERR_INVALID, "null argument");
    if (!sphr->spherical || geom->dim != 3) return fail(MX_ERR_INVALID, "Error: only call MakeS0mac for spherical");
    MX_TRY(check_type(sphr->s0mac_interp_type, 3, "s0mac_interp_type"));
    if (sphr->s0mac_interp_type == 1 && !s0_cart) MX_TRY(check_type(sphr->s0_interp_type, 3, "s0_interp_type"));
    if (sphr->s0mac_interp_type == 1 && s0_cart) MX_TRY(check_cover(geom, s0_cart, 2, -1, "s0_cart"));
    for (int d = 0; d < 3; ++d) MX_TRY(check_cover(geom, &s0mac[d], 1, d, "s0mac"));
    MX_TRY(require_device());
    Session S(stream);
    KSphr ks;
    MX_TRY(make_ksphr(S, sphr, &ks));
    DFab sm[3];
    for (int d = 0; d < 3; ++d) MX_TRY(S.fab(&s0mac[d], &sm[d], 0, 1, true, true));
    const double* p;
    int st;
    MX_TRY(S.base(s0, 0, &p, &st));
    DFab s0c;
    const bool use_cart = sphr->s0mac_interp_type == 1 && s0_cart;
    if (use_cart) MX_TRY(S.fab(s0_cart, &s0c, 0, 1, true, false));
    cart_make_s0mac(make_kgeom(geom), p, st, sm, use_cart ? &s0c : nullptr, ks, S.stream());
    return S.finish();
}

int mx_update_vel(const MxGeom* geom, const MxFab* uold, const MxFab* unew, const MxFab* umac,
                  const MxFab* uedge, const MxFab* force, const MxFab* sponge, const MxFab* w0_cart,
                  const MxFab* w0mac, double dt, int do_sponge, int spherical, void* stream) {
    MX_TRY(check_geom(geom));
    if (!uold || !unew || !umac || !uedge || !force) return fail(MX_ERR_INVALID, "null argument");
    if (spherical && geom->dim != 3) return fail(MX_ERR_INVALID, "UpdateVel: Spherical is not valid for DIM < 3");
    if (spherical ? !w0mac : !w0_cart) return fail(MX_ERR_INVALID, "w0_cart (planar) / w0mac (spherical) is required");
    if (do_sponge && !sponge) return fail(MX_ERR_INVALID, "do_sponge: sponge is required");
    const int dim = geom->dim;
    if (uold->ncomp < dim || unew->ncomp < dim || force->ncomp < dim)
        return fail(MX_ERR_INVALID, "velocity fabs need dim components");
    MX_TRY(check_cover(geom, uold, 0, -1, "uold"));
    MX_TRY(check_cover(geom, unew, 0, -1, "unew"));
    MX_TRY(check_cover(geom, force, 0, -1, "force"));
    MX_TRY(require_device());
    Session S(stream);
    DFab uo, un, fo, sp, w0c, um[3], ue[3], wm[3];
    memset(&sp, 0, sizeof(sp));
    memset(&w0c, 0, sizeof(w0c));
    MX_TRY(S.fab(uold, &uo, 0, dim, true, false));
    MX_TRY(S.fab(unew, &un, 0, dim, true, true));
    MX_TRY(S.fab(force, &fo, 0, dim, true, false));
    if (do_sponge) {
        MX_TRY(check_cover(geom, sponge, 0, -1, "sponge"));
        MX_TRY(S.fab(sponge, &sp, 0, 1, true, false));
    }
    for (int d = 0; d < dim; ++d) {
        MX_TRY(check_cover(geom, &umac[d], 0, d, "umac"));
        MX_TRY(check_cover(geom, &uedge[d], 0, d, "uedge"));
        if (uedge[d].ncomp < dim) return fail(MX_ERR_INVALID, "uedge needs dim components");
        MX_TRY(S.fab(&umac[d], &um[d], 0, 1, true, false));
        MX_TRY(S.fab(&uedge[d], &ue[d], 0, dim, true, false));
        if (spherical) {
            MX_TRY(check_cover(geom, &w0mac[d], 0, d, "w0mac"));
            MX_TRY(S.fab(&w0mac[d], &wm[d], 0, 1, true, false));
        }
    }
    if (!spherical) {
        if (w0_cart->ncomp < dim) return fail(MX_ERR_INVALID, "w0_cart needs dim components");
        const int last = dim - 1;
        for (int d = 0; d < dim; ++d) {
            const int ng = d == last ? 1 : 0;
            if (w0_cart->lo[d] > geom->valid_lo[d] || w0_cart->hi[d] < geom->valid_hi[d] + ng)
                return fail(MX_ERR_INVALID, "w0_cart needs one ghost cell in the last direction");
        }
        MX_TRY(S.fab(w0_cart, &w0c, last, 1, true, false));
    }
    cart_update_vel(make_kgeom(geom), uo, un, um, ue, fo, sp, w0c, wm, dt, do_sponge, spherical, S.stream());
    return S.finish();
}


// ---- MakeUtrans / VelPred -----------------------------------------------------------------
static int vel_params(const MxGeom* g, const MxVelPredParams* p, KVel* e) {
    if (p == nullptr || p->bc == nullptr) return fail(MX_ERR_INVALID, "null velocity-predictor parameters");
    if (p->ppm_type < 0 || p->ppm_type > 2)
        return fail(MX_ERR_UNSUPPORTED, "ppm_type %d (reference supports 0, 1, 2)", p->ppm_type);
    if (p->ppm_type == 0 && p->slope_order != 0 && p->slope_order != 2 && p->slope_order != 4)
        return fail(MX_ERR_UNSUPPORTED, "slope_order %d (reference supports 0, 2, 4)", p->slope_order);
    if (p->spherical && g->dim != 3) return fail(MX_ERR_INVALID, "spherical is 3-D only");
    memset(e, 0, sizeof(*e));
    e->dt = p->dt;
    e->eps = p->rel_eps;
    e->ppm_type = p->ppm_type;
    e->slope_order = p->slope_order;
    e->trace = p->ppm_trace_forces != 0;
    e->spherical = p->spherical != 0;
    for (int d = 0; d < 3; ++d) {
        const double h = d < g->dim ? g->dx[d] : 1.0;
        int ex = 0;
        e->h[d] = h;
        e->inv_h[d] = 1.0 / h;
        e->inv_exact[d] = std::frexp(h, &ex) == 0.5 ? 1 : 0;  // power of two: x * (1/h) is exact
    }
    for (int d = 0; d < g->dim; ++d) {
        e->pbc_lo[d] = p->phys_bc_lo[d];
        e->pbc_hi[d] = p->phys_bc_hi[d];
        if (e->pbc_lo[d] < 0 || e->pbc_lo[d] > 5 || e->pbc_hi[d] < 0 || e->pbc_hi[d] > 5)
            return fail(MX_ERR_INVALID, "invalid boundary type");  // the reference's amrex::Error
        for (int n = 0; n < g->dim; ++n) {
            e->bclo[n][d] = p->bc[n * 6 + d];
            e->bchi[n][d] = p->bc[n * 6 + 3 + d];
        }
    }
    return MX_OK;
}

// w0_cart (planar: last component, one ghost cell in the last direction) or w0mac (spherical)
static int vel_w0(Session& S, const MxGeom* geom, int spherical, const MxFab* w0_cart,
                  const MxFab* w0mac, DFab* w0c, DFab* wm) {
    const int dim = geom->dim;
    memset(w0c, 0, sizeof(*w0c));
    if (spherical) {
        if (!w0mac) return fail(MX_ERR_INVALID, "spherical: w0mac is required");
        for (int d = 0; d < dim; ++d) {
            MX_TRY(check_cover(geom, &w0mac[d], 0, d, "w0mac"));
            MX_TRY(S.fab(&w0mac[d], &wm[d], 0, 1, true, false));
        }
        return MX_OK;
    }
    if (!w0_cart) return fail(MX_ERR_INVALID, "planar: w0_cart is required");
    if (w0_cart->ncomp < dim) return fail(MX_ERR_INVALID, "w0_cart needs dim components");
    const int last = dim - 1;
    for (int d = 0; d < dim; ++d) {
        const int ng = d == last ? 1 : 0;
        if (w0_cart->lo[d] > geom->valid_lo[d] || w0_cart->hi[d] < geom->valid_hi[d] + ng)
            return fail(MX_ERR_INVALID, "w0_cart needs one ghost cell in the last direction");
    }
    return S.fab(w0_cart, w0c, last, 1, true, false);
}

int mx_make_utrans(const MxGeom* geom, const MxFab* utilde, const MxFab* ufull, const MxFab* utrans,
                   const MxFab* w0_cart, const MxFab* w0mac, const MxVelPredParams* par, void* stream) {
    MX_TRY(check_geom(geom));
    if (!utilde || !ufull || !utrans) return fail(MX_ERR_INVALID, "null argument");
    KVel e;
    MX_TRY(vel_params(geom, par, &e));
    const int dim = geom->dim;
    if (utilde->ncomp < dim || ufull->ncomp < dim) return fail(MX_ERR_INVALID, "velocity fabs need dim components");
    MX_TRY(check_cover(geom, utilde, e.ppm_type == 2 ? 4 : 3, -1, "utilde"));
    MX_TRY(check_cover(geom, ufull, 1, -1, "ufull"));
    MX_TRY(require_device());
    Session S(stream);
    DFab ut, uf, w0c, wm[3], tr[3];
    MX_TRY(S.fab(utilde, &ut, 0, dim, true, false));
    MX_TRY(S.fab(ufull, &uf, 0, dim, true, false));
    for (int d = 0; d < dim; ++d) {
        MX_TRY(check_cover(geom, &utrans[d], 0, d, "utrans"));
        // ghost faces keep the caller's values: read-modify-write unless the fab has none
        MX_TRY(S.fab(&utrans[d], &tr[d], 0, 1, !exact_face_box(geom, &utrans[d], d), true));
    }
    MX_TRY(vel_w0(S, geom, e.spherical, w0_cart, w0mac, &w0c, wm));
    vel_make_utrans(make_kgeom(geom), e, ut, uf, tr, w0c, e.spherical ? wm : nullptr, S.stream());
    return S.finish();
}

int mx_vel_pred(const MxGeom* geom, const MxFab* utilde, const MxFab* ufull, const MxFab* utrans,
                const MxFab* umac, const MxFab* w0mac, const MxFab* w0_cart, const MxFab* force,
                const MxVelPredParams* par, void* stream) {
    MX_TRY(check_geom(geom));
    if (!utilde || !ufull || !utrans || !umac || !force) return fail(MX_ERR_INVALID, "null argument");
    KVel e;
    MX_TRY(vel_params(geom, par, &e));
    const int dim = geom->dim;
    if (e.trace && dim == 2)
        return fail(MX_ERR_UNSUPPORTED,
                    "VelPred with ppm_trace_forces in 2-D: the reference reads Ipfy / Imfy without "
                    "filling them (MaestroVelPred.cpp:134-139)");
    if (e.trace && e.ppm_type == 0)
        return fail(MX_ERR_UNSUPPORTED, "ppm_trace_forces needs ppm_type 1 or 2");
    if (utilde->ncomp < dim || ufull->ncomp < dim || force->ncomp < dim)
        return fail(MX_ERR_INVALID, "velocity fabs need dim components");
    const int ngs = e.ppm_type == 2 ? 4 : 3;
    MX_TRY(check_cover(geom, utilde, ngs, -1, "utilde"));
    MX_TRY(check_cover(geom, ufull, 1, -1, "ufull"));
    MX_TRY(check_cover(geom, force, e.trace ? ngs : 1, -1, "force"));
    MX_TRY(require_device());
    Session S(stream);
    DFab ut, uf, fo, w0c, wm[3], tr[3], um[3];
    MX_TRY(S.fab(utilde, &ut, 0, dim, true, false));
    MX_TRY(S.fab(ufull, &uf, 0, dim, true, false));
    MX_TRY(S.fab(force, &fo, 0, dim, true, false));
    for (int d = 0; d < dim; ++d) {
        MX_TRY(check_cover(geom, &utrans[d], 1, d, "utrans"));
        MX_TRY(check_cover(geom, &umac[d], 0, d, "umac"));
        MX_TRY(S.fab(&utrans[d], &tr[d], 0, 1, true, false));
        MX_TRY(S.fab(&umac[d], &um[d], 0, 1, !exact_face_box(geom, &umac[d], d), true));
    }
    MX_TRY(vel_w0(S, geom, e.spherical, w0_cart, w0mac, &w0c, wm));
    const KGeom kg = make_kgeom(geom);
    void* ws = nullptr;
    CUDA_OK(S.alloc(&ws, sizeof(double) * vel_pred_workspace(kg)));
    vel_pred(kg, e, ut, uf, tr, um, e.spherical ? wm : nullptr, w0c, fo, static_cast<double*>(ws), S.stream());
    return S.finish();
}

}  // extern "C"

// ==== fused advances: device-resident composition of the entry points above =================
namespace {
// device fab over box [lo, hi] with ncomp components out of the session's workspace
int ws_fab(Session& S, const int lo[3], const int hi[3], int ncomp, MxFab* out, bool zero) {
    for (int d = 0; d < 3; ++d) {
        out->lo[d] = lo[d];
        out->hi[d] = hi[d];
    }
    out->ncomp = ncomp;
    size_t n = (size_t)ncomp;
    for (int d = 0; d < 3; ++d) n *= (size_t)(hi[d] - lo[d] + 1);
    void* p = nullptr;
    CUDA_OK(S.alloc(&p, n * sizeof(double)));
    if (zero) CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(double), S.stream()));
    out->p = static_cast<double*>(p);
    return MX_OK;
}
void valid_box3(const MxGeom* g, int nodal_dir, int lo[3], int hi[3]) {
    for (int d = 0; d < 3; ++d) {
        lo[d] = d < g->dim ? g->valid_lo[d] : 0;
        hi[d] = d < g->dim ? g->valid_hi[d] + (d == nodal_dir ? 1 : 0) : 0;
    }
}
// device view of a caller fab as an MxFab (uploads comps [c0, c0+nc) of a host fab)
int dev_view(Session& S, const MxFab* f, int c0, int nc, bool in, bool out, MxFab* v, DFab* d) {
    MX_TRY(S.fab(f, d, c0, nc, in, out));
    *v = *f;
    v->p = d->p;
    return MX_OK;
}
}  // namespace

extern "C" {

int mx_density_advance(const MxGeom* geom, const MxFab* state, const MxFab* snew,
                       const MxFab* etarhoflux, const MxFab* umac, const MxFab* force,
                       int species_force_is_zero, const MxEdgeParams* epar, int rho_bccomp,
                       const MxBase* rho0_old, const MxBase* rho0_edge_old, const MxFab* r0mac_old,
                       const MxBase* rho0_new, const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                       const MxBase* rho0_predicted_edge, const MxBase* w0, const MxFluxParams* fpar,
                       const MxBase* rho0_pert, const MxSphr* sphr, const MxUpdateParams* upar,
                       const MxFab* sedge, const MxFab* sflux, void* stream) {
    MX_TRY(check_geom(geom));
    if (!state || !snew || !umac || !force || !epar || !fpar || !upar)
        return fail(MX_ERR_INVALID, "null argument");
    if (geom->dim != 3) return fail(MX_ERR_UNSUPPORTED, "mx_density_advance is 3-D only");
    if (upar->start_comp != upar->first_spec || upar->first_spec != fpar->first_spec ||
        upar->num_spec != fpar->num_spec || upar->rho_comp != fpar->rho_comp)
        return fail(MX_ERR_INVALID, "update parameters do not describe the species update of this pass");
    const bool pert = fpar->species_pred_type == MX_PRED_RHOPRIME_AND_X;
    if (pert && (!rho0_pert || !rho0_pert->p)) return fail(MX_ERR_INVALID, "rho0_pert is required");
    if (snew->ncomp != state->ncomp) return fail(MX_ERR_INVALID, "snew / state component counts differ");
    MX_TRY(check_cover(geom, snew, 0, -1, "snew"));
    MX_TRY(require_device());
    Session S(stream);
    const int rho = fpar->rho_comp, c0 = fpar->first_spec, nc = fpar->num_spec;
    // 1. device views of the inputs (host fabs are uploaded once, here)
    MxFab v_state, v_force, v_umac[3], v_snew, v_eta;
    DFab d_state, d_force, d_umac[3], d_snew, d_eta;
    MX_TRY(dev_view(S, state, c0, nc, true, false, &v_state, &d_state));
    MX_TRY(S.upload(state, d_state, rho, 1));
    if (species_force_is_zero && !is_device_ptr(force->p)) {
        MX_TRY(dev_view(S, force, rho, 1, true, false, &v_force, &d_force));
        CUDA_OK(cudaMemsetAsync(d_force.comp(c0), 0, sizeof(double) * (size_t)d_force.sn * nc, S.stream()));
    } else {
        MX_TRY(dev_view(S, force, c0, nc, true, false, &v_force, &d_force));
        MX_TRY(S.upload(force, d_force, rho, 1));
    }
    for (int d = 0; d < 3; ++d) MX_TRY(dev_view(S, &umac[d], 0, 1, true, false, &v_umac[d], &d_umac[d]));
    // snew: the species / density components are fully overwritten on the valid zones and
    // nothing else is touched: no upload, and only the valid box travels back
    MX_TRY(dev_view(S, snew, c0, nc, false, false, &v_snew, &d_snew));
    {
        int vlo[3], vhi[3];
        valid_box3(geom, -1, vlo, vhi);
        MX_TRY(S.download_box(snew, d_snew, c0, nc, vlo, vhi));
        MX_TRY(S.download_box(snew, d_snew, rho, 1, vlo, vhi));
    }
    const bool do_eta = !fpar->spherical && fpar->evolve_base_state && !fpar->use_exact_base_state;
    if (etarhoflux) MX_TRY(dev_view(S, etarhoflux, 0, 1, true, do_eta, &v_eta, &d_eta));
    // 2. edge states + fluxes into workspace (or the caller's fabs)
    MxFab w_sedge[3], w_sflux[3];
    DFab dd;
    for (int d = 0; d < 3; ++d) {
        int lo[3], hi[3];
        valid_box3(geom, d, lo, hi);
        if (sedge) {
            MX_TRY(dev_view(S, &sedge[d], c0, nc, false, true, &w_sedge[d], &dd));
            MX_TRY(S.download_later(&sedge[d], dd, rho, 1));
        } else {
            MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_sedge[d], false));
        }
        if (sflux) {
            MX_TRY(dev_view(S, &sflux[d], c0, nc, false, true, &w_sflux[d], &dd));
            MX_TRY(S.download_later(&sflux[d], dd, rho, 1));
        } else {
            MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_sflux[d], false));
        }
    }
    MX_TRY(mx_density_edge_flux(geom, &v_state, w_sedge, w_sflux, etarhoflux ? &v_eta : nullptr, v_umac,
                                &v_force, epar, rho_bccomp, rho0_old, rho0_edge_old, r0mac_old, rho0_new,
                                rho0_edge_new, r0mac_new, rho0_predicted_edge, w0, fpar, S.stream()));
    // 3. the old state as UpdateScal sees it: rho' + rho0 -> rho, X rho -> rho X (valid zones)
    MxFab w_sold, w_zero;
    int lo[3], hi[3];
    valid_box3(geom, -1, lo, hi);
    MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_sold, false));
    MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_zero, true));  // scal_force.setVal(0), :241-243
    const KGeom kg = make_kgeom(geom);
    const DFab d_sold = make_dfab(&w_sold, w_sold.p);
    cart_copy_valid(kg, d_state, d_sold, c0, nc, S.stream());
    cart_copy_valid(kg, d_state, d_sold, rho, 1, S.stream());
    if (pert) MX_TRY(mx_put_in_pert_form(geom, &w_sold, rho0_pert, fpar->lev, rho, 0, sphr, S.stream()));
    MX_TRY(mx_convert_by_rho(geom, &w_sold, rho, c0, nc, 0, S.stream()));
    // 4. UpdateScal (species + density)
    MX_TRY(mx_update_scal(geom, &w_sold, &v_snew, w_sflux, &w_zero, upar, nullptr, nullptr, S.stream()));
    return S.finish();
}

int mx_enthalpy_advance(const MxGeom* geom, const MxFab* state, const MxFab* snew, const MxFab* umac,
                        const MxFab* force, const MxFab* force_update, const MxEdgeParams* epar,
                        const MxBase* rho0_old, const MxBase* rho0_edge_old, const MxBase* rho0_new,
                        const MxBase* rho0_edge_new, const MxBase* rhoh0_old,
                        const MxBase* rhoh0_edge_old, const MxBase* rhoh0_new,
                        const MxBase* rhoh0_edge_new, const MxFluxParams* fpar,
                        const MxBase* rhoh0_pert, const MxUpdateParams* upar, unsigned char* eos_mask,
                        long long* n_eos, void* stream) {
    MX_TRY(check_geom(geom));
    if (!state || !snew || !umac || !force || !force_update || !epar || !fpar || !upar)
        return fail(MX_ERR_INVALID, "null argument");
    if (geom->dim != 3) return fail(MX_ERR_UNSUPPORTED, "mx_enthalpy_advance is 3-D only");
    if (fpar->spherical) return fail(MX_ERR_UNSUPPORTED, "mx_enthalpy_advance: planar base state only");
    if (fpar->enthalpy_pred_type != MX_PRED_RHOH && fpar->enthalpy_pred_type != MX_PRED_RHOHPRIME)
        return fail(MX_ERR_UNSUPPORTED, "mx_enthalpy_advance: enthalpy_pred_type %d (only predict_rhoh / "
                    "predict_rhohprime)", fpar->enthalpy_pred_type);
    if (upar->start_comp != upar->rhoh_comp || upar->rhoh_comp != fpar->rhoh_comp)
        return fail(MX_ERR_INVALID, "update parameters do not describe the enthalpy update of this pass");
    const bool pert = fpar->enthalpy_pred_type == MX_PRED_RHOHPRIME;
    if (pert && (!rhoh0_pert || !rhoh0_pert->p)) return fail(MX_ERR_INVALID, "rhoh0_pert is required");
    if (snew->ncomp != state->ncomp) return fail(MX_ERR_INVALID, "snew / state component counts differ");
    MX_TRY(check_cover(geom, snew, 0, -1, "snew"));
    MX_TRY(require_device());
    Session S(stream);
    const int rhoh = fpar->rhoh_comp, rho = fpar->rho_comp;
    MxFab v_state, v_force, v_fupd, v_umac[3], v_snew;
    DFab d_state, d_force, d_fupd, d_umac[3], d_snew;
    MX_TRY(dev_view(S, state, rhoh, 1, true, false, &v_state, &d_state));
    MX_TRY(dev_view(S, force, rhoh, 1, true, false, &v_force, &d_force));
    MX_TRY(dev_view(S, force_update, rhoh, 1, true, false, &v_fupd, &d_fupd));
    for (int d = 0; d < 3; ++d) MX_TRY(dev_view(S, &umac[d], 0, 1, true, false, &v_umac[d], &d_umac[d]));
    MX_TRY(dev_view(S, snew, rhoh, 1, false, false, &v_snew, &d_snew));
    {
        int vlo[3], vhi[3];
        valid_box3(geom, -1, vlo, vhi);
        MX_TRY(S.download_box(snew, d_snew, rhoh, 1, vlo, vhi));
    }
    MX_TRY(S.upload(snew, d_snew, rho, 1));  // the new density decides the EOS zones
    MxFab w_sedge[3], w_sflux[3];
    for (int d = 0; d < 3; ++d) {
        int lo[3], hi[3];
        valid_box3(geom, d, lo, hi);
        MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_sedge[d], false));
        MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_sflux[d], false));
    }
    MX_TRY(mx_enthalpy_edge_flux(geom, &v_state, w_sedge, w_sflux, v_umac, &v_force, epar, rho0_old,
                                 rho0_edge_old, rho0_new, rho0_edge_new, rhoh0_old, rhoh0_edge_old,
                                 rhoh0_new, rhoh0_edge_new, fpar, S.stream()));
    MxFab w_sold;
    int lo[3], hi[3];
    valid_box3(geom, -1, lo, hi);
    MX_TRY(ws_fab(S, lo, hi, state->ncomp, &w_sold, false));
    const KGeom kg = make_kgeom(geom);
    cart_copy_valid(kg, d_state, make_dfab(&w_sold, w_sold.p), rhoh, 1, S.stream());
    if (pert) MX_TRY(mx_put_in_pert_form(geom, &w_sold, rhoh0_pert, fpar->lev, rhoh, 0, nullptr, S.stream()));
    // eos_mask / n_eos: device or host, exactly as mx_update_scal takes them
    MX_TRY(mx_update_scal(geom, &w_sold, &v_snew, w_sflux, &v_fupd, upar, eos_mask, n_eos, S.stream()));
    return S.finish();
}

}  // extern "C"

extern "C" int mx_make_w0_planar_ml(const MxW0Params* par, const MxRadialGeom* rg, const double* w0_old,
                                    const double* Sbar, const double* p0_old, const double* p0_new,
                                    const double* gamma1bar_old, const double* gamma1bar_new,
                                    const double* p0_minus_peosbar, const double* etarho_cc, double* w0,
                                    double* w0_force, double* p0bdot, void* stream) {
    if (!par || !rg || !w0_old || !Sbar || !p0_old || !p0_new || !gamma1bar_old || !gamma1bar_new ||
        !p0_minus_peosbar || !etarho_cc || !w0 || !w0_force)
        return fail(MX_ERR_INVALID, "null argument");
    const int nlev = rg->max_radial_level + 1, nf = rg->finest_radial_level + 1;
    if (rg->nr_fine < 2 || nlev < 1 || nf < 1 || nf > nlev || rg->chunk_len < 2 || !rg->dr || !rg->nr ||
        !rg->base_cutoff_density_coord || !rg->numdisjointchunks || !rg->r_start_coord || !rg->r_end_coord)
        return fail(MX_ERR_INVALID, "MxRadialGeom is incomplete");
    for (int n = 0; n < nf; ++n) {
        if (rg->numdisjointchunks[n] < 1 || rg->numdisjointchunks[n] >= rg->chunk_len)
            return fail(MX_ERR_INVALID, "numdisjointchunks(%d) does not fit the coordinate tables", n);
        for (int j = 1; j <= rg->numdisjointchunks[n]; ++j) {
            const int lo = rg->r_start_coord[n * rg->chunk_len + j], hi = rg->r_end_coord[n * rg->chunk_len + j];
            if (lo < 0 || hi < lo || hi >= rg->nr[n] || rg->nr[n] > rg->nr_fine)
                return fail(MX_ERR_INVALID, "chunk %d of radial level %d is outside the level", j, n);
        }
    }
    MX_TRY(require_device());
    Session S(stream);
    const size_t lc = (size_t)nlev * rg->nr_fine, le = (size_t)nlev * (rg->nr_fine + 1);
    auto up = [&](const void* h, size_t bytes, const void** d) -> int {
        void* p = nullptr;
        CUDA_OK(S.alloc(&p, bytes));
        CUDA_OK(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, S.stream()));
        *d = p;
        return MX_OK;
    };
    const void *d_dr, *d_nr, *d_cut, *d_nc, *d_rs, *d_re, *d_w0o, *d_S, *d_p0o, *d_p0n, *d_g1o, *d_g1n, *d_pmp, *d_ecc;
    MX_TRY(up(rg->dr, sizeof(double) * nlev, &d_dr));
    MX_TRY(up(rg->nr, sizeof(int) * nlev, &d_nr));
    MX_TRY(up(rg->base_cutoff_density_coord, sizeof(int) * nlev, &d_cut));
    MX_TRY(up(rg->numdisjointchunks, sizeof(int) * nf, &d_nc));
    MX_TRY(up(rg->r_start_coord, sizeof(int) * nf * rg->chunk_len, &d_rs));
    MX_TRY(up(rg->r_end_coord, sizeof(int) * nf * rg->chunk_len, &d_re));
    MX_TRY(up(w0_old, sizeof(double) * le, &d_w0o));
    MX_TRY(up(Sbar, sizeof(double) * lc, &d_S));
    MX_TRY(up(p0_old, sizeof(double) * lc, &d_p0o));
    MX_TRY(up(p0_new, sizeof(double) * lc, &d_p0n));
    MX_TRY(up(gamma1bar_old, sizeof(double) * lc, &d_g1o));
    MX_TRY(up(gamma1bar_new, sizeof(double) * lc, &d_g1n));
    MX_TRY(up(p0_minus_peosbar, sizeof(double) * lc, &d_pmp));
    MX_TRY(up(etarho_cc, sizeof(double) * lc, &d_ecc));
    void *d_w0, *d_f, *d_scr, *d_pb;
    CUDA_OK(S.alloc(&d_w0, sizeof(double) * le));
    CUDA_OK(S.alloc(&d_f, sizeof(double) * lc));
    CUDA_OK(S.alloc(&d_scr, sizeof(double) * (2 * (size_t)rg->nr_fine + 4)));
    CUDA_OK(S.alloc(&d_pb, sizeof(double)));
    CUDA_OK(cudaMemsetAsync(d_pb, 0, sizeof(double), S.stream()));
    w0_planar_ml(*par, rg->max_radial_level, rg->finest_radial_level, rg->nr_fine, rg->chunk_len,
                 (const double*)d_dr, (const int*)d_nr, (const int*)d_cut, (const int*)d_nc, (const int*)d_rs,
                 (const int*)d_re, (const double*)d_w0o, (const double*)d_S, (const double*)d_p0o,
                 (const double*)d_p0n, (const double*)d_g1o, (const double*)d_g1n, (const double*)d_pmp,
                 (const double*)d_ecc, (double*)d_w0, (double*)d_f, (double*)d_scr, (double*)d_pb, S.stream());
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(w0, d_w0, sizeof(double) * le, cudaMemcpyDeviceToHost, S.stream()));
    CUDA_OK(cudaMemcpyAsync(w0_force, d_f, sizeof(double) * lc, cudaMemcpyDeviceToHost, S.stream()));
    double pb = 0.0;
    CUDA_OK(cudaMemcpyAsync(&pb, d_pb, sizeof(double), cudaMemcpyDeviceToHost, S.stream()));
    CUDA_OK(cudaStreamSynchronize(S.stream()));
    if (p0bdot && par->add_pb) *p0bdot = pb;
    return MX_OK;
}
