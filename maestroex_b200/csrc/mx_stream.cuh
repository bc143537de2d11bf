// mx_stream.cuh — small device helpers shared by the streaming kernels (warp shuffles of
// doubles, cp.async staging, the correctly rounded x / h_d sequence).  Included by mx_fused.cu,
// which owns FusedArgs and the host-side dispatch.
#ifndef MX_STREAM_CUH
#define MX_STREAM_CUH

MXD double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
MXD double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

MXD void cp_async8(double* dst_smem, const double* src, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = pred ? 8 : 0;  // src-size 0: zero-fill, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(n)
                 : "memory");
}
MXD void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
MXD void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// ---- TMA / mbarrier (sm_90+) -------------------------------------------------------------
MXD unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
MXD void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
MXD void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
                 : "memory");
}
MXD void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MX_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MX_DONE_%=;\n\t"
        "bra MX_WAIT_%=;\n\t"
        "MX_DONE_%=:\n\t}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// one box {PW, PH, 1, 1} of the state fab (x, y, z, component) -> shared memory; completion is
// signalled on `bar` (complete_tx of the box's bytes); out-of-range elements arrive as zeros
MXD void tma_load_plane(double* dst, const void* tmap, unsigned long long* bar, int c0, int c1,
                        int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

MXD void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }
MXD void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// x / h_D for the loop-invariant mesh spacing, correctly rounded (== IEEE division): a power
// of two multiplies exactly; otherwise q0 = RN(x * RN(1/h)) is refined twice with exact FMA
// residuals.  After the first step q is a faithful rounding of x/h, so the second step is
// Markstein's theorem (RN(q + RN(x - h q) * RN(1/h)) = RN(x / h)).  Barring under/overflow
// of the intermediate products (|x| = |u| dt is never near either); tests/test_divconst.py
// checks the sequence against exact rational arithmetic.
template <int D, class T>
MXD T div_h(const FusedArgs& A, T x) {
    const double y = A.inv_h[D];
    T q = x * y;
    if (!A.inv_exact[D]) {
        const double h = A.h[D];
        T r = fma_(-h, q, x);
        q = fma_(r, y, q);
        r = fma_(-h, q, x);
        q = fma_(r, y, q);
    }
    return q;
}

constexpr int ST_MAXZC = 64;  // largest z chunk (size of the base-state tables)


#endif
