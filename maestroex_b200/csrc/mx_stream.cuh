// mx_stream.cuh — small device helpers shared by the streaming kernels (warp shuffles of
// doubles, cp.async staging, the correctly rounded x / h_d sequence).  Included by mx_fused.cu,
// which owns FusedArgs and the host-side dispatch.
#ifndef MX_STREAM_CUH
#define MX_STREAM_CUH

MXD double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
MXD double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
MXD D2 shfl_dn1(D2 v) { return D2{shfl_dn1(v.a), shfl_dn1(v.b)}; }

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
