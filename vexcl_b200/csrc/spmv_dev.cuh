// Device helpers shared by the sparse kernels (spmv.cu, distapply.cu): unfused arithmetic, L2 cache policies,
// streaming / keeping loads, the ELL row body.
#pragma once
#include "common.cuh"
#include "spmat.hpp"

namespace vexb {

template <class T> __device__ __forceinline__ T t_mul(T a, T b);
template <> __device__ __forceinline__ double t_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float t_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <class T> __device__ __forceinline__ T t_add(T a, T b);
template <> __device__ __forceinline__ double t_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float t_add<float>(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// L2 residency control.  The matrix streams through once per product, x is gathered ~nnz/ncols
// times: matrix traffic is marked evict-first and x evict-last, so the stream does not push x out
// of the 126 MB L2 (x gathers that miss L1 then cost an L2 hit, not an HBM round trip).
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ double ldg_keep(const double *p, uint64_t policy) {
    double v; asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy)); return v;
}
__device__ __forceinline__ float ldg_keep(const float *p, uint64_t policy) {
    float v; asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy)); return v;
}
__device__ __forceinline__ int ldg_stream(const int *p, uint64_t policy) {
    int v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy)); return v;
}
__device__ __forceinline__ short ldg_stream(const short *p, uint64_t policy) {
    short v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(policy)); return v;
}
// Column of an ELL slot.  32-bit storage holds it directly (-1 = padding); 16-bit storage (spmv.col16) holds its
// distance from (row + shift of its slot), with -32768 = padding: 2 bytes less HBM traffic per stored entry for banded matrices.
__device__ __forceinline__ int ell_column(int raw, size_t, int) { return raw; }
__device__ __forceinline__ int ell_shift_of(const EllShifts &sh, int slot) { return sh.s[slot < kEllShiftSlots ? slot : kEllShiftSlots - 1]; }
__device__ __forceinline__ int ell_column(short raw, size_t row, int shift) { return raw == (short)-32768 ? -1 : (int)row + shift + (int)raw; }
__device__ __forceinline__ double ldg_stream(const double *p, uint64_t policy) {
    double v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy)); return v;
}
__device__ __forceinline__ float ldg_stream(const float *p, uint64_t policy) {
    float v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy)); return v;
}

template <class T>
__device__ __forceinline__ void store_y(T *y, size_t r, T sum, T alpha, int append) {
    const T v = t_mul<T>(alpha, sum);
    y[r] = append ? t_add<T>(y[r], v) : v;
}

// One row of a hybrid-ELL strip (hybrid_ell.inl:252-268): ELL slots in order, then the CSR tail; products and sums rounded
// separately.  W > 0: fully unrolled, all 2W streaming loads and then the W gathers of x in flight at once.
template <class T, int W, class C>
__device__ __forceinline__ T hell_row_sum(size_t i, size_t pitch, int w_dyn, const C *__restrict__ ell_col, const EllShifts &shift,
                                          const T *__restrict__ ell_val, const int *__restrict__ tail_ptr,
                                          const int *__restrict__ tail_col, const T *__restrict__ tail_val,
                                          const T *__restrict__ x, uint64_t stream, uint64_t keep) {
    T sum = T(0);
    if (W > 0) {
        int c[W > 0 ? W : 1]; T v[W > 0 ? W : 1]; T xv[W > 0 ? W : 1];
#pragma unroll
        for (int j = 0; j < W; ++j) { c[j] = ell_column(ldg_stream(ell_col + i + (size_t)j * pitch, stream), i, ell_shift_of(shift, j)); v[j] = ldg_stream(ell_val + i + (size_t)j * pitch, stream); }
#pragma unroll
        for (int j = 0; j < W; ++j) xv[j] = (c[j] != -1) ? ldg_keep(x + c[j], keep) : T(0);
#pragma unroll
        for (int j = 0; j < W; ++j) if (c[j] != -1) sum = t_add<T>(sum, t_mul<T>(v[j], xv[j]));
    } else {
        // any width: plain dependent loop.  Measured faster on irregular matrices than batching 4 columns
        // (4.2 vs 3.3 TB/s effective at average width 12): occupancy hides the latency, and a padded slot
        // (column -1) costs 4 bytes, not 12, because its value is never fetched.
        for (int j = 0; j < w_dyn; ++j) {
            const int c = ell_column(ldg_stream(ell_col + i + (size_t)j * pitch, stream), i, shift.s[0]);   // run-time widths use one shift (spmv.cu build())
            if (c != -1) sum = t_add<T>(sum, t_mul<T>(ldg_stream(ell_val + i + (size_t)j * pitch, stream), ldg_keep(x + c, keep)));
        }
    }
    if (tail_ptr) {
        for (int j = tail_ptr[i], e = tail_ptr[i + 1]; j < e; ++j) sum = t_add<T>(sum, t_mul<T>(tail_val[j], __ldg(x + tail_col[j])));
    }
    return sum;
}

} // namespace vexb
