// Hand-written bodies for the recognised expression shapes (exprhost.hpp) and
// 256-bit global memory access helpers (LDG.E.256 / STG.E.256 on sm_100a).
#pragma once
#include "exprhost.hpp"

namespace vexb {

template <class T> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
};

// Each shape: K vector inputs, value f(v[0..K), s[0..2)).
template <int ID> struct Shape;
#define VEXB_SHAPE(ID, KK, EXPR) \
    template <> struct Shape<ID> { static constexpr int K = KK; \
        template <class T> static __device__ __forceinline__ T f(const T *v, const T *s) { typedef Arith<T> A; (void)v; (void)s; return EXPR; } };
VEXB_SHAPE(SH_COPY,    1, v[0])
VEXB_SHAPE(SH_FILL,    0, s[0])
VEXB_SHAPE(SH_ADD,     2, A::add(v[0], v[1]))
VEXB_SHAPE(SH_SUB,     2, A::sub(v[0], v[1]))
VEXB_SHAPE(SH_MUL,     2, A::mul(v[0], v[1]))
VEXB_SHAPE(SH_DIV,     2, A::div(v[0], v[1]))
VEXB_SHAPE(SH_SQR,     1, A::mul(v[0], v[0]))
VEXB_SHAPE(SH_SCALE,   1, A::mul(s[0], v[0]))
VEXB_SHAPE(SH_MULADD,  3, A::add(v[0], A::mul(v[1], v[2])))
VEXB_SHAPE(SH_AXPY,    2, A::add(A::mul(s[0], v[0]), v[1]))
VEXB_SHAPE(SH_XPAY,    2, A::add(v[0], A::mul(s[0], v[1])))
VEXB_SHAPE(SH_XMAY,    2, A::sub(v[0], A::mul(s[0], v[1])))
VEXB_SHAPE(SH_AXPBY,   2, A::add(A::mul(s[0], v[0]), A::mul(s[1], v[1])))
VEXB_SHAPE(SH_ABSDIFF, 2, A::abs(A::sub(v[0], v[1])))
#undef VEXB_SHAPE

struct SweepArgs {
    const void *v[3];
    double s[2];
    const void *sp[2];      // non-NULL: the scalar lives in device memory (VEXB_TERM_DSCALAR), of the kernel's type T
};

template <class T>
__device__ __forceinline__ T sweep_scalar(const SweepArgs &a, int k) {
    return a.sp[k] ? *static_cast<const T *>(a.sp[k]) : static_cast<T>(a.s[k]);
}

// 32 bytes = 4 doubles or 8 floats, moved with one 256-bit instruction.
struct alignas(32) Vec256 { unsigned long long w[4]; };

__device__ __forceinline__ Vec256 ldg256(const void *p) {
    Vec256 r;
    asm volatile("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(void *p, const Vec256 &r) {
    asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(r.w[0]), "l"(r.w[1]), "l"(r.w[2]), "l"(r.w[3]) : "memory");
}

template <class T> struct Lanes;
template <> struct Lanes<double> {
    static constexpr int E = 4;
    static __device__ __forceinline__ double get(const Vec256 &r, int j) { return __longlong_as_double((long long)r.w[j]); }
    static __device__ __forceinline__ void set(Vec256 &r, int j, double x) { r.w[j] = (unsigned long long)__double_as_longlong(x); }
};
template <> struct Lanes<float> {
    static constexpr int E = 8;
    static __device__ __forceinline__ float get(const Vec256 &r, int j) {
        const unsigned u = (j & 1) ? (unsigned)(r.w[j >> 1] >> 32) : (unsigned)r.w[j >> 1];
        return __uint_as_float(u);
    }
    static __device__ __forceinline__ void set(Vec256 &r, int j, float x) {
        const unsigned long long u = __float_as_uint(x);
        if (j & 1) r.w[j >> 1] = (r.w[j >> 1] & 0xffffffffull) | (u << 32);
        else       r.w[j >> 1] = (r.w[j >> 1] & 0xffffffff00000000ull) | u;
    }
};

inline bool aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

} // namespace vexb
