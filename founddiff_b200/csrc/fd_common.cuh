// Shared device helpers for the founddiff_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/founddiff_b200.h"

#define FD_DEVINL __device__ __forceinline__

// fp16 stores saturate to +-65504 instead of producing inf (one F2FP.SATFINITE either way): the fp16-stored tensors are
// normalisation-bounded, this only makes an out-of-range value degrade gracefully.
FD_DEVINL __half2 fd_floats2half2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return *reinterpret_cast<__half2*>(&r);
}
FD_DEVINL __half fd_float2half_sat(float v) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return *reinterpret_cast<__half*>(&r);
}


// Every extern "C" entry point returns 0 or a cudaError_t value; nothing throws across the ABI.
#define FD_LAUNCH_CHECK()                        \
    do {                                         \
        cudaError_t e__ = cudaPeekAtLastError(); \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

template <typename T> struct fd_type;
template <> struct fd_type<float> {
    static FD_DEVINL float ld(const float* p) { return *p; }
    static FD_DEVINL void st(float* p, float v) { *p = v; }
};
template <> struct fd_type<__nv_bfloat16> {
    static FD_DEVINL float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static FD_DEVINL void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct fd_type<__half> {
    static FD_DEVINL float ld(const __half* p) { return __half2float(*p); }
    static FD_DEVINL void st(__half* p, float v) { *p = fd_float2half_sat(v); }
};

template <typename T> FD_DEVINL float fd_ld(const T* p) { return fd_type<T>::ld(p); }
template <typename T> FD_DEVINL void fd_st(T* p, float v) { fd_type<T>::st(p, v); }

// Vector of 8 (16-bit types) or 4 (float) elements = 16 bytes.
template <typename T> struct fd_vec { static constexpr int N = 16 / sizeof(T); };

template <typename T, int N> FD_DEVINL void fd_ldv(const T* p, float (&v)[N]);
template <> FD_DEVINL void fd_ldv<float, 4>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> FD_DEVINL void fd_ldv<__nv_bfloat16, 8>(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> FD_DEVINL void fd_ldv<__half, 8>(const __half* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T, int N> FD_DEVINL void fd_stv(T* p, const float (&v)[N]);
template <> FD_DEVINL void fd_stv<float, 4>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> FD_DEVINL void fd_stv<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
}
template <> FD_DEVINL void fd_stv<__half, 8>(__half* p, const float (&v)[8]) {
    uint4 t;
    __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = fd_floats2half2_sat(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
}

// MUFU.EX2 + MUFU.RCP, no IEEE-division slow path (FCHK + branch): the full-precision divide cost ~2x the instructions
FD_DEVINL float fd_silu(float x) {      // FMUL, MUFU.EX2 (ftz: no denormal rescaling code), FADD, MUFU.RCP, FMUL
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    return __fdividef(x, 1.f + e);
}
// SiLU for results that are stored in a 16-bit type: x * sigmoid(x) = 0.5 x (1 + tanh(x / 2)) with ONE MUFU op
// (tanh.approx.f32, 2^-11 relative) instead of EX2 + RCP
FD_DEVINL float fd_silu16(float x) {
    float t;
    const float hx = 0.5f * x;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
}
FD_DEVINL float fd_softplus20(float x) { return x <= 20.f ? log1pf(__expf(x)) : x; }

FD_DEVINL float fd_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
FD_DEVINL float fd_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// packed fp32 pairs in 64-bit registers (fma / mul / add .f32x2 = FFMA2 / FMUL2 / FADD2 on sm_100a: one issue slot, two results)
typedef unsigned long long u64;
FD_DEVINL u64 f2_pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
FD_DEVINL void f2_unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
FD_DEVINL u64 f2_fma(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
FD_DEVINL u64 f2_mul(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
FD_DEVINL u64 f2_add(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// ---- warp-level tensor-core helpers (mma.sync m16n8k16, fp32 accumulate) for the small, memory-bound GEMMs that sit
// inside fused kernels (Gram matrices, x_proj/dt_proj); the large convolutions / projections use tcgen05 (fd_conv_tc.cu).
FD_DEVINL void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
FD_DEVINL void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <typename T> FD_DEVINL void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> FD_DEVINL void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> FD_DEVINL void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// dtype dispatch for the extern "C" wrappers
#define FD_DISPATCH_DTYPE(dtype, T, ...)                                           \
    switch (dtype) {                                                               \
        case FD_F32: { using T = float; __VA_ARGS__; break; }                      \
        case FD_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }             \
        case FD_F16: { using T = __half; __VA_ARGS__; break; }                     \
        default: return FD_ERR_BAD_ARGUMENT;                                       \
    }

static inline int fd_cdiv(long a, long b) { return (int)((a + b - 1) / b); }
