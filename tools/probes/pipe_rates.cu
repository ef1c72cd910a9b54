// Issue / pipe rate probe for the scan inner loop on sm_100a: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), MUFU.EX2, and the
// state-update mix (FMUL, EX2, FMUL, FFMA, FFMA per state) in scalar and packed form.  Prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ex2(float a) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float x[8];
    u64 X[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; X[i] = pk(x[i], x[i] + 1.f); }
    const float a = s, b = 1.f - s;
    const u64 A = pk(a, a), B = pk(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = ffma(x[i], a, b);            // 32 FFMA
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], A, B);            // 32 FFMA2
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = ex2(x[i]);                   // 32 MUFU
        } else if (MODE == 3) {                                                 // scalar state update x 8 states: 40 instr, 8 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float e = ex2(fmul(a, x[(i + 1) & 7] ));
                const float db = fmul(b, a);
                x[i] = ffma(e, x[i], db);
                x[(i + 4) & 7] = ffma(x[i], b, x[(i + 4) & 7]);
            }
        } else if (MODE == 4) {                                                 // packed: 4 pairs: FMUL2, 2 EX2, FMUL2, FFMA2, FFMA2 = 24 instr, 8 MUFU
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 t = mul2(A, X[(i + 1) & 7]);
                const float e0 = ex2(__uint_as_float((unsigned)t)), e1 = ex2(__uint_as_float((unsigned)(t >> 32)));
                const u64 db = mul2(B, A);
                X[i] = fma2(pk(e0, e1), X[i], db);
                X[i + 4] = fma2(X[i], B, X[i + 4]);
            }
        } else if (MODE == 5) {                                                 // 16 FFMA + 8 MUFU interleaved (independent)
#pragma unroll
            for (int i = 0; i < 8; ++i) { x[i] = ffma(x[i], a, b); x[i] = ffma(x[i], a, b); }
#pragma unroll
            for (int i = 0; i < 8; ++i) X[i] = pk(ex2(__uint_as_float((unsigned)X[i])), 0.f);
        } else if (MODE == 6) {                                                 // 32 FFMA2 + 8 MUFU
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) X[i] = fma2(X[i], A, B);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = ex2(x[i]);
        }
    }
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) acc += x[i] + __uint_as_float((unsigned)X[i]) + __uint_as_float((unsigned)(X[i] >> 32));
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE> void run(const char* name, int instr_per_iter, int warps_per_sm) {
    float* out; cudaMalloc(&out, 4);
    const int iters = 4096, sms = 148;
    const int blocks = sms * (warps_per_sm / 8 > 0 ? warps_per_sm / 8 : 1);
    const int threads = warps_per_sm >= 8 ? 256 : warps_per_sm * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double wi = (double)iters * instr_per_iter * warps_per_sm;
    printf("%-28s warps/SM %2d  %.3f ms  warp-instr/clk/SM %.2f  (clk %d kHz nominal)\n", name, warps_per_sm, ms, wi / cycles, clk);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) {
        if (w == 4) { run<0>("FFMA x32", 32, 4); run<1>("FFMA2 x32", 32, 4); run<2>("EX2 x32", 32, 4); run<3>("state mix scalar (40, 8 MUFU)", 40, 4); run<4>("state mix packed (24, 8 MUFU)", 24, 4); run<5>("16 FFMA + 8 EX2", 24, 4); run<6>("32 FFMA2 + 8 EX2", 40, 4); }
        if (w == 8) { run<0>("FFMA x32", 32, 8); run<1>("FFMA2 x32", 32, 8); run<2>("EX2 x32", 32, 8); run<3>("state mix scalar (40, 8 MUFU)", 40, 8); run<4>("state mix packed (24, 8 MUFU)", 24, 8); run<5>("16 FFMA + 8 EX2", 24, 8); run<6>("32 FFMA2 + 8 EX2", 40, 8); }
        if (w == 16) { run<0>("FFMA x32", 32, 16); run<1>("FFMA2 x32", 32, 16); run<2>("EX2 x32", 32, 16); run<3>("state mix scalar (40, 8 MUFU)", 40, 16); run<4>("state mix packed (24, 8 MUFU)", 24, 16); run<5>("16 FFMA + 8 EX2", 24, 16); run<6>("32 FFMA2 + 8 EX2", 40, 16); }
        if (w == 32) { run<0>("FFMA x32", 32, 32); run<1>("FFMA2 x32", 32, 32); run<2>("EX2 x32", 32, 32); run<3>("state mix scalar (40, 8 MUFU)", 40, 32); run<4>("state mix packed (24, 8 MUFU)", 24, 32); run<5>("16 FFMA + 8 EX2", 24, 32); run<6>("32 FFMA2 + 8 EX2", 40, 32); }
    }
    return 0;
}
