// S6 selective scan forward (the op behind selective_scan_cuda_core.fwd, src/emamba2.py:152-154).
//
// Mapping: one WARP per (batch, channel) row; the row is consumed in chunks of 32 lanes x 8 consecutive
// time steps.  Per state n the lane composes its 8 (a, b) pairs locally, a 5-round shuffle scan composes the
// lane aggregates, and the running state h[n] (held redundantly in every lane) carries across chunks — i.e. a
// chunked warp-shuffle scan with no shared memory and no block barrier.  Loads/stores are 16-byte vectors,
// consecutive lanes touching consecutive addresses.  exp(dt*A) is one MUFU.EX2 per (step, state): the kernel
// is bound by the SFU/FMA pipes rather than by HBM for dstate >= 4 ("selective scan").
#include "fd_common.cuh"

namespace {

constexpr int kItems = 8;
constexpr int kChunk = 32 * kItems;
constexpr int kWarpsPerBlock = 4;

FD_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T>
FD_DEVINL void load_items(const T* __restrict__ row, int l0, int L, bool vec_ok, float (&v)[kItems]) {
    if (vec_ok && l0 + kItems <= L) {
        if constexpr (sizeof(T) == 2) {
            fd_ldv<T, 8>(row + l0, v);
        } else {
            float a[4], b[4];
            fd_ldv<float, 4>(reinterpret_cast<const float*>(row) + l0, a);
            fd_ldv<float, 4>(reinterpret_cast<const float*>(row) + l0 + 4, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) v[i] = (l0 + i < L) ? fd_ld(row + l0 + i) : 0.f;
    }
}

template <typename T>
FD_DEVINL void store_items(T* __restrict__ row, int l0, int L, bool vec_ok, const float (&v)[kItems]) {
    if (vec_ok && l0 + kItems <= L) {
        if constexpr (sizeof(T) == 2) {
            fd_stv<T, 8>(row + l0, v);
        } else {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = v[i]; b[i] = v[4 + i]; }
            fd_stv<float, 4>(reinterpret_cast<float*>(row) + l0, a);
            fd_stv<float, 4>(reinterpret_cast<float*>(row) + l0 + 4, b);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i)
            if (l0 + i < L) fd_st(row + l0 + i, v[i]);
    }
}

template <typename T, int NS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) selective_scan_kernel(
    const T* __restrict__ u, const T* __restrict__ delta, const float* __restrict__ A, const float* __restrict__ Bm,
    const float* __restrict__ Cm, const float* __restrict__ D, const float* __restrict__ delta_bias, T* __restrict__ y,
    long rows, int dim, int L, int N, int G, int softplus, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / dim), d = (int)(row % dim);
    const int g = d / (dim / G);
    const T* ur = u + row * (long)L;
    const T* dr = delta + row * (long)L;
    T* yr = y + row * (long)L;
    const float* Bg = Bm + ((long)b * G + g) * N * (long)L;
    const float* Cg = Cm + ((long)b * G + g) * N * (long)L;
    const float bias = delta_bias ? delta_bias[d] : 0.f;
    const float Dd = D ? D[d] : 0.f;

    float A2[NS], h[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        A2[n] = (n < N) ? A[(long)d * N + n] * 1.4426950408889634f : 0.f;
        h[n] = 0.f;
    }

    for (int c0 = 0; c0 < L; c0 += kChunk) {
        const int l0 = c0 + lane * kItems;
        float dt[kItems], dtu[kItems], yacc[kItems];
        load_items<T>(dr, l0, L, vec_ok, dt);
        load_items<T>(ur, l0, L, vec_ok, dtu);
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            float t = dt[i] + bias;
            if (softplus) t = fd_softplus20(t);
            if (l0 + i >= L) t = 0.f;  // identity element beyond the end of the row
            yacc[i] = Dd * dtu[i];
            dtu[i] = t * dtu[i];
            dt[i] = t;
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            if (n < N) {
                float Bn[kItems], Cn[kItems];
                load_items<float>(Bg + (long)n * L, l0, L, vec_ok, Bn);
                load_items<float>(Cg + (long)n * L, l0, L, vec_ok, Cn);
                float a[kItems], ap = 1.f, bp = 0.f;
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    a[i] = ex2_approx(dt[i] * A2[n]);
                    Bn[i] = dtu[i] * Bn[i];
                    ap *= a[i];
                    bp = fmaf(a[i], bp, Bn[i]);
                }
                // inclusive scan of the lane aggregates: (a2,b2) o (a1,b1) = (a1*a2, a2*b1 + b2)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float au = __shfl_up_sync(0xffffffffu, ap, o);
                    const float bu = __shfl_up_sync(0xffffffffu, bp, o);
                    if (lane >= o) { bp = fmaf(ap, bu, bp); ap *= au; }
                }
                float ae = __shfl_up_sync(0xffffffffu, ap, 1);
                float be = __shfl_up_sync(0xffffffffu, bp, 1);
                if (lane == 0) { ae = 1.f; be = 0.f; }
                float hh = fmaf(ae, h[n], be);  // state entering this lane's first step
                const float at = __shfl_sync(0xffffffffu, ap, 31);
                const float bt = __shfl_sync(0xffffffffu, bp, 31);
                h[n] = fmaf(at, h[n], bt);
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    hh = fmaf(a[i], hh, Bn[i]);
                    yacc[i] = fmaf(hh, Cn[i], yacc[i]);
                }
            }
        }
        store_items<T>(yr, l0, L, vec_ok, yacc);
    }
}

template <typename T>
int scan_launch(const void* u, const void* delta, const float* A, const float* Bm, const float* Cm, const float* D,
                const float* delta_bias, void* y, int batch, int dim, int L, int N, int G, int softplus,
                cudaStream_t st) {
    const long rows = (long)batch * dim;
    const int vec_ok = (L % kItems == 0) && ((((uintptr_t)u | (uintptr_t)delta | (uintptr_t)y) & 31) == 0) &&
                       ((((uintptr_t)Bm | (uintptr_t)Cm) & 31) == 0);
    const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
#define SCAN_CASE(NSV)                                                                                              \
    if (N <= NSV) {                                                                                                 \
        selective_scan_kernel<T, NSV><<<grid, kWarpsPerBlock * 32, 0, st>>>((const T*)u, (const T*)delta, A, Bm, Cm, D, \
                                                                            delta_bias, (T*)y, rows, dim, L, N, G,   \
                                                                            softplus, vec_ok);                       \
        FD_LAUNCH_CHECK();                                                                                          \
        return 0;                                                                                                   \
    }
    SCAN_CASE(4) SCAN_CASE(8) SCAN_CASE(16) SCAN_CASE(32) SCAN_CASE(64)
#undef SCAN_CASE
    return FD_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int fd_selective_scan_fwd(const void* u, const void* delta, const float* A, const float* Bm,
                                     const float* Cm, const float* D, const float* delta_bias, void* y, int batch,
                                     int dim, int seqlen, int dstate, int ngroups, int delta_softplus, int io_dtype,
                                     cudaStream_t stream) {
    if (!u || !delta || !A || !Bm || !Cm || !y) return FD_ERR_BAD_ARGUMENT;
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0 || ngroups <= 0 || dim % ngroups) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(io_dtype, T,
                      return scan_launch<T>(u, delta, A, Bm, Cm, D, delta_bias, y, batch, dim, seqlen, dstate, ngroups,
                                            delta_softplus, stream));
    return 0;
}
