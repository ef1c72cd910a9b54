// LayerNorm+modulate, GroupNorm statistics / apply, small dense layers, sinusoidal embedding.
// All HBM-bound row kernels: 16-byte vector accesses, sub-warp shuffles for the per-row reductions.
#include <type_traits>

#include "fd_common.cuh"

// ------------------------------------------------------------------------------------------------------
// Row kernels over channels-last data.  LPR lanes cooperate on one row; each lane owns NV 16-byte vectors
// (interleaved: vector j of lane i is vector i + j*LPR of the row) so that a warp's accesses are fully coalesced.
// A warp handles U consecutive groups of 32/LPR rows: all U*NV raw vectors are requested before anything is
// consumed (bytes in flight, not occupancy, carry these kernels), and the per-channel parameters of the lane's
// columns are fetched once and reused for the U rows while they belong to the same sample.
// ------------------------------------------------------------------------------------------------------
template <typename T> FD_DEVINL void fd_raw_to_f(const uint4& r, float (&v)[16 / sizeof(T)]) {
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
    } else {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f;
            if constexpr (std::is_same<T, __nv_bfloat16>::value) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            else f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
}
template <typename T> FD_DEVINL uint4 fd_f_to_raw(const float (&v)[16 / sizeof(T)]) {
    uint4 r;
    if constexpr (sizeof(T) == 4) {
        r.x = __float_as_uint(v[0]); r.y = __float_as_uint(v[1]); r.z = __float_as_uint(v[2]); r.w = __float_as_uint(v[3]);
    } else {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
            } else {
                __half2 h = fd_floats2half2_sat(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        r.x = w[0]; r.y = w[1]; r.z = w[2]; r.w = w[3];
    }
    return r;
}
FD_DEVINL void fd_ld_f32v(const float* p, float* v, int n) {   // n is 4 or 8, p 16-byte aligned
    for (int i = 0; i < n; i += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// mean / rstd of one row held as NV vectors per lane across LPR lanes (two-pass in registers)
template <int LPR, int NV, int VEC>
FD_DEVINL void fd_row_stats(const float (&v)[NV][VEC], int C, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) s += v[j][e];
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) { const float d = v[j][e] - mean; q += d * d; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q / (float)C + eps);
}

// LayerNorm over C + adaLN modulate (src/DADiff.py:450-451, 459, 461, 486-487).
template <typename T, typename TO, int LPR, int NV, int U>      // T: storage of x, TO: storage of out (same width)
__global__ void __launch_bounds__(256) ln_modulate_kernel(const T* __restrict__ x, TO* __restrict__ out,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ scale, int mod_stride,
                                                          unsigned rows, unsigned P, int C, float eps) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int RPW = 32 / LPR;  // rows per warp per group
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned row0 = warp * (RPW * U) + lane / LPR;
    uint4 raw[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned row = row0 + u * RPW;
        const T* xr = x + (long)(row < rows ? row : 0) * C;
#pragma unroll
        for (int j = 0; j < NV; ++j) raw[u][j] = *reinterpret_cast<const uint4*>(xr + (sub + j * LPR) * VEC);
    }
    // per-channel parameters folded to y = xhat * a + c with a = gamma*(1+scale), c = beta*(1+scale) + shift
    constexpr bool HOIST = NV <= 2;   // wider rows would spill; they read the (L1-resident) parameters per use instead
    float pa[HOIST ? NV : 1][VEC], pc[HOIST ? NV : 1][VEC];
    unsigned bcur = 0xffffffffu;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned row = row0 + u * RPW;
        const bool active = row < rows;
        const unsigned b = (active ? row : 0) / P;
        if (HOIST && b != bcur) {
            bcur = b;
            const float* sh = shift + (long)b * mod_stride;
            const float* sc = scale + (long)b * mod_stride;
#pragma unroll
            for (int j = 0; j < (HOIST ? NV : 0); ++j) {
                const int c0 = (sub + j * LPR) * VEC;
                float tsc[VEC], tsh[VEC];
                fd_ld_f32v(sc + c0, tsc, VEC);
                fd_ld_f32v(sh + c0, tsh, VEC);
                if (gamma) {
                    float tg[VEC], tb[VEC];
                    fd_ld_f32v(gamma + c0, tg, VEC);
                    fd_ld_f32v(beta + c0, tb, VEC);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) { pa[j][e] = tg[e] * (1.f + tsc[e]); pc[j][e] = fmaf(tb[e], 1.f + tsc[e], tsh[e]); }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) { pa[j][e] = 1.f + tsc[e]; pc[j][e] = tsh[e]; }
                }
            }
        }
        float v[NV][VEC];
#pragma unroll
        for (int j = 0; j < NV; ++j) fd_raw_to_f<T>(raw[u][j], v[j]);
        float mean, rstd;
        fd_row_stats<LPR, NV, VEC>(v, C, eps, mean, rstd);
        if (active) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                if constexpr (HOIST) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v[j][e] = fmaf((v[j][e] - mean) * rstd, pa[j][e], pc[j][e]);
                } else {
                    const int c0 = (sub + j * LPR) * VEC;
                    float tsc[VEC], tsh[VEC];
                    fd_ld_f32v(scale + (long)b * mod_stride + c0, tsc, VEC);
                    fd_ld_f32v(shift + (long)b * mod_stride + c0, tsh, VEC);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        float t = (v[j][e] - mean) * rstd;
                        if (gamma) t = fmaf(t, __ldg(gamma + c0 + e), __ldg(beta + c0 + e));
                        v[j][e] = fmaf(t, 1.f + tsc[e], tsh[e]);
                    }
                }
                *reinterpret_cast<uint4*>(out + (long)row * C + (sub + j * LPR) * VEC) = fd_f_to_raw<TO>(v[j]);
            }
        }
    }
}

template <typename T, typename TO>
static int ln_modulate_launch(const void* x, void* out, const float* gamma, const float* beta,
                              const float* shift, const float* scale, int mod_stride, int B, int P, int C,
                              float eps, cudaStream_t st) {
    constexpr int VEC = fd_vec<T>::N;
    if (C % VEC || mod_stride % 4 || ((uintptr_t)shift | (uintptr_t)scale | (uintptr_t)gamma | (uintptr_t)beta) % 16)
        return FD_ERR_UNSUPPORTED;
    const int vr = C / VEC;  // vectors per row
    const long rows = (long)B * P;
    if (rows >= (1L << 31)) return FD_ERR_UNSUPPORTED;
    const int lpr = vr >= 32 ? 32 : vr;
    const int nv = vr / lpr;
    if (lpr * nv != vr || (lpr & (lpr - 1))) return FD_ERR_UNSUPPORTED;
    const int rpw = 32 / lpr;
    const int warps = 8;
#define LN_CASE(L, N, UU)                                                                                    \
    if (lpr == L && nv == N) {                                                                               \
        const int grid = fd_cdiv(rows, (long)rpw * warps * UU);                                              \
        ln_modulate_kernel<T, TO, L, N, UU><<<grid, warps * 32, 0, st>>>((const T*)x, (TO*)out, gamma, beta, shift, scale, \
                                                                     mod_stride, (unsigned)rows, (unsigned)P, C, eps); \
        FD_LAUNCH_CHECK();                                                                                   \
        return 0;                                                                                            \
    }
    LN_CASE(4, 1, 4) LN_CASE(8, 1, 4) LN_CASE(16, 1, 4) LN_CASE(32, 1, 4) LN_CASE(32, 2, 2) LN_CASE(32, 4, 1) LN_CASE(32, 8, 1)
#undef LN_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_ln_modulate(const void* x, void* out, const float* gamma, const float* beta,
                              const float* shift, const float* scale, int mod_stride, int B, int P, int C,
                              float eps, int dtype, cudaStream_t stream) {
    if (!x || !out || !shift || !scale || B <= 0 || P <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((gamma == nullptr) != (beta == nullptr)) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T,
                      return (ln_modulate_launch<T, T>(x, out, gamma, beta, shift, scale, mod_stride, B, P, C, eps, stream)));
    return 0;
}

extern "C" int fd_ln_modulate_io(const void* x, void* out, const float* gamma, const float* beta, const float* shift,
                                 const float* scale, int mod_stride, int B, int P, int C, float eps, int in_dtype, int out_dtype,
                                 cudaStream_t stream) {
    if (in_dtype == out_dtype)
        return fd_ln_modulate(x, out, gamma, beta, shift, scale, mod_stride, B, P, C, eps, in_dtype, stream);
    if (!x || !out || !shift || !scale || B <= 0 || P <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((gamma == nullptr) != (beta == nullptr)) return FD_ERR_BAD_ARGUMENT;
    if (in_dtype == FD_F16 && out_dtype == FD_BF16)
        return ln_modulate_launch<__half, __nv_bfloat16>(x, out, gamma, beta, shift, scale, mod_stride, B, P, C, eps, stream);
    if (in_dtype == FD_BF16 && out_dtype == FD_F16)
        return ln_modulate_launch<__nv_bfloat16, __half>(x, out, gamma, beta, shift, scale, mod_stride, B, P, C, eps, stream);
    return FD_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------------
// SS2D tail on channels-last data (src/emamba2.py:365, 747-748): out = (LN(y) * gamma + beta) * z + local[b]
// with z = columns [z_off, z_off + C) of rows of pitch `ld` (the silu(z) half of xz).  Same row mapping as above.
// ------------------------------------------------------------------------------------------------------
template <typename T, int LPR, int NV, int U>
__global__ void __launch_bounds__(256) ln_gate_kernel(const T* __restrict__ y, const T* __restrict__ xz, int ld, int z_off,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ local, T* __restrict__ out, unsigned rows,
                                                      unsigned P, int C, float eps) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned row0 = warp * (RPW * U) + lane / LPR;
    uint4 raw[U][NV], rawz[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned row = row0 + u * RPW;
        const long rr = row < rows ? row : 0;
        const T* yr = y + rr * C;
        const T* zr = xz + rr * ld + z_off;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            raw[u][j] = *reinterpret_cast<const uint4*>(yr + (sub + j * LPR) * VEC);
            rawz[u][j] = *reinterpret_cast<const uint4*>(zr + (sub + j * LPR) * VEC);
        }
    }
    constexpr bool HOIST = NV <= 2;
    constexpr int NH = HOIST ? NV : 1;
    float pg[NH][VEC], pb[NH][VEC], pl[NH][VEC];
#pragma unroll
    for (int j = 0; j < (HOIST ? NV : 0); ++j) {
        fd_ld_f32v(gamma + (sub + j * LPR) * VEC, pg[j], VEC);
        fd_ld_f32v(beta + (sub + j * LPR) * VEC, pb[j], VEC);
    }
    unsigned bcur = 0xffffffffu;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned row = row0 + u * RPW;
        const bool active = row < rows;
        const unsigned b = (active ? row : 0) / P;
        if (HOIST && b != bcur) {
            bcur = b;
#pragma unroll
            for (int j = 0; j < (HOIST ? NV : 0); ++j) fd_ld_f32v(local + (long)b * C + (sub + j * LPR) * VEC, pl[j], VEC);
        }
        float v[NV][VEC];
#pragma unroll
        for (int j = 0; j < NV; ++j) fd_raw_to_f<T>(raw[u][j], v[j]);
        float mean, rstd;
        fd_row_stats<LPR, NV, VEC>(v, C, eps, mean, rstd);
        if (active) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float z[VEC];
                fd_raw_to_f<T>(rawz[u][j], z);
                if constexpr (HOIST) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const float n = fmaf((v[j][e] - mean) * rstd, pg[j][e], pb[j][e]);
                        v[j][e] = fmaf(n, z[e], pl[j][e]);
                    }
                } else {
                    const int c0 = (sub + j * LPR) * VEC;
                    float tg[VEC], tb[VEC], tl[VEC];
                    fd_ld_f32v(gamma + c0, tg, VEC);
                    fd_ld_f32v(beta + c0, tb, VEC);
                    fd_ld_f32v(local + (long)b * C + c0, tl, VEC);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const float n = fmaf((v[j][e] - mean) * rstd, tg[e], tb[e]);
                        v[j][e] = fmaf(n, z[e], tl[e]);
                    }
                }
                *reinterpret_cast<uint4*>(out + (long)row * C + (sub + j * LPR) * VEC) = fd_f_to_raw<T>(v[j]);
            }
        }
    }
}

template <typename T>
static int ln_gate_launch(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta,
                          const float* local, void* out, int B, int P, int C, float eps, cudaStream_t st) {
    constexpr int VEC = fd_vec<T>::N;
    if (C % VEC || ld % VEC || z_off % VEC || ((uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)local) % 16) return FD_ERR_UNSUPPORTED;
    const int vr = C / VEC;
    const long rows = (long)B * P;
    if (rows >= (1L << 31)) return FD_ERR_UNSUPPORTED;
    const int lpr = vr >= 32 ? 32 : vr;
    const int nv = vr / lpr;
    if (lpr * nv != vr || (lpr & (lpr - 1))) return FD_ERR_UNSUPPORTED;
    const int rpw = 32 / lpr, warps = 8;
#define LG_CASE(L, N, UU)                                                                                          \
    if (lpr == L && nv == N) {                                                                                     \
        const int grid = fd_cdiv(rows, (long)rpw * warps * UU);                                                    \
        ln_gate_kernel<T, L, N, UU><<<grid, warps * 32, 0, st>>>((const T*)y, (const T*)xz, ld, z_off, gamma, beta, local, \
                                                                 (T*)out, (unsigned)rows, (unsigned)P, C, eps);    \
        FD_LAUNCH_CHECK();                                                                                         \
        return 0;                                                                                                  \
    }
    LG_CASE(2, 1, 4) LG_CASE(4, 1, 4) LG_CASE(8, 1, 4) LG_CASE(16, 1, 4) LG_CASE(32, 1, 2) LG_CASE(32, 2, 1) LG_CASE(32, 4, 1)
    LG_CASE(32, 8, 1)
#undef LG_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_ln_gate(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta,
                          const float* local, void* out, int B, int P, int C, float eps, int dtype, cudaStream_t stream) {
    if (!y || !xz || !gamma || !beta || !local || !out || B <= 0 || P <= 0 || C <= 0 || ld < z_off + C) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, return ln_gate_launch<T>(y, xz, ld, z_off, gamma, beta, local, out, B, P, C, eps, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// Per-row LayerNorm statistic for a folded GEMM (fd_conv_params.ln_rstd): rstd = 1 / sqrt(var + eps), two-pass in registers as
// every other LayerNorm here.  LPR lanes share a row (one 16-byte vector each, NV vectors per lane), U row groups per warp in flight.
// ------------------------------------------------------------------------------------------------------
template <typename T, int LPR, int NV, int U>
__global__ void __launch_bounds__(256) row_rstd_kernel(const T* __restrict__ x, float* __restrict__ rstd, long rows, int C, float eps) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long row0 = warp * (RPW * U) + lane / LPR;
    uint4 raw[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long row = row0 + u * RPW;
        const T* xr = x + (row < rows ? row : 0) * C;
#pragma unroll
        for (int j = 0; j < NV; ++j) raw[u][j] = *reinterpret_cast<const uint4*>(xr + (sub + j * LPR) * VEC);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long row = row0 + u * RPW;
        float v[NV][VEC];
#pragma unroll
        for (int j = 0; j < NV; ++j) fd_raw_to_f<T>(raw[u][j], v[j]);
        float mean, rs;
        fd_row_stats<LPR, NV, VEC>(v, C, eps, mean, rs);
        if (sub == 0 && row < rows) rstd[row] = rs;
    }
}

extern "C" int fd_row_rstd(const void* x, float* rstd, long rows, int C, float eps, int dtype, cudaStream_t stream) {
    if (!x || !rstd || rows <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((dtype != FD_BF16 && dtype != FD_F16) || ((uintptr_t)x & 15)) return FD_ERR_UNSUPPORTED;
#define RR_CASE(CV, L, N, UU)                                                                                         \
    if (C == CV) {                                                                                                    \
        const int grid = fd_cdiv(rows, (long)(32 / L) * 8 * UU);                                                      \
        if (dtype == FD_BF16) row_rstd_kernel<__nv_bfloat16, L, N, UU><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, rstd, rows, C, eps); \
        else row_rstd_kernel<__half, L, N, UU><<<grid, 256, 0, stream>>>((const __half*)x, rstd, rows, C, eps);      \
        FD_LAUNCH_CHECK();                                                                                            \
        return 0;                                                                                                     \
    }
    RR_CASE(64, 8, 1, 4) RR_CASE(128, 16, 1, 4) RR_CASE(256, 32, 1, 4) RR_CASE(512, 32, 2, 2)
#undef RR_CASE
    return FD_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm + adaLN modulation folded into the weights of the 1x1 GEMM that consumes it (see fd_ln_fold in the header).
// One warp per (sample, output row): Wf = round(W g - rowmean(W g)), v = W h; fixed shuffle tree -> reproducible.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ln_fold_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ shift,
                                                      const float* __restrict__ scale, int mod_stride, T* __restrict__ Wf,
                                                      float* __restrict__ v, int Cout, int C) {
    const int lane = threadIdx.x & 31, o = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y;
    if (o >= Cout) return;
    const float* wr = W + (long)o * C;
    const float* sc = scale + (long)b * mod_stride;
    const float* sh = shift + (long)b * mod_stride;
    T* out = Wf + ((long)b * Cout + o) * C;
    float gs = 0.f, vs = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float s1 = 1.f + sc[c];
        const float g = gamma ? gamma[c] * s1 : s1;
        const float h = (beta ? beta[c] * s1 : 0.f) + sh[c];
        const float w = wr[c];
        gs = fmaf(w, g, gs);
        vs = fmaf(w, h, vs);
    }
    gs = fd_warp_sum(gs) / (float)C;          // row mean of W g: subtracting it makes the row blind to the input's mean
    vs = fd_warp_sum(vs);
    for (int c = lane; c < C; c += 32) {
        const float s1 = 1.f + sc[c];
        const float g = gamma ? gamma[c] * s1 : s1;
        fd_st(out + c, fmaf(wr[c], g, -gs));
    }
    if (lane == 0) v[(long)b * Cout + o] = vs;
}

extern "C" int fd_ln_fold(const float* W, const float* gamma, const float* beta, const float* shift, const float* scale, int mod_stride,
                          void* Wf, float* v, int B, int Cout, int C, int dtype, cudaStream_t stream) {
    if (!W || !shift || !scale || !Wf || !v || B <= 0 || Cout <= 0 || C <= 0 || mod_stride < C) return FD_ERR_BAD_ARGUMENT;
    if ((gamma == nullptr) != (beta == nullptr)) return FD_ERR_BAD_ARGUMENT;
    dim3 grid((unsigned)fd_cdiv(Cout, 8), (unsigned)B);
    FD_DISPATCH_DTYPE(dtype, T, (ln_fold_kernel<T><<<grid, 256, 0, stream>>>(W, gamma, beta, shift, scale, mod_stride, (T*)Wf, v, Cout, C)));
    FD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// GroupNorm statistics: sums[b, g, 0:2] += (sum, sumsq).  One block = 256 threads over a slab of pixels of one
// sample; thread t owns vector column (t % VR) so its group is fixed; smem + atomics finish the reduction.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ y, float* __restrict__ sums, float* __restrict__ ws, int P,
                                                       int C, int G, int pix_per_block) {
    constexpr int VEC = fd_vec<T>::N;
    __shared__ float s_s[256], s_q[256];
    __shared__ int s_last;
    const int b = blockIdx.y;
    const int vr = C / VEC;
    const int col = threadIdx.x % vr;
    const int rstep = blockDim.x / vr;
    const int r0 = threadIdx.x / vr;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(P, p0 + pix_per_block);
    float s = 0.f, q = 0.f;
    for (int p = p0 + r0; p < p1; p += rstep) {
        float v[VEC];
        fd_ldv<T, VEC>(y + ((long)b * P + p) * C + col * VEC, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { s += v[e]; q += v[e] * v[e]; }
    }
    s_s[threadIdx.x] = s;
    s_q[threadIdx.x] = q;
    __syncthreads();
    // Block sums per group in THREAD ORDER, block partials in BLOCK ORDER by the last block to arrive: the result depends on the
    // sample's data only (no floating-point atomics), so a slice gets the same statistics in any batch and on any run.
    const int nblk = gridDim.x;
    float* part = ws + ((long)b * nblk + blockIdx.x) * 2 * G;
    if (threadIdx.x < 2 * G) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        const float* src = which ? s_q : s_s;
        const int cpg_v = vr / G;                       // vector columns per group
        float t = 0.f;
        for (int r = 0; r < rstep; ++r)
            for (int c = g * cpg_v; c < (g + 1) * cpg_v; ++c) t += src[r * vr + c];
        __stcg(part + threadIdx.x, t);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int* counters = reinterpret_cast<int*>(ws + (long)gridDim.y * nblk * 2 * G);
        s_last = atomicAdd(counters + b, 1) == nblk - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 2 * G) {
        __threadfence();
        float t = 0.f;
        for (int k = 0; k < nblk; ++k) t += __ldcg(ws + ((long)b * nblk + k) * 2 * G + threadIdx.x);
        sums[(long)b * 2 * G + threadIdx.x] = t;
    }
}

extern "C" long fd_gn_stats_ws_floats(int B, int P, int G) {
    return (B > 0 && P > 0 && G > 0) ? (long)B * fd_cdiv(P, 256) * 2 * G + B : 0;
}

extern "C" int fd_gn_stats(const void* y, float* sums, float* ws, int B, int P, int C, int G, int dtype, cudaStream_t stream) {
    if (!y || !sums || !ws || B <= 0 || P <= 0 || C <= 0 || G <= 0 || C % G) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        const int vr = C / VEC;
        if (C % VEC || (C / G) % VEC || vr > 256 || 256 % vr || G > 64) return FD_ERR_UNSUPPORTED;
        const int ppb = 256;  // pixels per block
        dim3 grid(fd_cdiv(P, ppb), B);
        gn_stats_kernel<T><<<grid, 256, 0, stream>>>((const T*)y, sums, ws, P, C, G, ppb);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

// out = silu(gn(y)) + skip     (elementwise given the statistics)
// The grid stride is a multiple of C / VEC vectors (launcher), so a thread always lands on the same channel vector: the
// per-channel affine map (GroupNorm statistics, gamma/beta and the optional scale/shift folded into  t = v * A + Bc) is
// computed once per thread; the loop body is 4 independent vectors in flight, FMA + SiLU + add per element.
template <typename T>
__global__ void __launch_bounds__(256) gn_silu_add_kernel(const T* __restrict__ y, const float* __restrict__ sums,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int ss_stride, const T* __restrict__ skip, T* __restrict__ out, int P,
                                                          int C, int G, float eps, long nvec_per_sample) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int U = 4;
    const int b = blockIdx.y;
    const int cpg = C / G;
    const float inv_n = 1.f / ((float)cpg * (float)P);
    const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    if (i0 >= nvec_per_sample) return;
    const int c0 = (int)((i0 * VEC) % C);
    float ka[VEC], kb[VEC];
    {
        const int g = c0 / cpg;                      // a vector never straddles groups (cpg % VEC == 0)
        const float mean = sums[((long)b * G + g) * 2] * inv_n;
        const float var = fmaxf(sums[((long)b * G + g) * 2 + 1] * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float a = rstd * __ldg(gamma + c0 + e), c = __ldg(beta + c0 + e) - mean * a;
            if (scale) {
                const float sc = __ldg(scale + (long)b * ss_stride + c0 + e) + 1.f;
                a *= sc;
                c = fmaf(c, sc, __ldg(shift + (long)b * ss_stride + c0 + e));
            }
            ka[e] = a; kb[e] = c;
        }
    }
    const T* yb = y + (long)b * P * C;
    const T* sb = skip ? skip + (long)b * P * C : nullptr;
    T* ob = out + (long)b * P * C;
    for (long i = i0; i < nvec_per_sample; i += U * stride) {
        uint4 rv[U], rs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long j = i + u * stride;
            const long jj = j < nvec_per_sample ? j : i;          // clamped (valid) address, result discarded
            rv[u] = *reinterpret_cast<const uint4*>(yb + jj * VEC);
            if (sb) rs[u] = *reinterpret_cast<const uint4*>(sb + jj * VEC);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long j = i + u * stride;
            float v[VEC], sk[VEC];
            fd_raw_to_f<T>(rv[u], v);
            if (sb) fd_raw_to_f<T>(rs[u], sk);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float z = fmaf(v[e], ka[e], kb[e]);
                const float t = sizeof(T) == 2 ? fd_silu16(z) : fd_silu(z);
                v[e] = sb ? t + sk[e] : t;
            }
            if (j < nvec_per_sample) *reinterpret_cast<uint4*>(ob + j * VEC) = fd_f_to_raw<T>(v);
        }
    }
}

// grid.x such that the grid stride (grid.x * 256 vectors) is a multiple of the vectors per pixel: every thread then keeps its
// channel vector for the whole loop.  ~4 vectors per thread per loop trip, a few trips per thread.
static long gn_grid_x(long nvec, int vec_per_pixel) {
    long gx = fd_cdiv(nvec, 256L * 8);
    if (gx > 148L * 32) gx = 148L * 32;
    if (gx < 1) gx = 1;
    // 256 * gx % vec_per_pixel == 0  <=>  gx % (vec_per_pixel / gcd(vec_per_pixel, 256)) == 0
    long a = vec_per_pixel, b2 = 256;
    while (b2) { const long t = a % b2; a = b2; b2 = t; }
    const long m = vec_per_pixel / a;
    gx = (gx + m - 1) / m * m;
    return gx;
}

extern "C" int fd_gn_silu_add(const void* y, const float* sums, const float* gamma, const float* beta,
                              const void* skip, void* out, int B, int P, int C, int G, float eps, int dtype,
                              cudaStream_t stream) {
    if (!y || !sums || !gamma || !beta || !out || B <= 0 || P <= 0 || C <= 0 || G <= 0 || C % G) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        if ((C / G) % VEC) return FD_ERR_UNSUPPORTED;
        const long nvec = (long)P * C / VEC;
        if ((((uintptr_t)y | (uintptr_t)skip | (uintptr_t)out) & 15)) return FD_ERR_UNSUPPORTED;
        dim3 grid((unsigned)gn_grid_x(nvec, C / VEC), B);
        gn_silu_add_kernel<T><<<grid, 256, 0, stream>>>((const T*)y, sums, gamma, beta, nullptr, nullptr, 0, (const T*)skip,
                                                        (T*)out, P, C, G, eps, nvec);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

// lucidrains Block with the time-embedding scale/shift (src/denoising_diffusion_pytorch.py:183-199):
//   out = silu( GN(y) * (scale[b] + 1) + shift[b] ) + skip        scale/shift: fp32, element (b, c) at [b*ss_stride + c]
extern "C" int fd_gn_scale_shift_silu(const void* y, const float* sums, const float* gamma, const float* beta,
                                      const float* scale, const float* shift, int ss_stride, const void* skip, void* out, int B,
                                      int P, int C, int G, float eps, int dtype, cudaStream_t stream) {
    if (!y || !sums || !gamma || !beta || !scale || !shift || !out || B <= 0 || P <= 0 || C <= 0 || G <= 0 || C % G || ss_stride < C)
        return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        if ((C / G) % VEC) return FD_ERR_UNSUPPORTED;
        const long nvec = (long)P * C / VEC;
        if ((((uintptr_t)y | (uintptr_t)skip | (uintptr_t)out) & 15)) return FD_ERR_UNSUPPORTED;
        dim3 grid((unsigned)gn_grid_x(nvec, C / VEC), B);
        gn_silu_add_kernel<T><<<grid, 256, 0, stream>>>((const T*)y, sums, gamma, beta, scale, shift, ss_stride, (const T*)skip,
                                                        (T*)out, P, C, G, eps, nvec);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// Small dense layers: one warp per output element pair (b, n); K <= a few thousand.
// ------------------------------------------------------------------------------------------------------
FD_DEVINL float fd_act(float v, int act) {
    switch (act) {
        case 1: return fd_silu(v);
        case 2: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
        case 3: return fmaxf(v, 0.f);
        default: return v;
    }
}

__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ bias,
                                                           const float* __restrict__ add, float* __restrict__ out,
                                                           int B, int K, int N, int act_in, int act_out) {
    const int lane = threadIdx.x & 31;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= (long)B * N) return;
    const int b = (int)(warp / N), n = (int)(warp % N);
    const float* xr = x + (long)b * K;
    const float* wr = W + (long)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += fd_act(xr[k], act_in) * __ldg(wr + k);
    acc = fd_warp_sum(acc);
    if (lane == 0) {
        float v = fd_act(acc + (bias ? bias[n] : 0.f), act_out);
        if (add) v += add[(long)b * N + n];
        out[(long)b * N + n] = v;
    }
}

extern "C" int fd_linear_small(const float* x, const float* W, const float* bias, const float* add, float* out,
                               int B, int K, int N, int act_in, int act_out, cudaStream_t stream) {
    if (!x || !W || !out || B <= 0 || K <= 0 || N <= 0) return FD_ERR_BAD_ARGUMENT;
    const long warps = (long)B * N;
    linear_small_kernel<<<fd_cdiv(warps, 8), 256, 0, stream>>>(x, W, bias, add, out, B, K, N, act_in, act_out);
    FD_LAUNCH_CHECK();
    return 0;
}

__global__ void time_sinusoid_kernel(const float* __restrict__ time, float* __restrict__ out, int B, int dim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dim) return;
    const int b = i / dim, j = i % dim, half = dim / 2;
    const int k = j < half ? j : j - half;
    // emb = log(10000) / (half - 1); freq = exp(-k * emb)     (src/DADiff.py:180-184)
    // torch evaluates `emb` in double, rounds it to fp32, then multiplies the fp32 arange by it.
    const float nemb = (float)(-(log(10000.0) / (double)(half - 1)));
    const float freq = expf((float)k * nemb);
    const float a = time[b] * freq;
    out[i] = j < half ? sinf(a) : cosf(a);
}

extern "C" int fd_time_sinusoid(const float* time, float* out, int B, int dim, cudaStream_t stream) {
    if (!time || !out || B <= 0 || dim < 4 || dim % 2) return FD_ERR_BAD_ARGUMENT;
    time_sinusoid_kernel<<<fd_cdiv((long)B * dim, 128), 128, 0, stream>>>(time, out, B, dim);
    FD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// 2x2 / stride-2 average pooling over channels-last data: one thread per (output pixel, 16-byte channel vector).
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) avgpool2x2_nhwc_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C,
                                                              long total) {
    constexpr int VEC = fd_vec<T>::N;
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int vr = C / VEC;
    const int cv = (int)(i % vr);
    long pix = i / vr;
    const int Wo = W / 2, Ho = H / 2;
    const int xo = (int)(pix % Wo); pix /= Wo;
    const int yo = (int)(pix % Ho);
    const long b = pix / Ho;
    const T* p00 = in + (((b * H + 2 * yo) * W + 2 * xo) * (long)C) + cv * VEC;
    float a[VEC], t[VEC];
    fd_raw_to_f<T>(*reinterpret_cast<const uint4*>(p00), a);
    fd_raw_to_f<T>(*reinterpret_cast<const uint4*>(p00 + C), t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[e] += t[e];
    fd_raw_to_f<T>(*reinterpret_cast<const uint4*>(p00 + (long)W * C), t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[e] += t[e];
    fd_raw_to_f<T>(*reinterpret_cast<const uint4*>(p00 + (long)W * C + C), t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[e] = (a[e] + t[e]) * 0.25f;
    *reinterpret_cast<uint4*>(out + (((b * Ho + yo) * Wo + xo) * (long)C) + cv * VEC) = fd_f_to_raw<T>(a);
}

extern "C" int fd_avgpool2x2_nhwc(const void* in, void* out, int B, int H, int W, int C, int dtype, cudaStream_t stream) {
    if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((H & 1) || (W & 1)) return FD_ERR_UNSUPPORTED;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        if (C % VEC || (((uintptr_t)in | (uintptr_t)out) & 15)) return FD_ERR_UNSUPPORTED;
        const long total = (long)B * (H / 2) * (W / 2) * (C / VEC);
        avgpool2x2_nhwc_kernel<T><<<(unsigned)fd_cdiv(total, 256), 256, 0, stream>>>((const T*)in, (T*)out, H, W, C, total);
    });
    FD_LAUNCH_CHECK();
    return 0;
}
