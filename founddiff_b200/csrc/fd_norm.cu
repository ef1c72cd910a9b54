// LayerNorm+modulate, GroupNorm statistics / apply, small dense layers, sinusoidal embedding.
// All HBM-bound row kernels: 16-byte vector accesses, sub-warp shuffles for the per-row reductions.
#include "fd_common.cuh"

// ------------------------------------------------------------------------------------------------------
// LayerNorm over C + adaLN modulate (src/DADiff.py:450-451, 459, 461, 486-487).
// LPR lanes cooperate on one row; each lane owns NV 16-byte vectors (interleaved: vector j of lane i is
// vector i + j*LPR of the row) so that a warp's loads are fully coalesced.
// ------------------------------------------------------------------------------------------------------
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ scale, int mod_stride,
                                                          long rows, int P, int C, float eps) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int RPW = 32 / LPR;  // rows per warp
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long row = warp * RPW + lane / LPR;
    const bool active = row < rows;
    float v[NV][VEC];
    float s = 0.f;
    const T* xr = x + (active ? row : 0) * (long)C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        fd_ldv<T, VEC>(xr + (sub + j * LPR) * VEC, v[j]);
#pragma unroll
        for (int e = 0; e < VEC; ++e) s += v[j][e];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) { float d = v[j][e] - mean; q += d * d; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (!active) return;
    const int b = (int)(row / P);
    const float* sh = shift + (long)b * mod_stride;
    const float* sc = scale + (long)b * mod_stride;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c0 = (sub + j * LPR) * VEC;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float y = (v[j][e] - mean) * rstd;
            if (gamma) y = y * __ldg(gamma + c0 + e) + __ldg(beta + c0 + e);
            v[j][e] = y * (1.f + __ldg(sc + c0 + e)) + __ldg(sh + c0 + e);
        }
        fd_stv<T, VEC>(out + row * (long)C + c0, v[j]);
    }
}

template <typename T>
static int ln_modulate_launch(const void* x, void* out, const float* gamma, const float* beta,
                              const float* shift, const float* scale, int mod_stride, int B, int P, int C,
                              float eps, cudaStream_t st) {
    constexpr int VEC = fd_vec<T>::N;
    if (C % VEC) return FD_ERR_UNSUPPORTED;
    const int vr = C / VEC;  // vectors per row
    const long rows = (long)B * P;
    const int lpr = vr >= 32 ? 32 : vr;
    const int nv = vr / lpr;
    if (lpr * nv != vr || (lpr & (lpr - 1))) return FD_ERR_UNSUPPORTED;
    const int rpw = 32 / lpr;
    const int warps = 8;
    const int grid = fd_cdiv(rows, (long)rpw * warps);
#define LN_CASE(L, N)                                                                                        \
    if (lpr == L && nv == N) {                                                                               \
        ln_modulate_kernel<T, L, N><<<grid, warps * 32, 0, st>>>((const T*)x, (T*)out, gamma, beta, shift,   \
                                                                 scale, mod_stride, rows, P, C, eps);        \
        FD_LAUNCH_CHECK();                                                                                   \
        return 0;                                                                                            \
    }
    LN_CASE(4, 1) LN_CASE(8, 1) LN_CASE(16, 1) LN_CASE(32, 1) LN_CASE(32, 2) LN_CASE(32, 4) LN_CASE(32, 8)
#undef LN_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_ln_modulate(const void* x, void* out, const float* gamma, const float* beta,
                              const float* shift, const float* scale, int mod_stride, int B, int P, int C,
                              float eps, int dtype, cudaStream_t stream) {
    if (!x || !out || !shift || !scale || B <= 0 || P <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((gamma == nullptr) != (beta == nullptr)) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T,
                      return ln_modulate_launch<T>(x, out, gamma, beta, shift, scale, mod_stride, B, P, C, eps, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// SS2D tail on channels-last data (src/emamba2.py:365, 747-748): out = (LN(y) * gamma + beta) * z + local[b]
// with z = columns [z_off, z_off + C) of rows of pitch `ld` (the silu(z) half of xz).  Same row mapping as above.
// ------------------------------------------------------------------------------------------------------
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(256) ln_gate_kernel(const T* __restrict__ y, const T* __restrict__ xz, int ld, int z_off,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ local, T* __restrict__ out, long rows, int P,
                                                      int C, float eps) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long row = warp * RPW + lane / LPR;
    const bool active = row < rows;
    float v[NV][VEC];
    float s = 0.f;
    const T* yr = y + (active ? row : 0) * (long)C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        fd_ldv<T, VEC>(yr + (sub + j * LPR) * VEC, v[j]);
#pragma unroll
        for (int e = 0; e < VEC; ++e) s += v[j][e];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) { float d = v[j][e] - mean; q += d * d; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (!active) return;
    const int b = (int)(row / P);
    const T* zr = xz + row * (long)ld + z_off;
    const float* lc = local + (long)b * C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c0 = (sub + j * LPR) * VEC;
        float z[VEC];
        fd_ldv<T, VEC>(zr + c0, z);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float n = (v[j][e] - mean) * rstd * __ldg(gamma + c0 + e) + __ldg(beta + c0 + e);
            v[j][e] = n * z[e] + __ldg(lc + c0 + e);
        }
        fd_stv<T, VEC>(out + row * (long)C + c0, v[j]);
    }
}

template <typename T>
static int ln_gate_launch(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta,
                          const float* local, void* out, int B, int P, int C, float eps, cudaStream_t st) {
    constexpr int VEC = fd_vec<T>::N;
    if (C % VEC || ld % VEC || z_off % VEC) return FD_ERR_UNSUPPORTED;
    const int vr = C / VEC;
    const long rows = (long)B * P;
    const int lpr = vr >= 32 ? 32 : vr;
    const int nv = vr / lpr;
    if (lpr * nv != vr || (lpr & (lpr - 1))) return FD_ERR_UNSUPPORTED;
    const int rpw = 32 / lpr, warps = 8;
    const int grid = fd_cdiv(rows, (long)rpw * warps);
#define LG_CASE(L, N)                                                                                              \
    if (lpr == L && nv == N) {                                                                                     \
        ln_gate_kernel<T, L, N><<<grid, warps * 32, 0, st>>>((const T*)y, (const T*)xz, ld, z_off, gamma, beta, local, \
                                                             (T*)out, rows, P, C, eps);                            \
        FD_LAUNCH_CHECK();                                                                                         \
        return 0;                                                                                                  \
    }
    LG_CASE(2, 1) LG_CASE(4, 1) LG_CASE(8, 1) LG_CASE(16, 1) LG_CASE(32, 1) LG_CASE(32, 2) LG_CASE(32, 4) LG_CASE(32, 8)
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
// GroupNorm statistics: sums[b, g, 0:2] += (sum, sumsq).  One block = 256 threads over a slab of pixels of one
// sample; thread t owns vector column (t % VR) so its group is fixed; smem + atomics finish the reduction.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ y, float* __restrict__ sums, int P,
                                                       int C, int G, int pix_per_block) {
    constexpr int VEC = fd_vec<T>::N;
    extern __shared__ float sm[];  // [G][2]
    const int b = blockIdx.y;
    const int vr = C / VEC;
    const int col = threadIdx.x % vr;
    const int rstep = blockDim.x / vr;
    const int r0 = threadIdx.x / vr;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(P, p0 + pix_per_block);
    float s = 0.f, q = 0.f;
    for (int p = p0 + r0; p < p1; p += rstep) {
        float v[VEC];
        fd_ldv<T, VEC>(y + ((long)b * P + p) * C + col * VEC, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { s += v[e]; q += v[e] * v[e]; }
    }
    const int g = (col * VEC) / (C / G);
    atomicAdd(&sm[2 * g], s);
    atomicAdd(&sm[2 * g + 1], q);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&sums[(long)b * 2 * G + i], sm[i]);
}

extern "C" int fd_gn_stats(const void* y, float* sums, int B, int P, int C, int G, int dtype, cudaStream_t stream) {
    if (!y || !sums || B <= 0 || P <= 0 || C <= 0 || G <= 0 || C % G) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        const int vr = C / VEC;
        if (C % VEC || (C / G) % VEC || vr > 256 || 256 % vr) return FD_ERR_UNSUPPORTED;
        const int ppb = 256;  // pixels per block
        dim3 grid(fd_cdiv(P, ppb), B);
        gn_stats_kernel<T><<<grid, 256, 2 * G * sizeof(float), stream>>>((const T*)y, sums, P, C, G, ppb);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

// out = silu(gn(y)) + skip     (elementwise given the statistics)
template <typename T>
__global__ void __launch_bounds__(256) gn_silu_add_kernel(const T* __restrict__ y, const float* __restrict__ sums,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int ss_stride, const T* __restrict__ skip, T* __restrict__ out, int P,
                                                          int C, int G, float eps, long nvec_per_sample) {
    constexpr int VEC = fd_vec<T>::N;
    const int b = blockIdx.y;
    const int cpg = C / G;
    const float inv_n = 1.f / ((float)cpg * (float)P);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec_per_sample; i += (long)gridDim.x * blockDim.x) {
        const int c0 = (int)((i * VEC) % C);
        const int g = c0 / cpg;
        const float mean = sums[((long)b * G + g) * 2] * inv_n;
        const float var = fmaxf(sums[((long)b * G + g) * 2 + 1] * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        const long off = (long)b * P * C + i * VEC;
        float v[VEC], s[VEC];
        fd_ldv<T, VEC>(y + off, v);
        if (skip) fd_ldv<T, VEC>(skip + off, s);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float t = (v[e] - mean) * rstd * __ldg(gamma + c0 + e) + __ldg(beta + c0 + e);
            if (scale) t = t * (__ldg(scale + (long)b * ss_stride + c0 + e) + 1.f) + __ldg(shift + (long)b * ss_stride + c0 + e);
            t = fd_silu(t);
            v[e] = skip ? t + s[e] : t;
        }
        fd_stv<T, VEC>(out + off, v);
    }
}

extern "C" int fd_gn_silu_add(const void* y, const float* sums, const float* gamma, const float* beta,
                              const void* skip, void* out, int B, int P, int C, int G, float eps, int dtype,
                              cudaStream_t stream) {
    if (!y || !sums || !gamma || !beta || !out || B <= 0 || P <= 0 || C <= 0 || G <= 0 || C % G) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        if ((C / G) % VEC) return FD_ERR_UNSUPPORTED;
        const long nvec = (long)P * C / VEC;
        dim3 grid((unsigned)min((long)fd_cdiv(nvec, 256), 148L * 16), B);
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
        dim3 grid((unsigned)min((long)fd_cdiv(nvec, 256), 148L * 16), B);
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
