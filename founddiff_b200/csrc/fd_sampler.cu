// Sampler-side kernels: noise injection, init_conv 7x7 over cat(x_t, x_input), and the fused
// final_conv + model_predictions + posterior / DDIM update (src/DADiff.py:1153-1209, 1221-1230, 1317-1344).
#include "fd_common.cuh"

__global__ void sampler_init_kernel(const float* __restrict__ ldct, const float* __restrict__ noise, float noise_scale,
                                    float* __restrict__ x_input, float* __restrict__ x_t, float* __restrict__ first,
                                    long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float xi = ldct[i] * 2.f - 1.f;              // normalize_to_neg_one_to_one (:106-110)
        const float xt = xi + noise_scale * noise[i];      // :1294-1295
        x_input[i] = xi;
        x_t[i] = xt;
        if (first) first[i] = (xt + 1.f) * 0.5f;           // input_add_noise, un-normalised (:1296, 1359)
    }
}

extern "C" int fd_sampler_init(const float* ldct, const float* noise, float noise_scale, float* x_input, float* x_t,
                               float* first, long n, cudaStream_t stream) {
    if (!ldct || !noise || !x_input || !x_t || n <= 0) return FD_ERR_BAD_ARGUMENT;
    sampler_init_kernel<<<(unsigned)min((long)fd_cdiv(n, 256), 148L * 8), 256, 0, stream>>>(ldct, noise, noise_scale,
                                                                                          x_input, x_t, first, n);
    FD_LAUNCH_CHECK();
    return 0;
}

__global__ void unnormalize_kernel(const float* __restrict__ x, float* __restrict__ out, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = (x[i] + 1.f) * 0.5f;
}

extern "C" int fd_unnormalize(const float* x, float* out, long n, cudaStream_t stream) {
    if (!x || !out || n <= 0) return FD_ERR_BAD_ARGUMENT;
    unnormalize_kernel<<<(unsigned)min((long)fd_cdiv(n, 256), 148L * 8), 256, 0, stream>>>(x, out, n);
    FD_LAUNCH_CHECK();
    return 0;
}

// One pixel per LPP lanes; each lane reads one 16-byte vector of the C-channel feature row.
template <typename T>
__global__ void __launch_bounds__(256) final_conv_update_kernel(
    const T* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ x_input, const float* __restrict__ x_t, const float* __restrict__ noise,
    const float* __restrict__ coef, float* __restrict__ x_next, float* __restrict__ pred_res,
    float* __restrict__ pred_noise, float* __restrict__ x_start, long npix, int C) {
    constexpr int VEC = fd_vec<T>::N;
    const int lpp = C / VEC;  // lanes per pixel (power of two <= 32)
    const int lane = threadIdx.x & 31;
    const int sub = lane % lpp;
    const int ppw = 32 / lpp;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long pix = warp * ppw + lane / lpp;
    const bool active = pix < npix;
    float v[VEC];
    float acc = 0.f;
    if (active) {
        fd_ldv<T, VEC>(feat + pix * (long)C + sub * VEC, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc += v[e] * __ldg(w + sub * VEC + e);
    }
    for (int o = lpp / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (!active || sub != 0) return;
    const float c_xt = coef[0], c_res = coef[1], c_x0 = coef[2], c_noise = coef[3], acs = coef[4], bcs = coef[5];
    const float xi = x_input[pix], xt = x_t[pix];
    const float pr = fminf(fmaxf(acc + bias[0], -1.f), 1.f);      // :1165-1166, 1204
    const float x0 = fminf(fmaxf(xi - pr, -1.f), 1.f);            // :1206-1207
    if (pred_res) pred_res[pix] = pr;
    if (x_start) x_start[pix] = x0;
    if (pred_noise) pred_noise[pix] = (xt - xi - (acs - 1.f) * pr) / bcs;   // :1120-1124
    float xn = c_xt * xt + c_res * pr + c_x0 * x0;
    if (noise) xn += c_noise * noise[pix];
    x_next[pix] = xn;
}

extern "C" int fd_final_conv_update(const void* feat, const float* w, const float* bias, const float* x_input,
                                    const float* x_t, const float* noise, const float* coef, float* x_next,
                                    float* pred_res, float* pred_noise, float* x_start, long npix, int C, int dtype,
                                    cudaStream_t stream) {
    if (!feat || !w || !bias || !x_input || !x_t || !coef || !x_next || npix <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        const int lpp = C / VEC;
        if (C % VEC || lpp > 32 || (lpp & (lpp - 1))) return FD_ERR_UNSUPPORTED;
        const long warps = (npix + (32 / lpp) - 1) / (32 / lpp);
        final_conv_update_kernel<T><<<fd_cdiv(warps, 8), 256, 0, stream>>>((const T*)feat, w, bias, x_input, x_t, noise,
                                                                          coef, x_next, pred_res, pred_noise, x_start,
                                                                          npix, C);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// init_conv 7x7 pad 3 over two fp32 single-channel images (Unet.init_conv, src/DADiff.py:558, 700 + cat :1160).
// Block: 16x16 output pixels; the 22x22x2 input patch and the (Cout,2,7,7) weights live in shared memory;
// each thread produces all Cout channels of one pixel, 16 channels per pass.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) init_conv7x7_kernel(const float* __restrict__ x_t, const float* __restrict__ x_in,
                                                           const float* __restrict__ weight,
                                                           const float* __restrict__ bias, T* __restrict__ out, int H,
                                                           int W, int Cout) {
    extern __shared__ float sm[];
    float* sw = sm;                      // [98][Cout]  (tap-major so that a pass reads 16 consecutive floats)
    float* sb = sw + 98 * Cout;          // [Cout]
    float* sp = sb + Cout;               // [2][22][23]
    const int b = blockIdx.z, ty0 = blockIdx.y * 16, tx0 = blockIdx.x * 16;
    for (int i = threadIdx.x; i < 98 * Cout; i += 256) {
        const int co = i / 98, tap = i % 98;             // weight memory order: (co, ci, kh, kw)
        sw[tap * Cout + co] = weight[i];
    }
    for (int i = threadIdx.x; i < Cout; i += 256) sb[i] = bias[i];
    const long img = (long)b * H * W;
    for (int i = threadIdx.x; i < 2 * 22 * 22; i += 256) {
        const int ci = i / 484, r = (i % 484) / 22, c = i % 22;
        const int h = ty0 + r - 3, w = tx0 + c - 3;
        float v = 0.f;
        if (h >= 0 && h < H && w >= 0 && w < W) v = (ci == 0 ? x_t : x_in)[img + (long)h * W + w];
        sp[(ci * 22 + r) * 23 + c] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    const int h = ty0 + ty, w = tx0 + tx;
    if (h >= H || w >= W) return;
    T* orow = out + (img + (long)h * W + w) * Cout;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = sb[c0 + j];
        for (int ci = 0; ci < 2; ++ci)
            for (int kh = 0; kh < 7; ++kh)
#pragma unroll
                for (int kw = 0; kw < 7; ++kw) {
                    const float v = sp[(ci * 22 + ty + kh) * 23 + tx + kw];
                    const float* wp = sw + ((ci * 7 + kh) * 7 + kw) * Cout + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
                }
        constexpr int VEC = fd_vec<T>::N;
#pragma unroll
        for (int j = 0; j < 16; j += VEC) {
            float t[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) t[e] = acc[j + e];
            fd_stv<T, VEC>(orow + c0 + j, t);
        }
    }
}

extern "C" int fd_init_conv7x7(const float* x_t, const float* x_input, const float* weight, const float* bias,
                               void* out, int B, int H, int W, int Cout, int dtype, cudaStream_t stream) {
    if (!x_t || !x_input || !weight || !bias || !out || B <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout % 16)
        return FD_ERR_BAD_ARGUMENT;
    const size_t smem = (size_t)(98 * Cout + Cout + 2 * 22 * 23) * sizeof(float);
    if (smem > 200 * 1024) return FD_ERR_UNSUPPORTED;
    dim3 grid(fd_cdiv(W, 16), fd_cdiv(H, 16), B);
    FD_DISPATCH_DTYPE(dtype, T, {
        cudaError_t e = cudaFuncSetAttribute(init_conv7x7_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        init_conv7x7_kernel<T><<<grid, 256, smem, stream>>>(x_t, x_input, weight, bias, (T*)out, H, W, Cout);
    });
    FD_LAUNCH_CHECK();
    return 0;
}
