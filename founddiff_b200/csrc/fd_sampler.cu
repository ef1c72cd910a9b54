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

// Secondary path: epsilon-prediction GaussianDiffusion update (src/denoising_diffusion_pytorch.py:556-576 model_predictions,
// :547-554 q_posterior, :588-595 p_sample, :612-646 ddim_sample) as one fused elementwise kernel:
//   x0     = sr * x_t - srm1 * eps            (predict_start_from_noise);  clipped to [-1, 1] when coef[6] != 0
//   x_next = a0 * x0 + a1 * x_t + a2 * eps + a3 * noise
// coef (DEVICE fp32[8]) = {sr, srm1, a0, a1, a2, a3, clip, 0}:  ancestral {.., c1[t], c2[t], 0, exp(.5 logvar[t]) | 0};
// DDIM {.., sqrt(abar_next), 0, c, sigma};  last DDIM pair {.., 1, 0, 0, 0}.
__global__ void ddpm_update_kernel(const float* __restrict__ x_t, const float* __restrict__ eps, const float* __restrict__ noise,
                                   const float* __restrict__ coef, float* __restrict__ x_next, float* __restrict__ x_start, long n) {
    const float sr = coef[0], srm1 = coef[1], a0 = coef[2], a1 = coef[3], a2 = coef[4], a3 = coef[5];
    const bool clip = coef[6] != 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float xt = x_t[i], e = eps[i];
        float x0 = sr * xt - srm1 * e;
        if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        if (x_start) x_start[i] = x0;
        float xn = a0 * x0 + a1 * xt + a2 * e;
        if (noise) xn += a3 * noise[i];
        x_next[i] = xn;
    }
}

extern "C" int fd_ddpm_update(const float* x_t, const float* eps, const float* noise, const float* coef, float* x_next,
                              float* x_start, long n, cudaStream_t stream) {
    if (!x_t || !eps || !coef || !x_next || n <= 0) return FD_ERR_BAD_ARGUMENT;
    ddpm_update_kernel<<<(unsigned)min((long)fd_cdiv(n, 256), 148L * 8), 256, 0, stream>>>(x_t, eps, noise, coef, x_next, x_start, n);
    FD_LAUNCH_CHECK();
    return 0;
}

// One pixel per LPP lanes; each lane reads one 16-byte vector of the C-channel feature row.
// MODE selects the objective branch of model_predictions (src/DADiff.py:1168-1207); feat1 / w1 / bias1 are the second
// Unet's final_conv operands (num_unet = 2, src/DADiff.py:817-820) and are only read by the two-output modes.
//   FD_OBJ_PRED_RES       o0 = pred_res                    (:1202-1207, also 'pred_res_noise' with test_res_or_noise = "res")
//   FD_OBJ_PRED_NOISE     o0 = pred_noise                  (:1194-1201, also test_res_or_noise = "noise" :1180-1187)
//   FD_OBJ_PRED_RES_NOISE o0 = pred_res, o1 = pred_noise   (:1169-1175)
//   FD_OBJ_PRED_X0_NOISE  o0 = x_start,  o1 = pred_noise   (:1188-1192)
// coef (DEVICE fp32[8]) = {c_xt, c_res, c_x0, c_noise, alphas_cumsum[t], betas_cumsum[t], one_minus_alphas_cumsum[t], 0}
template <typename T, int MODE>
__global__ void __launch_bounds__(256) final_conv_update_kernel(
    const T* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
    const T* __restrict__ feat1, const float* __restrict__ w1, const float* __restrict__ bias1,
    const float* __restrict__ x_input, const float* __restrict__ x_t, const float* __restrict__ noise,
    const float* __restrict__ coef, float* __restrict__ x_next, float* __restrict__ pred_res,
    float* __restrict__ pred_noise, float* __restrict__ x_start, long npix, int C) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr bool TWO = (MODE == FD_OBJ_PRED_RES_NOISE || MODE == FD_OBJ_PRED_X0_NOISE);
    constexpr int U = TWO ? 2 : 4;   // pixel groups per warp with all their loads in flight at once (one group per warp left
                                     // the kernel latency-bound at 2.5 TB/s: a single 16-byte load, then a dependent scalar tail)
    const int lpp = C / VEC;  // lanes per pixel (power of two <= 32)
    const int lane = threadIdx.x & 31;
    const int sub = lane % lpp;
    const int ppw = 32 / lpp;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long pix0 = warp * (ppw * U) + lane / lpp;
    float wv[VEC], wv1[TWO ? VEC : 1];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        wv[e] = __ldg(w + sub * VEC + e);
        if (TWO) wv1[e] = __ldg(w1 + sub * VEC + e);
    }
    float v[U][VEC], v1[TWO ? U : 1][VEC], xi[U], xt[U], nz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long pix = pix0 + (long)u * ppw;
        const long pc = pix < npix ? pix : npix - 1;             // clamped: loads stay unconditional
        fd_ldv<T, VEC>(feat + pc * (long)C + sub * VEC, v[u]);
        if (TWO) fd_ldv<T, VEC>(feat1 + pc * (long)C + sub * VEC, v1[u]);
        xi[u] = x_input[pc];
        xt[u] = x_t[pc];
        nz[u] = noise ? noise[pc] : 0.f;
    }
    const float c_xt = coef[0], c_res = coef[1], c_x0 = coef[2], c_noise = coef[3], acs = coef[4], bcs = coef[5];
    const float omacs = coef[6], b0 = bias[0], b1 = TWO ? bias1[0] : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long pix = pix0 + (long)u * ppw;
        float acc = 0.f, acc1 = 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            acc = fmaf(v[u][e], wv[e], acc);
            if (TWO) acc1 = fmaf(v1[u][e], wv1[e], acc1);
        }
        for (int o = lpp / 2; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (TWO) acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
        }
        if (pix >= npix || sub != 0) continue;
        const float o0 = acc + b0;
        float pr, pn, x0;
        if (MODE == FD_OBJ_PRED_RES) {
            pr = fminf(fmaxf(o0, -1.f), 1.f);                         // :1165-1166, 1204
            x0 = fminf(fmaxf(xi[u] - pr, -1.f), 1.f);                 // :1206-1207
            pn = (xt[u] - xi[u] - (acs - 1.f) * pr) / bcs;            // :1120-1124
        } else if (MODE == FD_OBJ_PRED_NOISE) {
            pn = o0;
            x0 = (xt[u] - acs * xi[u] - bcs * pn) / omacs;            // :1126-1130
            x0 = fminf(fmaxf(x0, -1.f), 1.f);
            pr = fminf(fmaxf(xi[u] - x0, -1.f), 1.f);                 // :1199-1200
        } else if (MODE == FD_OBJ_PRED_RES_NOISE) {
            pr = fminf(fmaxf(o0, -1.f), 1.f);
            pn = acc1 + b1;
            x0 = fminf(fmaxf(xt[u] - acs * pr - bcs * pn, -1.f), 1.f);   // :1132-1136, 1175
        } else {
            pr = fminf(fmaxf(xi[u] - o0, -1.f), 1.f);                 // :1189, 1191
            pn = acc1 + b1;
            x0 = fminf(fmaxf(o0, -1.f), 1.f);                         // :1192
        }
        if (pred_res) pred_res[pix] = pr;
        if (x_start) x_start[pix] = x0;
        if (pred_noise) pred_noise[pix] = pn;
        x_next[pix] = c_xt * xt[u] + c_res * pr + c_x0 * x0 + c_noise * nz[u];
    }
}

extern "C" int fd_final_conv_update_obj(const void* feat, const float* w, const float* bias, const void* feat1,
                                        const float* w1, const float* bias1, const float* x_input, const float* x_t,
                                        const float* noise, const float* coef, float* x_next, float* pred_res,
                                        float* pred_noise, float* x_start, long npix, int C, int dtype, int objective,
                                        cudaStream_t stream) {
    if (!feat || !w || !bias || !x_input || !x_t || !coef || !x_next || npix <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if (objective < FD_OBJ_PRED_RES || objective > FD_OBJ_PRED_X0_NOISE) return FD_ERR_BAD_ARGUMENT;
    const bool two = objective == FD_OBJ_PRED_RES_NOISE || objective == FD_OBJ_PRED_X0_NOISE;
    if (two && (!feat1 || !w1 || !bias1)) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        const int lpp = C / VEC;
        if (C % VEC || lpp > 32 || (lpp & (lpp - 1))) return FD_ERR_UNSUPPORTED;
        const bool two_k = objective == FD_OBJ_PRED_RES_NOISE || objective == FD_OBJ_PRED_X0_NOISE;
        const long ppw_u = (long)(32 / lpp) * (two_k ? 2 : 4);       // pixels per warp (U groups, see the kernel)
        const long warps = (npix + ppw_u - 1) / ppw_u;
        const unsigned grid = (unsigned)fd_cdiv(warps, 8);
        const T *f0 = (const T*)feat, *f1 = (const T*)feat1;
        switch (objective) {
            case FD_OBJ_PRED_RES:
                final_conv_update_kernel<T, FD_OBJ_PRED_RES><<<grid, 256, 0, stream>>>(
                    f0, w, bias, f1, w1, bias1, x_input, x_t, noise, coef, x_next, pred_res, pred_noise, x_start, npix, C);
                break;
            case FD_OBJ_PRED_NOISE:
                final_conv_update_kernel<T, FD_OBJ_PRED_NOISE><<<grid, 256, 0, stream>>>(
                    f0, w, bias, f1, w1, bias1, x_input, x_t, noise, coef, x_next, pred_res, pred_noise, x_start, npix, C);
                break;
            case FD_OBJ_PRED_RES_NOISE:
                final_conv_update_kernel<T, FD_OBJ_PRED_RES_NOISE><<<grid, 256, 0, stream>>>(
                    f0, w, bias, f1, w1, bias1, x_input, x_t, noise, coef, x_next, pred_res, pred_noise, x_start, npix, C);
                break;
            default:
                final_conv_update_kernel<T, FD_OBJ_PRED_X0_NOISE><<<grid, 256, 0, stream>>>(
                    f0, w, bias, f1, w1, bias1, x_input, x_t, noise, coef, x_next, pred_res, pred_noise, x_start, npix, C);
                break;
        }
    });
    FD_LAUNCH_CHECK();
    return 0;
}

extern "C" int fd_final_conv_update(const void* feat, const float* w, const float* bias, const float* x_input,
                                    const float* x_t, const float* noise, const float* coef, float* x_next,
                                    float* pred_res, float* pred_noise, float* x_start, long npix, int C, int dtype,
                                    cudaStream_t stream) {
    return fd_final_conv_update_obj(feat, w, bias, nullptr, nullptr, nullptr, x_input, x_t, noise, coef, x_next,
                                    pred_res, pred_noise, x_start, npix, C, dtype, FD_OBJ_PRED_RES, stream);
}

// ------------------------------------------------------------------------------------------------------
// init_conv 7x7 pad 3 over two fp32 single-channel images (Unet.init_conv, src/DADiff.py:558, 700 + cat :1160).
// fp32 CUDA-core kernel (the two input images are fp32 and K = 98 is tiny): block = 16 rows x 64 cols of output pixels,
// 256 threads; a thread owns 4 horizontally adjacent pixels x 16 output channels per pass, so every weight vector read
// from shared memory feeds 4 pixels and every input value feeds 16 channels (64 FMAs per 4 + 10 shared loads).
// ------------------------------------------------------------------------------------------------------
constexpr int IC_TH = 16, IC_TW = 64;
constexpr int IC_PH = IC_TH + 6, IC_PW = IC_TW + 6 + 2;     // patch with halo (+2: row pitch 72 keeps 16-byte alignment)

template <typename T>
__global__ void __launch_bounds__(256) init_conv7x7_kernel(const float* __restrict__ x_t, const float* __restrict__ x_in,
                                                           const float* __restrict__ weight,
                                                           const float* __restrict__ bias, T* __restrict__ out, int H,
                                                           int W, int Cout) {
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                                  // [98][Cout]  tap-major
    float* sb = sw + 98 * Cout;                      // [Cout]
    float* sp = sb + Cout;                           // [2][IC_PH][IC_PW]
    const int b = blockIdx.z, ty0 = blockIdx.y * IC_TH, tx0 = blockIdx.x * IC_TW;
    for (int i = threadIdx.x; i < 98 * Cout; i += 256) {
        const int co = i / 98, tap = i % 98;         // weight memory order: (co, ci, kh, kw)
        sw[tap * Cout + co] = weight[i];
    }
    for (int i = threadIdx.x; i < Cout; i += 256) sb[i] = bias[i];
    const long img = (long)b * H * W;
    for (int i = threadIdx.x; i < 2 * IC_PH * IC_PW; i += 256) {
        const int ci = i / (IC_PH * IC_PW), r = (i % (IC_PH * IC_PW)) / IC_PW, c = i % IC_PW;
        const int h = ty0 + r - 3, w = tx0 + c - 3;
        float v = 0.f;
        if (h >= 0 && h < H && w >= 0 && w < W) v = (ci == 0 ? x_t : x_in)[img + (long)h * W + w];
        sp[i] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / 16, tx4 = (threadIdx.x % 16) * 4;      // 16 rows x 16 groups of 4 pixels
    const int h = ty0 + ty;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
        float acc[4][16];
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[px][j] = sb[c0 + j];
        for (int ci = 0; ci < 2; ++ci)
            for (int kh = 0; kh < 7; ++kh) {
                const float* prow = sp + (ci * IC_PH + ty + kh) * IC_PW + tx4;
                float in[10];
#pragma unroll
                for (int j = 0; j < 10; ++j) in[j] = prow[j];
#pragma unroll
                for (int kw = 0; kw < 7; ++kw) {
                    const float4* wp = reinterpret_cast<const float4*>(sw + ((ci * 7 + kh) * 7 + kw) * Cout + c0);
                    const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                    const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                    for (int px = 0; px < 4; ++px)
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[px][j] = fmaf(in[px + kw], wv[j], acc[px][j]);
                }
            }
        if (h < H) {
            constexpr int VEC = fd_vec<T>::N;
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int w = tx0 + tx4 + px;
                if (w >= W) continue;
                T* orow = out + (img + (long)h * W + w) * Cout + c0;
#pragma unroll
                for (int j = 0; j < 16; j += VEC) {
                    float t[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) t[e] = acc[px][j + e];
                    fd_stv<T, VEC>(orow + j, t);
                }
            }
        }
    }
}

extern "C" int fd_init_conv7x7(const float* x_t, const float* x_input, const float* weight, const float* bias,
                               void* out, int B, int H, int W, int Cout, int dtype, cudaStream_t stream) {
    if (!x_t || !x_input || !weight || !bias || !out || B <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout % 16)
        return FD_ERR_BAD_ARGUMENT;
    const size_t smem = (size_t)(98 * Cout + Cout + 2 * IC_PH * IC_PW) * sizeof(float);
    if (smem > 200 * 1024) return FD_ERR_UNSUPPORTED;
    dim3 grid(fd_cdiv(W, IC_TW), fd_cdiv(H, IC_TH), B);
    FD_DISPATCH_DTYPE(dtype, T, {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(init_conv7x7_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        init_conv7x7_kernel<T><<<grid, 256, smem, stream>>>(x_t, x_input, weight, bias, (T*)out, H, W, Cout);
    });
    FD_LAUNCH_CHECK();
    return 0;
}
