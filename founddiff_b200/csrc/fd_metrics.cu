// Evaluation metrics of Trainer.test on the device (src/DADiff.py:1883-1888 -> src/util.py:188-236): per-slice sum of
// squared errors (PSNR, RMSE) and the SSIM map sum (kornia semantics: 11x11 Gaussian window, sigma 1.5, reflect border,
// C1 = 0.01^2, C2 = 0.03^2, map clamped to [0, 1] before the mean), so that an evaluation loop never leaves the GPU.
// One block = a 32 x 8 output tile: both images' (32+10) x (8+10) reflect-padded patches in shared memory, separable
// window (horizontal pass into shared memory for the five moments, vertical pass in registers), block reduction, one
// atomicAdd pair per block.
#include "fd_common.cuh"

namespace {

constexpr int MT_W = 32, MT_H = 8, MT_R = 5, MT_K = 2 * MT_R + 1;
constexpr int MT_PW = MT_W + 2 * MT_R, MT_PH = MT_H + 2 * MT_R;

struct GaussWin {
    float w[MT_K];
};

FD_DEVINL int reflect(int i, int n) {          // torch 'reflect' padding (no edge repeat); n > MT_R
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    // partial tiles load the full halo patch: far outside the image one fold is not enough (W = 16: x up to 36 -> -6).  Those
    // patch entries only feed masked outputs; the clamp keeps the read inside the buffer.
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(MT_W * MT_H) slice_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                                      float* __restrict__ out, int H, int W, GaussWin gw, float C1,
                                                                      float C2) {
    __shared__ float s_p[MT_PH][MT_PW], s_t[MT_PH][MT_PW];
    __shared__ float s_h[5][MT_PH][MT_W];
    __shared__ float s_red[2][MT_W * MT_H / 32];
    const int b = blockIdx.z, x0 = blockIdx.x * MT_W, y0 = blockIdx.y * MT_H;
    const int tx = threadIdx.x % MT_W, ty = threadIdx.x / MT_W;
    const float* pb = pred + (long)b * H * W;
    const float* tb = target + (long)b * H * W;
    for (int i = threadIdx.x; i < MT_PH * MT_PW; i += MT_W * MT_H) {
        const int py = i / MT_PW, px = i % MT_PW;
        const int yy = reflect(y0 + py - MT_R, H), xx = reflect(x0 + px - MT_R, W);
        s_p[py][px] = pb[(long)yy * W + xx];
        s_t[py][px] = tb[(long)yy * W + xx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MT_PH * MT_W; i += MT_W * MT_H) {       // horizontal pass: five moments per patch row
        const int py = i / MT_W, px = i % MT_W;
        float m1 = 0.f, m2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
        for (int k = 0; k < MT_K; ++k) {
            const float p = s_p[py][px + k], t = s_t[py][px + k], w = gw.w[k];
            m1 = fmaf(w, p, m1); m2 = fmaf(w, t, m2);
            m11 = fmaf(w, p * p, m11); m22 = fmaf(w, t * t, m22); m12 = fmaf(w, p * t, m12);
        }
        s_h[0][py][px] = m1; s_h[1][py][px] = m2; s_h[2][py][px] = m11; s_h[3][py][px] = m22; s_h[4][py][px] = m12;
    }
    __syncthreads();
    float sse = 0.f, ssim = 0.f;
    const int x = x0 + tx, y = y0 + ty;
    if (x < W && y < H) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < MT_K; ++k) {
            const float w = gw.w[k];
            mu1 = fmaf(w, s_h[0][ty + k][tx], mu1); mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
            e11 = fmaf(w, s_h[2][ty + k][tx], e11); e22 = fmaf(w, s_h[3][ty + k][tx], e22);
            e12 = fmaf(w, s_h[4][ty + k][tx], e12);
        }
        const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1s, s2 = e22 - mu2s, s12 = e12 - mu12;
        const float v = ((2.f * mu12 + C1) * (2.f * s12 + C2)) / ((mu1s + mu2s + C1) * (s1 + s2 + C2));
        ssim = fminf(fmaxf(v, 0.f), 1.f);
        const float d = s_p[ty + MT_R][tx + MT_R] - s_t[ty + MT_R][tx + MT_R];
        sse = d * d;
    }
    sse = fd_warp_sum(sse);
    ssim = fd_warp_sum(ssim);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_red[0][warp] = sse; s_red[1][warp] = ssim; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float a = 0.f;
        for (int i = 0; i < MT_W * MT_H / 32; ++i) a += s_red[threadIdx.x][i];
        atomicAdd(out + 2 * b + threadIdx.x, a);
    }
}

}  // namespace

extern "C" int fd_slice_metrics(const float* pred, const float* target, float* out, int B, int H, int W, float max_val,
                                cudaStream_t stream) {
    if (!pred || !target || !out || B <= 0 || H <= MT_R || W <= MT_R) return FD_ERR_BAD_ARGUMENT;
    GaussWin gw;
    float sum = 0.f;
    for (int i = 0; i < MT_K; ++i) {           // kornia get_gaussian_kernel1d(11, 1.5), float32
        const float x = (float)(i - MT_R);
        gw.w[i] = expf(-(x * x) / (2.f * 1.5f * 1.5f));
        sum += gw.w[i];
    }
    for (int i = 0; i < MT_K; ++i) gw.w[i] /= sum;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * 2 * B, stream);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(fd_cdiv(W, MT_W), fd_cdiv(H, MT_H), B);
    const float C1 = (0.01f * max_val) * (0.01f * max_val), C2 = (0.03f * max_val) * (0.03f * max_val);
    slice_metrics_kernel<<<grid, MT_W * MT_H, 0, stream>>>(pred, target, out, H, W, gw, C1, C2);
    FD_LAUNCH_CHECK();
    return 0;
}
