// TransposedAttention (channel attention, src/DADiff.py:252-285): depthwise 3x3 over qkv fused with the per-head
// 32x32 Gram matrix q.k^T and the squared norms, then softmax + folding of the attention matrix into the output
// projection (W_eff = W_proj . blockdiag(attn)), so the attn@v + project_out pair becomes one per-sample 1x1 GEMM.
#include "fd_common.cuh"

namespace {

constexpr int HD = 32;            // channels per head (heads = C/32, src/DADiff.py:468)
constexpr int TPH = 8, TPW = 32;  // pixel tile: 8 rows x 32 cols = 256 pixels

// grid: (C/32 heads, tiles, B).  Each block: dwconv for the head's q, k, v channels on a 8x32 pixel tile.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_qkv_gram_kernel(const T* __restrict__ qkv, const float* __restrict__ w,
                                                              T* __restrict__ v_out, float* __restrict__ gram,
                                                              float* __restrict__ qk_sq, int H, int W, int C) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int NVH = HD / VEC;                    // vectors per 32-channel segment
    constexpr int HP = (TPH + 2) * (TPW + 2);        // halo pixels
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_in = reinterpret_cast<T*>(smem_raw);                         // [HP][3*HD]
    float* s_q = reinterpret_cast<float*>(s_in + HP * 3 * HD);        // [256][HD+1]
    float* s_k = s_q + TPH * TPW * (HD + 1);                          // [256][HD+1]
    float* s_w = s_k + TPH * TPW * (HD + 1);                          // [9][3*HD]

    const int head = blockIdx.x;
    const int tiles_w = (W + TPW - 1) / TPW;
    const int ty0 = (blockIdx.y / tiles_w) * TPH, tx0 = (blockIdx.y % tiles_w) * TPW;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;

    for (int i = tid; i < 9 * 3 * HD; i += 256) {
        const int tap = i / (3 * HD), sc = i % (3 * HD);   // sc: section*32 + c
        const int ch = (sc / HD) * C + head * HD + sc % HD;
        s_w[i] = w[(long)ch * 9 + tap];
    }
    for (int i = tid; i < HP * 3 * NVH; i += 256) {
        const int pix = i / (3 * NVH), sv = i % (3 * NVH);
        const int sec = sv / NVH, vc = sv % NVH;
        const int h = ty0 + pix / (TPW + 2) - 1, ww = tx0 + pix % (TPW + 2) - 1;
        float v[VEC];
        if (h >= 0 && h < H && ww >= 0 && ww < W) {
            fd_ldv<T, VEC>(qkv + (((long)b * H + h) * W + ww) * (3 * C) + sec * C + head * HD + vc * VEC, v);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = 0.f;
        }
        fd_stv<T, VEC>(s_in + pix * 3 * HD + sec * HD + vc * VEC, v);
    }
    __syncthreads();
    // depthwise conv: item = (pixel, section, vector)
    for (int i = tid; i < TPH * TPW * 3 * NVH; i += 256) {
        const int pix = i / (3 * NVH), sv = i % (3 * NVH);
        const int sec = sv / NVH, vc = sv % NVH;
        const int py = pix / TPW, px = pix % TPW;
        const bool inside = (ty0 + py < H) && (tx0 + px < W);
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                float v[VEC];
                fd_ldv<T, VEC>(s_in + ((py + dy) * (TPW + 2) + px + dx) * 3 * HD + sec * HD + vc * VEC, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] = fmaf(v[e], s_w[(dy * 3 + dx) * 3 * HD + sec * HD + vc * VEC + e], acc[e]);
            }
        if (sec == 2) {
            if (inside) fd_stv<T, VEC>(v_out + (((long)b * H + ty0 + py) * W + tx0 + px) * C + head * HD + vc * VEC, acc);
        } else {
            float* dst = (sec == 0 ? s_q : s_k) + pix * (HD + 1) + vc * VEC;
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = inside ? acc[e] : 0.f;
        }
    }
    __syncthreads();
    // Gram: thread (i, j-quad): 32 x 8 threads, each 4 outputs (i, j0..j0+3)
    {
        const int gi = tid / 8, gj = (tid % 8) * 4;
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = 0; p < TPH * TPW; ++p) {
            const float qv = s_q[p * (HD + 1) + gi];
            const float* kp = s_k + p * (HD + 1) + gj;
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = fmaf(qv, kp[j], g[j]);
        }
        float* gg = gram + (((long)b * (C / HD) + head) * HD + gi) * HD + gj;
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(gg + j, g[j]);
    }
    if (tid < 2 * HD) {
        const float* s = (tid < HD ? s_q : s_k) + (tid % HD);
        float acc = 0.f;
        for (int p = 0; p < TPH * TPW; ++p) { const float v = s[p * (HD + 1)]; acc = fmaf(v, v, acc); }
        atomicAdd(qk_sq + ((long)b * 2 + tid / HD) * C + head * HD + tid % HD, acc);
    }
}

// grid: (heads, B); block 256.  attn in shared memory, then weff rows.
template <typename T>
__global__ void __launch_bounds__(256) attn_weff_kernel(const float* __restrict__ gram, const float* __restrict__ qk_sq,
                                                        const float* __restrict__ temperature,
                                                        const float* __restrict__ proj_w, T* __restrict__ weff, int C) {
    __shared__ float s_a[HD][HD + 1];
    const int head = blockIdx.x, b = blockIdx.y, heads = C / HD;
    const int tid = threadIdx.x;
    const float temp = temperature[head];
    for (int i = tid; i < HD * HD; i += 256) {
        const int r = i / HD, c = i % HD;
        const float qn = fmaxf(sqrtf(qk_sq[((long)b * 2 + 0) * C + head * HD + r]), 1e-12f);   // F.normalize eps (:273-274)
        const float kn = fmaxf(sqrtf(qk_sq[((long)b * 2 + 1) * C + head * HD + c]), 1e-12f);
        s_a[r][c] = gram[(((long)b * heads + head) * HD + r) * HD + c] / (qn * kn) * temp;
    }
    __syncthreads();
    if (tid < HD) {  // row softmax (:277)
        float m = -INFINITY;
        for (int c = 0; c < HD; ++c) m = fmaxf(m, s_a[tid][c]);
        float s = 0.f;
        for (int c = 0; c < HD; ++c) { const float e = expf(s_a[tid][c] - m); s_a[tid][c] = e; s += e; }
        const float inv = 1.f / s;
        for (int c = 0; c < HD; ++c) s_a[tid][c] *= inv;
    }
    __syncthreads();
    // weff[b, o, head*32 + j] = sum_i proj_w[o, head*32 + i] * attn[i][j]
    for (int i = tid; i < C * HD; i += 256) {
        const int o = i / HD, j = i % HD;
        const float* pw = proj_w + (long)o * C + head * HD;
        float acc = 0.f;
#pragma unroll 8
        for (int ii = 0; ii < HD; ++ii) acc = fmaf(__ldg(pw + ii), s_a[ii][j], acc);
        fd_st(weff + ((long)b * C + o) * C + head * HD + j, acc);
    }
}

}  // namespace

extern "C" int fd_dwconv3x3_qkv_gram(const void* qkv, const float* w, void* v, float* gram, float* qk_sq, int B, int H, int W,
                                     int C, int dtype, cudaStream_t stream) {
    if (!qkv || !w || !v || !gram || !qk_sq || B <= 0 || H <= 0 || W <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if (C % HD) return FD_ERR_UNSUPPORTED;
    dim3 grid(C / HD, fd_cdiv(H, TPH) * fd_cdiv(W, TPW), B);
    FD_DISPATCH_DTYPE(dtype, T, {
        const size_t smem = (size_t)(TPH + 2) * (TPW + 2) * 3 * HD * sizeof(T) +
                            (size_t)(2 * TPH * TPW * (HD + 1) + 9 * 3 * HD) * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(dwconv_qkv_gram_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dwconv_qkv_gram_kernel<T><<<grid, 256, smem, stream>>>((const T*)qkv, w, (T*)v, gram, qk_sq, H, W, C);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

extern "C" int fd_attn_weff(const float* gram, const float* qk_sq, const float* temperature, const float* proj_w, void* weff,
                            int B, int C, int dtype, cudaStream_t stream) {
    if (!gram || !qk_sq || !temperature || !proj_w || !weff || B <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if (C % HD) return FD_ERR_UNSUPPORTED;
    dim3 grid(C / HD, B);
    FD_DISPATCH_DTYPE(dtype, T,
                      (attn_weff_kernel<T><<<grid, 256, 0, stream>>>(gram, qk_sq, temperature, proj_w, (T*)weff, C)));
    FD_LAUNCH_CHECK();
    return 0;
}
