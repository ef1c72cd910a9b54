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

// ---------------------------------------------------------------------------------------------------------
// 16-bit storage types: same fusion, but the Gram matrix and the squared norms run on the tensor cores
// (mma.sync m16n8k16, fp32 accumulate): G = Q^T K, |q|^2 = diag(Q^T Q), |k|^2 = diag(K^T K) with q, k rounded once
// to the storage type (cosine similarities stay consistent).  A block walks TPB consecutive 8x32 tiles of one
// (sample, head) and keeps the accumulators in registers, so global atomics are issued once per block instead of
// once per tile.
constexpr int QK_LD = 40;      // padded row (elements) of the q / k tiles: conflict-free ldmatrix
constexpr int TPB = 8;         // tiles per block

template <typename T>
__global__ void __launch_bounds__(256) dwconv_qkv_gram_mma_kernel(const T* __restrict__ qkv, const float* __restrict__ w,
                                                                  T* __restrict__ v_out, float* __restrict__ gram,
                                                                  float* __restrict__ qk_sq, int H, int W, int C, int ntiles) {
    constexpr int VEC = 8, NVH = HD / VEC;
    constexpr int HP = (TPH + 2) * (TPW + 2);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_in = reinterpret_cast<T*>(smem_raw);                 // [HP][3*HD]
    T* s_q = s_in + HP * 3 * HD;                              // [256][QK_LD]
    T* s_k = s_q + TPH * TPW * QK_LD;                         // [256][QK_LD]
    float* s_w = reinterpret_cast<float*>(s_k + TPH * TPW * QK_LD);   // [9][3*HD]
    float* s_red = s_w + 9 * 3 * HD;                          // [3][32][32] cross-warp reduction

    const int head = blockIdx.x, b = blockIdx.z;
    const int tiles_w = (W + TPW - 1) / TPW;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 9 * 3 * HD; i += 256) {
        const int tap = i / (3 * HD), sc = i % (3 * HD);
        s_w[i] = w[(long)((sc / HD) * C + head * HD + sc % HD) * 9 + tap];
    }
    for (int i = tid; i < 3 * HD * HD; i += 256) s_red[i] = 0.f;

    float acc[2][4][4];        // Q^T K            [m tile][n tile][frag]
    float accd[2][2][2][4];    // Q^T Q, K^T K diagonal 16x16 blocks  [which][m tile][n sub-tile][frag]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int ns = 0; ns < 2; ++ns)
#pragma unroll
                for (int e = 0; e < 4; ++e) accd[a][mt][ns][e] = 0.f;
    }

    const int t_begin = blockIdx.y * TPB, t_end = min(ntiles, t_begin + TPB);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int ty0 = (tile / tiles_w) * TPH, tx0 = (tile % tiles_w) * TPW;
        __syncthreads();       // previous tile's mma reads of s_q/s_k and conv reads of s_in are done
        for (int i = tid; i < HP * 3 * NVH; i += 256) {
            const int pix = i / (3 * NVH), sv = i % (3 * NVH);
            const int sec = sv / NVH, vc = sv % NVH;
            const int h = ty0 + pix / (TPW + 2) - 1, ww = tx0 + pix % (TPW + 2) - 1;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (h >= 0 && h < H && ww >= 0 && ww < W)
                val = *reinterpret_cast<const uint4*>(qkv + (((long)b * H + h) * W + ww) * (3 * C) + sec * C + head * HD + vc * VEC);
            *reinterpret_cast<uint4*>(s_in + pix * 3 * HD + sec * HD + vc * VEC) = val;
        }
        __syncthreads();
        for (int i = tid; i < TPH * TPW * 3 * NVH; i += 256) {
            const int pix = i / (3 * NVH), sv = i % (3 * NVH);
            const int sec = sv / NVH, vc = sv % NVH;
            const int py = pix / TPW, px = pix % TPW;
            const bool inside = (ty0 + py < H) && (tx0 + px < W);
            float a8[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) a8[e] = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    float v[VEC];
                    fd_ldv<T, VEC>(s_in + ((py + dy) * (TPW + 2) + px + dx) * 3 * HD + sec * HD + vc * VEC, v);
                    const float* wp = s_w + (dy * 3 + dx) * 3 * HD + sec * HD + vc * VEC;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) a8[e] = fmaf(v[e], wp[e], a8[e]);
                }
            if (sec == 2) {
                if (inside) fd_stv<T, VEC>(v_out + (((long)b * H + ty0 + py) * W + tx0 + px) * C + head * HD + vc * VEC, a8);
            } else {
                if (!inside) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) a8[e] = 0.f;
                }
                fd_stv<T, VEC>((sec == 0 ? s_q : s_k) + pix * QK_LD + vc * VEC, a8);
            }
        }
        __syncthreads();
        // each warp owns 32 pixels (2 k-steps of 16) of the tile
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int p0 = warp * 32 + ks * 16;
            uint32_t aq[2][4], ak[2][4], bq[2][4], bk[2][4];
            // A fragments (operand stored [k = pixel][m = channel]) : ldmatrix.trans
            const int ar = p0 + (lane & 7) + 8 * (lane >> 4), ac = 8 * ((lane >> 3) & 1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                ldmatrix_x4_trans(aq[mt], s_q + ar * QK_LD + mt * 16 + ac);
                ldmatrix_x4_trans(ak[mt], s_k + ar * QK_LD + mt * 16 + ac);
            }
            // B fragments (operand stored [k = pixel][n = channel]) : ldmatrix.trans, two n-tiles per x4
            const int br = p0 + (lane & 7) + 8 * ((lane >> 3) & 1), bc = 8 * (lane >> 4);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                ldmatrix_x4_trans(bq[np], s_q + br * QK_LD + np * 16 + bc);
                ldmatrix_x4_trans(bk[np], s_k + br * QK_LD + np * 16 + bc);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_16816<T>(acc[mt][nt], aq[mt], bk[nt >> 1][(nt & 1) * 2], bk[nt >> 1][(nt & 1) * 2 + 1]);
#pragma unroll
                for (int ns = 0; ns < 2; ++ns) {   // the norms only need the diagonal 16x16 blocks
                    mma_16816<T>(accd[0][mt][ns], aq[mt], bq[mt][ns * 2], bq[mt][ns * 2 + 1]);
                    mma_16816<T>(accd[1][mt][ns], ak[mt], bk[mt][ns * 2], bk[mt][ns * 2 + 1]);
                }
            }
        }
    }
    // cross-warp reduction in shared memory, then one global atomic per entry
    {
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int row = mt * 16 + g, col = nt * 8 + 2 * t4;
                atomicAdd(&s_red[row * HD + col], acc[mt][nt][0]);
                atomicAdd(&s_red[row * HD + col + 1], acc[mt][nt][1]);
                atomicAdd(&s_red[(row + 8) * HD + col], acc[mt][nt][2]);
                atomicAdd(&s_red[(row + 8) * HD + col + 1], acc[mt][nt][3]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int ns = 0; ns < 2; ++ns) {
                    float* r = s_red + (1 + a) * HD * HD;
                    const int row = mt * 16 + g, col = mt * 16 + ns * 8 + 2 * t4;
                    if (row == col) atomicAdd(&r[row], accd[a][mt][ns][0]);
                    if (row == col + 1) atomicAdd(&r[row], accd[a][mt][ns][1]);
                    if (row + 8 == col) atomicAdd(&r[row + 8], accd[a][mt][ns][2]);
                    if (row + 8 == col + 1) atomicAdd(&r[row + 8], accd[a][mt][ns][3]);
                }
        }
    }
    __syncthreads();
    float* gg = gram + ((long)b * (C / HD) + head) * HD * HD;
    for (int i = tid; i < HD * HD; i += 256) atomicAdd(gg + i, s_red[i]);
    if (tid < 2 * HD) {
        const int which = tid / HD, c = tid % HD;
        atomicAdd(qk_sq + ((long)b * 2 + which) * C + head * HD + c, s_red[(1 + which) * HD * HD + c]);
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
    const int ntiles = fd_cdiv(H, TPH) * fd_cdiv(W, TPW);
    if (dtype == FD_F32) {
        using T = float;
        dim3 grid(C / HD, ntiles, B);
        const size_t smem = (size_t)(TPH + 2) * (TPW + 2) * 3 * HD * sizeof(T) +
                            (size_t)(2 * TPH * TPW * (HD + 1) + 9 * 3 * HD) * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(dwconv_qkv_gram_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dwconv_qkv_gram_kernel<T><<<grid, 256, smem, stream>>>((const T*)qkv, w, (T*)v, gram, qk_sq, H, W, C);
        FD_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid(C / HD, fd_cdiv(ntiles, TPB), B);
    const size_t smem = ((size_t)(TPH + 2) * (TPW + 2) * 3 * HD + 2 * (size_t)TPH * TPW * QK_LD) * 2 +
                        (size_t)(9 * 3 * HD + 3 * HD * HD) * sizeof(float);
    if (dtype == FD_BF16) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(dwconv_qkv_gram_mma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dwconv_qkv_gram_mma_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>((const __nv_bfloat16*)qkv, w, (__nv_bfloat16*)v, gram, qk_sq, H, W, C, ntiles);
    } else if (dtype == FD_F16) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(dwconv_qkv_gram_mma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dwconv_qkv_gram_mma_kernel<__half><<<grid, 256, smem, stream>>>((const __half*)qkv, w, (__half*)v, gram, qk_sq, H, W, C, ntiles);
    } else {
        return FD_ERR_BAD_ARGUMENT;
    }
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
