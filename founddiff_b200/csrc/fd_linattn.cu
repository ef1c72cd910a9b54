// LinearAttention of the lucidrains Unet (src/denoising_diffusion_pytorch.py:227-255), heads of 32 channels over N tokens:
//   q = softmax_d(q) * scale;  k = softmax_n(k);  v = v / N;  context[d,e] = sum_n k[d,n] v[e,n];
//   out[e,n] = sum_d context[d,e] q[d,n];  then to_out = Conv1x1(+bias) -> channel LayerNorm(g).
// Decomposition (all channels-last, qkv (B, N, 3*heads*32) = [q | k | v]):
//   fd_linattn_kmax     column max of k over the tokens (softmax_n stabiliser)
//   fd_linattn_context  ctx_raw[b,h,d,e] += sum_n exp(k - max)[n,d] v[n,e], ksum[b,hd] += sum_n exp(k - max)[n,d]
//                       — 256-token tiles staged with cp.async, exp applied in shared memory, products on the tensor cores
//   fd_linattn_weff     folds context / (ksum * N) * scale into the output projection:
//                       weff[b, o, h*32+d] = sum_e Wout[o, h*32+e] * ctx[b,h,d,e]  -> to_out(out) == qhat @ weff[b]^T + bias,
//                       a per-sample 1x1 GEMM on the tcgen05 path (fd_conv2d_tc, per_batch_weight)
//   fd_softmax_d32      qhat = softmax over each head's 32 channels of q (per token)
// The trailing channel LayerNorm is fd_ln_modulate with zero modulation.
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int HD = 32;
constexpr int LA_LD = 40;
constexpr int LA_TILE = 256;
constexpr int LA_PIX = 4096;

FD_DEVINL void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// grid (N / 1024 token ranges, B); block 256: thread = (row parity group, 8-channel vector of k)
template <typename T>
__global__ void __launch_bounds__(256) linattn_kmax_kernel(const T* __restrict__ qkv, float* __restrict__ kmax, int N, int HC) {
    const int NV = HC / 8;
    const int b = blockIdx.y;
    const int vec = threadIdx.x % NV, rstep = 256 / NV, r0 = threadIdx.x / NV;
    const int n0 = blockIdx.x * 1024, n1 = min(N, n0 + 1024);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    if (threadIdx.x < rstep * NV)
        for (int n = n0 + r0; n < n1; n += rstep) {
            float v[8];
            fd_ldv<T, 8>(qkv + ((long)b * N + n) * 3 * HC + HC + vec * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
        }
    if (threadIdx.x < rstep * NV) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomic_max_float(kmax + (long)b * HC + vec * 8 + e, m[e]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) linattn_context_kernel(const T* __restrict__ qkv, const float* __restrict__ kmax,
                                                              float* __restrict__ ctx_raw, float* __restrict__ ksum, int N, int heads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_kv = reinterpret_cast<T*>(smem_raw);                        // [2 buffers][2 (k, v)][256][LA_LD]
    float* s_red = reinterpret_cast<float*>(s_kv + 2 * 2 * LA_TILE * LA_LD);   // [32*32 + 32]
    const int HC = heads * HD, ld = 3 * HC;
    const int head = blockIdx.x, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long p_begin = (long)blockIdx.y * LA_PIX, p_end = min((long)N, p_begin + LA_PIX);
    const int ntiles = (int)((p_end - p_begin + LA_TILE - 1) / LA_TILE);
    for (int i = tid; i < HD * HD + HD; i += 256) s_red[i] = 0.f;

    auto stage = [&](int tile, int buf) {
        T* dst = s_kv + (size_t)buf * 2 * LA_TILE * LA_LD;
        for (int i = tid; i < LA_TILE * 8; i += 256) {
            const int pix = i >> 3, sv = i & 7, sec = sv >> 2, vc = sv & 3;
            const long p = p_begin + (long)tile * LA_TILE + pix;
            const bool ok = p < p_end;
            const T* src = qkv + ((long)b * N + (ok ? p : p_begin)) * ld + (1 + sec) * HC + head * HD + vc * 8;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (sec * LA_TILE + pix) * LA_LD + vc * 8);
            const int sz = ok ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // thread = (pixel row, vector) of the exp pass
    const int epix = tid >> 2, evc = tid & 3;                       // 64 pixel rows x 4 vectors per pass
    float mx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) mx[e] = kmax[(long)b * HC + head * HD + evc * 8 + e];

    stage(0, 0);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) { stage(tile + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        T* s_k = s_kv + (size_t)buf * 2 * LA_TILE * LA_LD;
        T* s_v = s_k + LA_TILE * LA_LD;
        // k <- exp(k - max) in place (rows past the end were zero-filled: force them to 0)
#pragma unroll
        for (int pass = 0; pass < LA_TILE / 64; ++pass) {
            const int pix = pass * 64 + epix;
            const bool ok = p_begin + (long)tile * LA_TILE + pix < p_end;
            float v[8];
            fd_ldv<T, 8>(s_k + pix * LA_LD + evc * 8, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[e] = ok ? __expf(v[e] - mx[e]) : 0.f;
                // column sums use the ROUNDED value so that normalisation matches the tensor-core product
            }
            fd_stv<T, 8>(s_k + pix * LA_LD + evc * 8, v);
            float w[8];
            fd_ldv<T, 8>(s_k + pix * LA_LD + evc * 8, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) csum[e] += w[e];
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int p0 = warp * 32 + ks * 16;
            uint32_t ak[2][4], bv[2][4];
            const int ar = p0 + (lane & 7) + 8 * (lane >> 4), ac = 8 * ((lane >> 3) & 1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) ldmatrix_x4_trans(ak[mt], s_k + ar * LA_LD + mt * 16 + ac);
            const int br = p0 + (lane & 7) + 8 * ((lane >> 3) & 1), bc = 8 * (lane >> 4);
#pragma unroll
            for (int np = 0; np < 2; ++np) ldmatrix_x4_trans(bv[np], s_v + br * LA_LD + np * 16 + bc);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_16816<T>(acc[mt][nt], ak[mt], bv[nt >> 1][(nt & 1) * 2], bv[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncthreads();
    }
    {
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int row = mt * 16 + g, col = nt * 8 + 2 * t4;
                atomicAdd(&s_red[row * HD + col], acc[mt][nt][0]);
                atomicAdd(&s_red[row * HD + col + 1], acc[mt][nt][1]);
                atomicAdd(&s_red[(row + 8) * HD + col], acc[mt][nt][2]);
                atomicAdd(&s_red[(row + 8) * HD + col + 1], acc[mt][nt][3]);
            }
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&s_red[HD * HD + evc * 8 + e], csum[e]);
    }
    __syncthreads();
    float* cc = ctx_raw + ((long)b * heads + head) * HD * HD;
    for (int i = tid; i < HD * HD; i += 256) atomicAdd(cc + i, s_red[i]);
    if (tid < HD) atomicAdd(ksum + (long)b * HC + head * HD + tid, s_red[HD * HD + tid]);
}

// grid (heads, B), block 256:  weff[b, o, h*32+d] = scale/(N*ksum[d]) * sum_e wout[o, h*32+e] * ctx_raw[b,h,d,e]
template <typename T>
__global__ void __launch_bounds__(256) linattn_weff_kernel(const float* __restrict__ ctx_raw, const float* __restrict__ ksum,
                                                           const float* __restrict__ wout, T* __restrict__ weff, int dim, int heads,
                                                           float scale_over_n) {
    __shared__ float s_c[HD][HD + 1];
    const int head = blockIdx.x, b = blockIdx.y, HC = heads * HD;
    for (int i = threadIdx.x; i < HD * HD; i += 256) {
        const int d = i / HD, e = i % HD;
        s_c[d][e] = ctx_raw[(((long)b * heads + head) * HD + d) * HD + e] * scale_over_n / ksum[(long)b * HC + head * HD + d];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < dim * HD; i += 256) {
        const int o = i / HD, d = i % HD;
        const float* w = wout + (long)o * HC + head * HD;
        float acc = 0.f;
#pragma unroll 8
        for (int e = 0; e < HD; ++e) acc = fmaf(__ldg(w + e), s_c[d][e], acc);
        fd_st(weff + ((long)b * dim + o) * HC + head * HD + d, acc);
    }
}

// one thread per (token, head): softmax over the head's 32 q channels
template <typename T>
__global__ void __launch_bounds__(256) softmax_d32_kernel(const T* __restrict__ qkv, T* __restrict__ qhat, long tokens, int heads) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= tokens * heads) return;
    const long tok = i / heads;
    const int head = (int)(i % heads);
    const int HC = heads * HD;
    const T* src = qkv + tok * 3 * HC + head * HD;
    float v[4][8], m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        fd_ldv<T, 8>(src + j * 8, v[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) m = fmaxf(m, v[j][e]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[j][e] = __expf(v[j][e] - m); s += v[j][e]; }
    const float inv = 1.f / s;
    T* dst = qhat + tok * HC + head * HD;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] *= inv;
        fd_stv<T, 8>(dst + j * 8, v[j]);
    }
}

template <typename T>
int linattn_launch(const void* qkv, float* kmax, float* ksum, float* ctx_raw, int B, int N, int heads, cudaStream_t stream) {
    const int HC = heads * HD;
    if (256 % (HC / 8) || HC / 8 > 256) return FD_ERR_UNSUPPORTED;
    linattn_kmax_kernel<T><<<dim3(fd_cdiv(N, 1024), B), 256, 0, stream>>>((const T*)qkv, kmax, N, HC);
    FD_LAUNCH_CHECK();
    const size_t smem = (size_t)2 * 2 * LA_TILE * LA_LD * sizeof(T) + (size_t)(HD * HD + HD) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linattn_context_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    linattn_context_kernel<T><<<dim3(heads, fd_cdiv(N, LA_PIX), B), 256, smem, stream>>>((const T*)qkv, kmax, ctx_raw, ksum, N, heads);
    FD_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// kmax (B, heads*32) must be pre-filled with -inf, ksum (B, heads*32) and ctx_raw (B, heads, 32, 32) with zeros.
extern "C" int fd_linattn_context(const void* qkv, float* kmax, float* ksum, float* ctx_raw, int B, int N, int heads, int dtype,
                                  cudaStream_t stream) {
    if (!qkv || !kmax || !ksum || !ctx_raw || B <= 0 || N <= 0 || heads <= 0) return FD_ERR_BAD_ARGUMENT;
    if (dtype == FD_BF16) return linattn_launch<__nv_bfloat16>(qkv, kmax, ksum, ctx_raw, B, N, heads, stream);
    if (dtype == FD_F16) return linattn_launch<__half>(qkv, kmax, ksum, ctx_raw, B, N, heads, stream);
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_linattn_weff(const float* ctx_raw, const float* ksum, const float* wout, void* weff, int B, int N, int heads,
                               int dim, float scale, int dtype, cudaStream_t stream) {
    if (!ctx_raw || !ksum || !wout || !weff || B <= 0 || N <= 0 || heads <= 0 || dim <= 0) return FD_ERR_BAD_ARGUMENT;
    dim3 grid(heads, B);
    FD_DISPATCH_DTYPE(dtype, T, (linattn_weff_kernel<T><<<grid, 256, 0, stream>>>(ctx_raw, ksum, wout, (T*)weff, dim, heads, scale / (float)N)));
    FD_LAUNCH_CHECK();
    return 0;
}

extern "C" int fd_softmax_d32(const void* qkv, void* qhat, int B, int N, int heads, int dtype, cudaStream_t stream) {
    if (!qkv || !qhat || B <= 0 || N <= 0 || heads <= 0) return FD_ERR_BAD_ARGUMENT;
    const long tokens = (long)B * N;
    const unsigned grid = (unsigned)fd_cdiv(tokens * heads, 256);
    if (dtype == FD_BF16) softmax_d32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)qhat, tokens, heads);
    else if (dtype == FD_F16) softmax_d32_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)qkv, (__half*)qhat, tokens, heads);
    else return FD_ERR_UNSUPPORTED;
    FD_LAUNCH_CHECK();
    return 0;
}
