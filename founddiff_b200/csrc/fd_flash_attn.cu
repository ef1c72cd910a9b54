// Bottleneck self-attention of the lucidrains Unet (src/denoising_diffusion_pytorch.py:257-279): heads of d = 32 over
// N = H*W tokens (4096 at the 512^2 bottleneck), softmax(q^T k * scale) v — a flash-style kernel: 64-query tiles, K/V
// tiles of 64 keys double-buffered in shared memory with cp.async, S = Q K^T and O += P V on the tensor cores
// (mma.sync m16n8k16, fp32 accumulate), online softmax in registers (exp2 with the scale folded in).
// Layout: qkv (B, N, 3*heads*32) channels-last as produced by the to_qkv 1x1 GEMM (q | k | v column blocks, 'b (h c)'
// channel order); out (B, N, heads*32).  16-bit storage types.
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int FA_D = 32;
constexpr int FA_BQ = 64;      // queries per block (4 warps x 16)
constexpr int FA_BK = 64;      // keys per tile
constexpr int FA_LD = 40;      // padded smem row (elements): conflict-free ldmatrix

FD_DEVINL float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <typename T> FD_DEVINL uint32_t pack2(float a, float b) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = fd_floats2half2_sat(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) flash_attn_d32_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int heads,
                                                             float scale_log2e) {
    __shared__ __align__(16) T s_q[FA_BQ * FA_LD];
    __shared__ __align__(16) T s_k[2][FA_BK * FA_LD];
    __shared__ __align__(16) T s_v[2][FA_BK * FA_LD];
    const int HC = heads * FA_D, ld = 3 * HC;
    const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * FA_BQ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const T* base = qkv + (long)b * N * ld;

    auto cp16 = [](void* dst, const void* src, bool ok) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
        const int sz = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
    };
    // Q tile: 64 rows x 4 vectors
    for (int i = tid; i < FA_BQ * 4; i += 128) {
        const int r = i >> 2, v = i & 3;
        const bool ok = q0 + r < N;
        cp16(s_q + r * FA_LD + v * 8, base + (long)(ok ? q0 + r : 0) * ld + head * FA_D + v * 8, ok);
    }
    auto stage = [&](int kt, int buf) {
        for (int i = tid; i < FA_BK * 8; i += 128) {
            const int r = i >> 3, sv = i & 7, isv = sv >> 2, v = sv & 3;
            const int key = kt * FA_BK + r;
            const bool ok = key < N;
            const T* src = base + (long)(ok ? key : 0) * ld + (1 + isv) * HC + head * FA_D + v * 8;
            cp16((isv ? s_v[buf] : s_k[buf]) + r * FA_LD + v * 8, src, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int ntiles = (N + FA_BK - 1) / FA_BK;
    stage(0, 0);

    uint32_t aq[2][4];
    float o[4][4], m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;

    for (int kt = 0; kt < ntiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) { stage(kt + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
                ldmatrix_x4(aq[ks], s_q + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * FA_LD + ks * 16 + 8 * (lane >> 4));
        }
        // S = Q K^T  (16 queries x 64 keys per warp)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bk[4];
                ldmatrix_x4(bk, s_k[buf] + (np * 16 + (lane & 7) + 8 * (lane >> 4)) * FA_LD + ks * 16 + 8 * ((lane >> 3) & 1));
                mma_16816<T>(s[2 * np], aq[ks], bk[0], bk[1]);
                mma_16816<T>(s[2 * np + 1], aq[ks], bk[2], bk[3]);
            }
        // mask keys past N (only in the last tile), online softmax
        const int kbase = kt * FA_BK;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = kbase + nt * 8 + 2 * t4 + (e & 1);
                float v = s[nt][e] * scale_log2e;
                if (key >= N) v = -INFINITY;
                s[nt][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = ex2f(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float ls[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = ex2f(s[nt][e] - m_run[e >> 1]);
                s[nt][e] = p;
                ls[e >> 1] += p;
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 1);
            ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 2);
            l_run[r] = l_run[r] * corr[r] + ls[r];
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            o[nt][0] *= corr[0]; o[nt][1] *= corr[0];
            o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
        }
        // O += P V  (P from the S accumulators, V tile [key][d] via ldmatrix.trans)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t ap[4];
            ap[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
            ap[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
            ap[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            ap[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {
                uint32_t bv[4];
                ldmatrix_x4_trans(bv, s_v[buf] + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * FA_LD + dp * 16 + 8 * (lane >> 4));
                mma_16816<T>(o[2 * dp], ap, bv[0], bv[1]);
                mma_16816<T>(o[2 * dp + 1], ap, bv[2], bv[3]);
            }
        }
        __syncthreads();       // this buffer may be refilled by the next-but-one stage
    }
    // normalise and store: thread holds rows g, g+8 of its warp's 16 queries; cols nt*8 + 2t, +1
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qi = q0 + warp * 16 + g + 8 * r;
        if (qi >= N) continue;
        const float inv = 1.f / l_run[r];
        T* orow = out + ((long)b * N + qi) * HC + head * FA_D;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * t4) = pack2<T>(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    }
}

}  // namespace

extern "C" int fd_flash_attn_d32(const void* qkv, void* out, int B, int N, int heads, float scale, int dtype, cudaStream_t stream) {
    if (!qkv || !out || B <= 0 || N <= 0 || heads <= 0) return FD_ERR_BAD_ARGUMENT;
    dim3 grid((unsigned)fd_cdiv(N, FA_BQ), (unsigned)heads, (unsigned)B);
    const float sl = scale * 1.4426950408889634f;
    if (dtype == FD_BF16) flash_attn_d32_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, N, heads, sl);
    else if (dtype == FD_F16) flash_attn_d32_kernel<__half><<<grid, 128, 0, stream>>>((const __half*)qkv, (__half*)out, N, heads, sl);
    else return FD_ERR_UNSUPPORTED;
    FD_LAUNCH_CHECK();
    return 0;
}
