// TransposedAttention (channel attention, src/DADiff.py:252-285): depthwise 3x3 over qkv fused with the per-head
// 32x32 Gram matrix q.k^T and the squared norms, then softmax + folding of the attention matrix into the output
// projection (W_eff = W_proj . blockdiag(attn)), so the attn@v + project_out pair becomes one per-sample 1x1 GEMM.
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int HD = 32;            // channels per head (heads = C/32, src/DADiff.py:468)
constexpr int TPH = 8, TPW = 32;  // pixel tile: 8 rows x 32 cols = 256 pixels
constexpr int GRAM_REC = HD * HD + 2 * HD;   // one block's partial: 32x32 Gram + the squared norms of q and k

// Reproducible reduction of the per-block Gram partials (no floating-point atomics): every block has stored its record in
// ws[(b, head), chunk]; the block that arrives last for (b, head) adds the records in CHUNK ORDER and writes gram / qk_sq.
// The grouping of pixels into chunks depends on the sample's geometry only, so a slice gets bit-identical attention in any
// batch.  Counters live behind the records (ws must be zero on entry).  Called by all 256 threads of the block.
__device__ __forceinline__ void gram_finish(float* __restrict__ ws, float* __restrict__ gram, float* __restrict__ qk_sq, int b, int head,
                                            int C, int nchunks, int chunk, int B) {
    __shared__ int s_last;
    const int heads = C / HD, tid = threadIdx.x;
    (void)chunk;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int* counters = reinterpret_cast<int*>(ws + (long)B * heads * nchunks * GRAM_REC);
        s_last = atomicAdd(counters + b * heads + head, 1) == nchunks - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* rec = ws + ((long)b * heads + head) * nchunks * GRAM_REC;
    for (int i = tid; i < GRAM_REC; i += 256) {
        float t = 0.f;
        for (int k = 0; k < nchunks; ++k) t += __ldcg(rec + (long)k * GRAM_REC + i);
        if (i < HD * HD) gram[((long)b * heads + head) * HD * HD + i] = t;
        else qk_sq[((long)b * 2 + (i - HD * HD) / HD) * C + head * HD + (i - HD * HD) % HD] = t;
    }
}

// grid: (C/32 heads, tiles, B).  Each block: dwconv for the head's q, k, v channels on a 8x32 pixel tile.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_qkv_gram_kernel(const T* __restrict__ qkv, const float* __restrict__ w,
                                                              T* __restrict__ v_out, float* __restrict__ gram,
                                                              float* __restrict__ qk_sq, float* __restrict__ ws, int H, int W, int C) {
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
        float* part = ws + (((long)b * (C / HD) + head) * gridDim.y + blockIdx.y) * GRAM_REC;
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcg(part + gi * HD + gj + j, g[j]);
    }
    if (tid < 2 * HD) {
        const float* s = (tid < HD ? s_q : s_k) + (tid % HD);
        float acc = 0.f;
        for (int p = 0; p < TPH * TPW; ++p) { const float v = s[p * (HD + 1)]; acc = fmaf(v, v, acc); }
        __stcg(ws + (((long)b * (C / HD) + head) * gridDim.y + blockIdx.y) * GRAM_REC + HD * HD + tid, acc);
    }
    gram_finish(ws, gram, qk_sq, b, head, C, (int)gridDim.y, (int)blockIdx.y, (int)gridDim.z);
}

// ---------------------------------------------------------------------------------------------------------
// 16-bit storage types: two streaming kernels instead of the fused CUDA-core one.
//
// (1) dwconv3x3_nhwc_kernel — depthwise 3x3 over a channels-last tensor with a REGISTER sliding window: a thread owns
//     one (column, 8-channel vector) and walks down the rows; per input row it loads the three horizontally adjacent
//     vectors (coalesced: an NHWC row is contiguous over (x, c)), feeds the three output rows they contribute to
//     (72 FMAs against weights held in registers) and emits one finished output vector.  No shared memory, no
//     barrier; ~115 instructions per output vector instead of ~250 for the tile/halo formulation.
// (2) gram_mma_kernel — streams q, k (the first 2C channels of the dwconv output) once: 256-pixel tiles staged with
//     cp.async, G = Q^T K and the diagonals of Q^T Q, K^T K on the tensor cores (mma.sync m16n8k16, fp32 accumulate),
//     accumulators kept in registers across the block's pixel range, one global atomic per entry per block.
constexpr int DW_RY = 32;      // output rows per block segment (2 halo rows are recomputed: 6 %)
constexpr int DW_V = 4;        // channels per thread (8-byte accesses; keeps the register window at ~80 registers)

template <typename T> FD_DEVINL uint2 dw_ld_raw(const T* p) { return *reinterpret_cast<const uint2*>(p); }
template <typename T> FD_DEVINL void dw_cvt(uint2 r, float (&v)[4]) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
        const float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}
template <typename T> FD_DEVINL void dw_st(T* p, const float (&v)[4]) {
    uint2 r;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    } else {
        __half2 a = fd_floats2half2_sat(v[0], v[1]), b = fd_floats2half2_sat(v[2], v[3]);
        r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint2*>(p) = r;
}

// A thread owns TWO horizontally adjacent pixels x 4 channels: the 4 input vectors x-1 .. x+2 of a row feed both
// (2 loads + 2 conversions per output instead of 3), weights in registers, 3-row software-pipelined ring.
template <typename T, bool SILU>
__global__ void __launch_bounds__(256, 2) dwconv3x3_nhwc_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                                const float* __restrict__ bias, T* __restrict__ out, int H,
                                                                int W, int C) {
    const int NV = C / DW_V;                            // vectors per pixel
    const long rowv = (long)W * NV;                     // vectors per image row
    const int W2 = (W + 1) / 2;
    const long f2 = (long)blockIdx.x * 256 + threadIdx.x;   // index over (pixel pair, vector)
    if (f2 >= (long)W2 * NV) return;
    const int cv = (int)(f2 % NV), x = 2 * (int)(f2 / NV);
    const long f = (long)x * NV + cv;                   // vector index of the left pixel within a row
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * DW_RY, y1 = min(H, y0 + DW_RY);
    float wr[9][DW_V], bs[DW_V];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long)t * C + cv * DW_V));
        wr[t][0] = wv.x; wr[t][1] = wv.y; wr[t][2] = wv.z; wr[t][3] = wv.w;
    }
    if (bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cv * DW_V));
        bs[0] = bv.x; bs[1] = bv.y; bs[2] = bv.z; bs[3] = bv.w;
    } else {
        bs[0] = bs[1] = bs[2] = bs[3] = 0.f;
    }
    const bool has_l = x > 0, has_1 = x + 1 < W, has_2 = x + 2 < W;
    const T* base = in + (long)b * H * rowv * DW_V;
    T* obase = out + (long)b * H * rowv * DW_V;
    float acc[3][2][DW_V];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < DW_V; ++e) acc[r][px][e] = bs[e];

    // Loads are UNCONDITIONAL (clamped addresses) and masked when consumed three rows later: a predicated load followed by
    // a select makes ptxas wait for the load right away, which serialises the whole prefetch ring on DRAM latency.
    const long offl = has_l ? (long)NV * DW_V : 0, off1 = has_1 ? (long)NV * DW_V : 0, off2 = has_2 ? 2L * NV * DW_V : 0;
    const int ymax = min(H - 1, y1);
    uint2 ring[3][4];                                   // raw vectors x-1, x, x+1, x+2 of input rows yi, yi+1, yi+2
    auto fetch = [&](int yr, auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        const T* rp = base + ((long)min(max(yr, 0), ymax) * rowv + f) * DW_V;
        ring[S][0] = dw_ld_raw<T>(rp - offl);
        ring[S][1] = dw_ld_raw<T>(rp);
        ring[S][2] = dw_ld_raw<T>(rp + off1);
        ring[S][3] = dw_ld_raw<T>(rp + off2);
    };
    // input row `yi` contributes to output rows yi+1 (tap row 0), yi (1), yi-1 (2); accumulator slot = output row mod 3
    auto step = [&](int yi, auto slot_c) {
        constexpr int S = decltype(slot_c)::value;      // == yi mod 3 (compile-time so that acc[] / ring[] stay in registers)
        float v[4][DW_V];
        {
            const bool rv = yi >= 0 && yi < H;
            const bool ok[4] = {rv && has_l, rv, rv && has_1, rv && has_2};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 r = ring[S][j];
                r.x = ok[j] ? r.x : 0u;
                r.y = ok[j] ? r.y : 0u;
                dw_cvt<T>(r, v[j]);
            }
        }
        fetch(yi + 3, slot_c);
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int e = 0; e < DW_V; ++e) {
                    const float xv = v[px + dx][e];
                    acc[(S + 1) % 3][px][e] = fmaf(xv, wr[0 * 3 + dx][e], acc[(S + 1) % 3][px][e]);   // output row yi+1
                    acc[S][px][e] = fmaf(xv, wr[1 * 3 + dx][e], acc[S][px][e]);                       // output row yi
                    acc[(S + 2) % 3][px][e] = fmaf(xv, wr[2 * 3 + dx][e], acc[(S + 2) % 3][px][e]);   // output row yi-1
                }
        const int yo = yi - 1;                          // output row yi-1 is complete (slot (S+2)%3)
        if (yo >= y0 && yo < y1) {
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                if (px == 1 && !has_1) continue;
                float o[DW_V];
#pragma unroll
                for (int e = 0; e < DW_V; ++e) o[e] = SILU ? fd_silu(acc[(S + 2) % 3][px][e]) : acc[(S + 2) % 3][px][e];
                dw_st<T>(obase + ((long)yo * rowv + f + (long)px * NV) * DW_V, o);
            }
        }
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < DW_V; ++e) acc[(S + 2) % 3][px][e] = bs[e];
    };
    int yi = y0 - 1;
    yi -= ((yi % 3) + 3) % 3;                            // round down to a multiple of 3 (extra rows only touch accumulators that are reset)
    fetch(yi, std::integral_constant<int, 0>{});
    fetch(yi + 1, std::integral_constant<int, 1>{});
    fetch(yi + 2, std::integral_constant<int, 2>{});
    for (; yi <= y1; yi += 3) {
        step(yi, std::integral_constant<int, 0>{});
        if (yi + 1 <= y1) step(yi + 1, std::integral_constant<int, 1>{});
        if (yi + 2 <= y1) step(yi + 2, std::integral_constant<int, 2>{});
    }
}

// v2 of the same kernel.  ncu/SASS of v1 (profiles/r1_final_ncu_dwconv_nhwc.txt): issue-bound (73 % issue slots, 120 registers) at
// 570 instructions per 3 rows of which only 216 are FFMA — the rest was 64-bit address arithmetic redone per load (clamp,
// multiply, LEA), the boundary SELs on every lane and the accumulator resets.  Here: (a) ONE running row pointer advanced by a
// block-uniform stride (the row clamp becomes a uniform select of 0 / row pitch), (b) warps with no edge column (99 % at
// W = 512) run a variant with no masks at all, row validity is a block-uniform branch, (c) a finished accumulator slot is
// re-seeded by the first FMA of the next output row (bias as the addend) instead of being reset.
constexpr int DW2_RY = 32;     // output rows per block segment (2 halo rows recomputed: 6 %)

template <typename T, bool SILU, bool EDGE, bool BIAS>
FD_DEVINL void dw_nhwc_rows(const T* __restrict__ in_b, T* __restrict__ out_b, const float (&wr)[9][DW_V],
                            const float (&bs)[DW_V], int H, long rowe, int C, long fe, int y0, int y1, bool has_l,
                            bool has_1, bool has_2, int pf) {
    const int ymax = min(H - 1, y1);
    int yi = y0 - 1;
    yi -= ((yi % 3) + 3) % 3;                            // round down to a multiple of 3
    const int offl = (!EDGE || has_l) ? C : 0, off1 = (!EDGE || has_1) ? C : 0, off2 = (!EDGE || has_2) ? 2 * C : 0;
    const T* p = in_b + (long)min(max(yi, 0), ymax) * rowe + fe;      // row the next fetch reads (clamped into the image)
    T* q = out_b + (long)(yi - 1) * rowe + fe;                        // output row of the current step (only dereferenced in range)
    int yf = yi;
    float acc[3][2][DW_V];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < DW_V; ++e) acc[r][px][e] = 0.f;
    uint2 ring[3][4];
    auto fetch = [&](auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        ring[S][0] = dw_ld_raw<T>(p - offl);
        ring[S][1] = dw_ld_raw<T>(p);
        ring[S][2] = dw_ld_raw<T>(p + off1);
        ring[S][3] = dw_ld_raw<T>(p + off2);
        if (pf > 0 && yf >= 0 && yf + pf <= ymax) {      // block-uniform: pull the row `pf` steps further down into L2
            const T* pp = p + (long)pf * rowe;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + off1));
        }
        p += (yf >= 0 && yf < ymax) ? rowe : 0;          // block-uniform
        ++yf;
    };
    auto step = [&](int yrow, auto slot_c) {
        constexpr int S = decltype(slot_c)::value;       // == yrow mod 3
        constexpr int S1 = (S + 1) % 3, S2 = (S + 2) % 3;
        float v[4][DW_V];
        {
            const bool rv = yrow >= 0 && yrow < H;       // block-uniform; straight-line masking keeps the loads in flight
            const bool ok[4] = {rv && (!EDGE || has_l), rv, rv && (!EDGE || has_1), rv && (!EDGE || has_2)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 r = ring[S][j];
                r.x = ok[j] ? r.x : 0u;
                r.y = ok[j] ? r.y : 0u;
                dw_cvt<T>(r, v[j]);
            }
        }
        fetch(slot_c);
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < DW_V; ++e) {
                float a1 = BIAS ? fmaf(v[px][e], wr[0][e], bs[e]) : v[px][e] * wr[0][e];   // output row yrow+1: first contribution
                a1 = fmaf(v[px + 1][e], wr[1][e], a1);
                acc[S1][px][e] = fmaf(v[px + 2][e], wr[2][e], a1);
                float a0 = fmaf(v[px][e], wr[3][e], acc[S][px][e]);               // output row yrow
                a0 = fmaf(v[px + 1][e], wr[4][e], a0);
                acc[S][px][e] = fmaf(v[px + 2][e], wr[5][e], a0);
                float a2 = fmaf(v[px][e], wr[6][e], acc[S2][px][e]);              // output row yrow-1 (complete after this)
                a2 = fmaf(v[px + 1][e], wr[7][e], a2);
                acc[S2][px][e] = fmaf(v[px + 2][e], wr[8][e], a2);
            }
        const int yo = yrow - 1;
        if (yo >= y0 && yo < y1) {                       // block-uniform
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                if (EDGE && px == 1 && !has_1) continue;
                float o[DW_V];
#pragma unroll
                for (int e = 0; e < DW_V; ++e) o[e] = SILU ? fd_silu(acc[S2][px][e]) : acc[S2][px][e];
                dw_st<T>(q + (px ? C : 0), o);
            }
        }
        q += rowe;
    };
    fetch(std::integral_constant<int, 0>{});
    fetch(std::integral_constant<int, 1>{});
    fetch(std::integral_constant<int, 2>{});
    for (; yi <= y1; yi += 3) {
        step(yi, std::integral_constant<int, 0>{});
        if (yi + 1 <= y1) step(yi + 1, std::integral_constant<int, 1>{});
        if (yi + 2 <= y1) step(yi + 2, std::integral_constant<int, 2>{});
    }
}

template <typename T, bool SILU, bool BIAS>
__global__ void __launch_bounds__(256, 2) dwconv3x3_nhwc_v2_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, T* __restrict__ out, int H,
                                                                   int W, int C, int pf) {
    const int NV = C / DW_V;
    const int W2 = (W + 1) / 2;
    const long f2 = (long)blockIdx.x * 256 + threadIdx.x;   // index over (pixel pair, vector)
    const bool live = f2 < (long)W2 * NV;
    const long f2c = live ? f2 : 0;
    const int cv = (int)(f2c % NV), x = 2 * (int)(f2c / NV);
    const int y0 = blockIdx.y * DW2_RY, y1 = min(H, y0 + DW2_RY);
    float wr[9][DW_V], bs[DW_V];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long)t * C + cv * DW_V));
        wr[t][0] = wv.x; wr[t][1] = wv.y; wr[t][2] = wv.z; wr[t][3] = wv.w;
    }
    if (BIAS) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cv * DW_V));
        bs[0] = bv.x; bs[1] = bv.y; bs[2] = bv.z; bs[3] = bv.w;
    } else {
        bs[0] = bs[1] = bs[2] = bs[3] = 0.f;
    }
    const bool has_l = x > 0, has_1 = x + 1 < W, has_2 = x + 2 < W;
    const long rowe = (long)W * C;
    const T* in_b = in + (long)blockIdx.z * H * rowe;
    T* out_b = out + (long)blockIdx.z * H * rowe;
    const long fe = (long)x * C + cv * DW_V;
    const bool edge = !(has_l && has_1 && has_2);
    if (!live) return;                                   // whole trailing warps only matter for the vote below when partially live
    if (__any_sync(__activemask(), edge)) dw_nhwc_rows<T, SILU, true, BIAS>(in_b, out_b, wr, bs, H, rowe, C, fe, y0, y1, has_l, has_1, has_2, pf);
    else dw_nhwc_rows<T, SILU, false, BIAS>(in_b, out_b, wr, bs, H, rowe, C, fe, y0, y1, true, true, true, pf);
}

constexpr int QK_LD = 40;      // padded row (elements) of the q / k tiles: conflict-free ldmatrix
constexpr int GR_TILE = 128;   // pixels per staged tile
constexpr int GR_STAGES = 4;   // cp.async ring: 3 tiles (24 KB per block, 2 blocks per SM) in flight — with two 256-pixel buffers (one
                               // tile in flight per block) the kernel was bound by latency x bytes in flight at 0.57 of the HBM roof
constexpr int GR_PIX = 4096;   // pixels per block

template <typename T>
__global__ void __launch_bounds__(256) gram_mma_kernel(const T* __restrict__ qkv, int ld, float* __restrict__ gram,
                                                       float* __restrict__ qk_sq, float* __restrict__ ws, int P, int C, int pf_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_qk = reinterpret_cast<T*>(smem_raw);                      // [GR_STAGES][2 (q,k)][GR_TILE][QK_LD]
    float* s_red = reinterpret_cast<float*>(s_qk + GR_STAGES * 2 * GR_TILE * QK_LD);   // [32*32 + 2*32]
    const int head = blockIdx.x, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long p_begin = (long)blockIdx.y * GR_PIX;
    const long p_end = min((long)P, p_begin + GR_PIX);
    const int ntiles = (int)((p_end - p_begin + GR_TILE - 1) / GR_TILE);
    for (int i = tid; i < HD * HD + 2 * HD; i += 256) s_red[i] = 0.f;

    auto stage = [&](int tile, int buf) {   // 256 pixels x (q 4 vectors + k 4 vectors), zero-filled past p_end
        T* dst = s_qk + (size_t)buf * 2 * GR_TILE * QK_LD;
        for (int i = tid; i < GR_TILE * 8; i += 256) {
            const int pix = i >> 3, sv = i & 7, sec = sv >> 2, vc = sv & 3;
            const long p = p_begin + (long)tile * GR_TILE + pix;
            const bool ok = p < p_end;
            const T* src = qkv + ((long)b * P + (ok ? p : p_begin)) * ld + sec * C + head * HD + vc * 8;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + (sec * GR_TILE + pix) * QK_LD + vc * 8);
            const int sz = ok ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[2][4][4], accd[2][2][2][4];
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
#pragma unroll
    for (int st = 0; st < GR_STAGES - 1; ++st) {
        if (st < ntiles) stage(st, st);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile % GR_STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(GR_STAGES - 2) : "memory");          // tile `tile` has landed (this thread's part)
        __syncthreads();                                 // ... for everybody, and every warp is done with tile - 1: its buffer is refilled
        if (tile + GR_STAGES - 1 < ntiles) stage(tile + GR_STAGES - 1, (tile + GR_STAGES - 1) % GR_STAGES);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const T* s_q = s_qk + (size_t)buf * 2 * GR_TILE * QK_LD;
        const T* s_k = s_q + GR_TILE * QK_LD;
#pragma unroll
        for (int ks = 0; ks < GR_TILE / 128; ++ks) {     // each warp owns GR_TILE / 8 pixels of the tile
            const int p0 = warp * (GR_TILE / 8) + ks * 16;
            uint32_t aq[2][4], ak[2][4], bq[2][4], bk[2][4];
            const int ar = p0 + (lane & 7) + 8 * (lane >> 4), ac = 8 * ((lane >> 3) & 1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                ldmatrix_x4_trans(aq[mt], s_q + ar * QK_LD + mt * 16 + ac);
                ldmatrix_x4_trans(ak[mt], s_k + ar * QK_LD + mt * 16 + ac);
            }
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
                for (int ns = 0; ns < 2; ++ns) {         // the norms only need the diagonal 16x16 blocks
                    mma_16816<T>(accd[0][mt][ns], aq[mt], bq[mt][ns * 2], bq[mt][ns * 2 + 1]);
                    mma_16816<T>(accd[1][mt][ns], ak[mt], bk[mt][ns * 2], bk[mt][ns * 2 + 1]);
                }
            }
        }
    }
    __syncthreads();
    // the 8 warps add their accumulators into s_red one after the other (warp order = pixel order): no shared-memory atomics
    for (int wsel = 0; wsel < 8; ++wsel) {
        if (warp == wsel) {
            const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int row = mt * 16 + g, col = nt * 8 + 2 * t4;
                    s_red[row * HD + col] += acc[mt][nt][0];
                    s_red[row * HD + col + 1] += acc[mt][nt][1];
                    s_red[(row + 8) * HD + col] += acc[mt][nt][2];
                    s_red[(row + 8) * HD + col + 1] += acc[mt][nt][3];
                }
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int ns = 0; ns < 2; ++ns) {
                        float* r = s_red + HD * HD + a * HD;
                        const int row = mt * 16 + g, col = mt * 16 + ns * 8 + 2 * t4;
                        if (row == col) r[row] += accd[a][mt][ns][0];
                        if (row == col + 1) r[row] += accd[a][mt][ns][1];
                        if (row + 8 == col) r[row + 8] += accd[a][mt][ns][2];
                        if (row + 8 == col + 1) r[row + 8] += accd[a][mt][ns][3];
                    }
            }
        }
        __syncthreads();
    }
    float* part = ws + (((long)b * (C / HD) + head) * gridDim.y + blockIdx.y) * GRAM_REC;
    for (int i = tid; i < GRAM_REC; i += 256) __stcg(part + i, s_red[i]);
    gram_finish(ws, gram, qk_sq, b, head, C, (int)gridDim.y, (int)blockIdx.y, (int)gridDim.z);
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

extern "C" long fd_gram_ws_floats(int B, int H, int W, int C, int dtype) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % HD) return 0;
    const long chunks = dtype == FD_F32 ? (long)fd_cdiv(H, TPH) * fd_cdiv(W, TPW) : fd_cdiv((long)H * W, GR_PIX);
    return (long)B * (C / HD) * chunks * GRAM_REC + (long)B * (C / HD);
}

extern "C" int fd_dwconv3x3_qkv_gram(const void* qkv, const float* w, void* v, float* gram, float* qk_sq, float* ws, int B, int H,
                                     int W, int C, int dtype, cudaStream_t stream) {
    if (!qkv || !w || !v || !gram || !qk_sq || !ws || B <= 0 || H <= 0 || W <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
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
        dwconv_qkv_gram_kernel<T><<<grid, 256, smem, stream>>>((const T*)qkv, w, (T*)v, gram, qk_sq, ws, H, W, C);
        FD_LAUNCH_CHECK();
        return 0;
    }
    return FD_ERR_UNSUPPORTED;   // 16-bit types: use fd_dwconv3x3_nhwc + fd_gram_qk
}

template <typename T>
static int dwconv_nhwc_launch(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int C, int silu,
                              cudaStream_t stream) {
    const long pairs = (long)((W + 1) / 2) * (C / DW_V);
    static const bool use_v1 = getenv("FD_DWCONV_V1") != nullptr;      // A/B switch for the measurement scripts
    if (!use_v1) {
        static const int pf = getenv("FD_DWCONV_PF") ? atoi(getenv("FD_DWCONV_PF")) : 4;     // L2 prefetch distance in rows (0 = off): 700 -> 632 us
        dim3 grid((unsigned)fd_cdiv(pairs, 256), (unsigned)fd_cdiv(H, DW2_RY), (unsigned)B);
        if (silu && bias) dwconv3x3_nhwc_v2_kernel<T, true, true><<<grid, 256, 0, stream>>>((const T*)in, w, bias, (T*)out, H, W, C, pf);
        else if (silu) dwconv3x3_nhwc_v2_kernel<T, true, false><<<grid, 256, 0, stream>>>((const T*)in, w, bias, (T*)out, H, W, C, pf);
        else if (bias) dwconv3x3_nhwc_v2_kernel<T, false, true><<<grid, 256, 0, stream>>>((const T*)in, w, bias, (T*)out, H, W, C, pf);
        else dwconv3x3_nhwc_v2_kernel<T, false, false><<<grid, 256, 0, stream>>>((const T*)in, w, bias, (T*)out, H, W, C, pf);
        FD_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((unsigned)fd_cdiv(pairs, 256), (unsigned)fd_cdiv(H, DW_RY), (unsigned)B);
    if (silu) dwconv3x3_nhwc_kernel<T, true><<<grid, 256, 0, stream>>>((const T*)in, w, bias, (T*)out, H, W, C);
    else dwconv3x3_nhwc_kernel<T, false><<<grid, 256, 0, stream>>>((const T*)in, w, bias, (T*)out, H, W, C);
    FD_LAUNCH_CHECK();
    return 0;
}

extern "C" int fd_dwconv3x3_nhwc(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int C,
                                 int silu, int dtype, cudaStream_t stream) {
    if (!in || !w || !out || in == out || B <= 0 || H <= 0 || W <= 0 || C <= 0) return FD_ERR_BAD_ARGUMENT;
    if (C % 8) return FD_ERR_UNSUPPORTED;
    if (dtype == FD_BF16) return dwconv_nhwc_launch<__nv_bfloat16>(in, w, bias, out, B, H, W, C, silu, stream);
    if (dtype == FD_F16) return dwconv_nhwc_launch<__half>(in, w, bias, out, B, H, W, C, silu, stream);
    return FD_ERR_UNSUPPORTED;
}

template <typename T>
static int gram_launch(const void* qkv, int ld, float* gram, float* qk_sq, float* ws, int B, int P, int C, cudaStream_t stream) {
    const size_t smem = (size_t)GR_STAGES * 2 * GR_TILE * QK_LD * sizeof(T) + (size_t)(HD * HD + 2 * HD) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gram_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((unsigned)(C / HD), (unsigned)fd_cdiv(P, GR_PIX), (unsigned)B);
    // L2 prefetch distance in 256-pixel tiles.  OFF: measured 235 -> 266 us at 16x262144x64 with 2 or 4 tiles (unlike the depthwise
    // kernels, where it gave 10 %): the q / k half-lines of the two heads are already shared through L2 by neighbouring blocks
    static const int pf_tiles = getenv("FD_GRAM_PF") ? atoi(getenv("FD_GRAM_PF")) : 0;
    gram_mma_kernel<T><<<grid, 256, smem, stream>>>((const T*)qkv, ld, gram, qk_sq, ws, P, C, pf_tiles);
    FD_LAUNCH_CHECK();
    return 0;
}

extern "C" int fd_gram_qk(const void* qkv, int ld, float* gram, float* qk_sq, float* ws, int B, int P, int C, int dtype,
                          cudaStream_t stream) {
    if (!qkv || !gram || !qk_sq || !ws || B <= 0 || P <= 0 || C <= 0 || ld < 2 * C) return FD_ERR_BAD_ARGUMENT;
    if (C % HD || ld % 8) return FD_ERR_UNSUPPORTED;
    if (dtype == FD_BF16) return gram_launch<__nv_bfloat16>(qkv, ld, gram, qk_sq, ws, B, P, C, stream);
    if (dtype == FD_F16) return gram_launch<__half>(qkv, ld, gram, qk_sq, ws, B, P, C, stream);
    return FD_ERR_UNSUPPORTED;
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
