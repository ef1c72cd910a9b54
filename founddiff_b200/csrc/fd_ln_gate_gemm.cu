// SS2D tail fused with out_proj and the Mamba block's gated residual (src/emamba2.py:365, 747-748; src/DADiff.py:486):
//
//     out[p, :] = addend[p, :] + gate[b, :] * ( W_out . ((LN_D(y[p, :]) * gamma + beta) * z[p, :] + local[b, :]) )
//
// in ONE pass over the pixels.  As two launches (fd_ln_gate, then the 1x1 fd_conv2d_tc) the gated row g was written and read
// back: (3 D + D + 2 C) elements of HBM traffic per pixel against (2 D + 2 C) here, 40 % less at D = 2 C — the pair was 9 % of
// the sampling step at the full-resolution level, both halves HBM-bound.
//
// The GEMM is small (K = D = 128, N = C = 64: 16 KFLOP per 640 bytes of traffic) and the kernel stays HBM-bound, so it runs on
// warp-level mma.sync (m16n8k16, fp32 accumulate) with the A operand built IN REGISTERS from the rows the warp has just
// normalised — no shared-memory round trip for A at all:
//   * a warp owns 16 consecutive pixels; lane (g = lane / 4, t = lane % 4) holds the 16-byte pieces t, t + 4, t + 8, t + 12 of
//     the rows of pixels g and g + 8 (4 x LDG.128 per row and tensor: the four lanes of a quad read 64 contiguous bytes per
//     instruction — with 64 contiguous bytes per LANE instead, every warp load touched 16 cache lines half a sector at a time and
//     the kernel sat at 3.2 TB/s on the L1 tag stage), so a row's LayerNorm statistics are two shuffles inside the quad;
//   * the K index of a GEMM may be permuted freely as long as A and B agree: k-tile j takes, from thread t, the four channels
//     32 (j / 2) + 8 t + 4 (j % 2) + {0, 1, 2, 3} as the fragment's (k = 2t, 2t+1, 2t+8, 2t+9) — exactly what the thread holds;
//   * the weight is re-laid once per block into fragment order in shared memory (one conflict-free LDS.64 per MMA).
// The row g is rounded to the operand type before the product, as the two-launch form did when it stored it.
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int LG_D = 128, LG_C = 64;          // d_inner and d_model of the full-resolution level
constexpr int LG_WARPS = 8;
constexpr int LG_KT = LG_D / 16, LG_NT = LG_C / 8;
constexpr int LG_OP = LG_C + 8;                // row pitch (floats) of a warp's output tile in shared memory

template <typename T> FD_DEVINL float2 lg_unpack(uint32_t w) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
    else return __half22float2(*reinterpret_cast<__half2*>(&w));
}
template <typename T> FD_DEVINL u64 lg_unpack2(uint32_t w) {           // two 16-bit values -> a packed fp32 pair
    const float2 f = lg_unpack<T>(w);
    return f2_pack(f.x, f.y);
}
template <typename T> FD_DEVINL uint32_t lg_pack(float a, float b) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = fd_floats2half2_sat(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(LG_WARPS * 32, 2) ln_gate_out_proj_kernel(
    const TI* __restrict__ y, const TI* __restrict__ xz, int ld, int z_off, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ local, const TI* __restrict__ w, const float* __restrict__ gate, int gate_stride,
    const TO* __restrict__ addend, TO* __restrict__ out, unsigned tiles, unsigned P, float eps) {
    __shared__ __align__(16) unsigned long long s_w[LG_NT * LG_KT * 32];      // B fragments: [n-tile][k-tile][lane] = 4 elements
    __shared__ __align__(16) float s_gamma[LG_D], s_beta[LG_D];
    extern __shared__ __align__(16) float s_out[];                              // [LG_WARPS][16][LG_OP] output tiles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < LG_NT * LG_KT * 32; i += LG_WARPS * 32) {
        const int ln = i & 31, j = (i >> 5) % LG_KT, nn = i / (32 * LG_KT);
        s_w[i] = *reinterpret_cast<const unsigned long long*>(w + (long)(8 * nn + (ln >> 2)) * LG_D + 32 * (j >> 1) + 8 * (ln & 3) + 4 * (j & 1));
    }
    for (int i = threadIdx.x; i < LG_D; i += LG_WARPS * 32) { s_gamma[i] = gamma[i]; s_beta[i] = beta[i]; }
    __syncthreads();

    for (unsigned tile = blockIdx.x * LG_WARPS + warp; tile < tiles; tile += gridDim.x * LG_WARPS) {
        const unsigned row0 = tile * 16 + g;                    // this lane's pixels: row0 and row0 + 8 (same sample: P % 16 == 0)
        const unsigned b = row0 / P;
        uint4 ry[2][4], rz[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const TI* yr = y + (long)(row0 + 8 * r) * LG_D + 8 * t;
            const TI* zr = xz + (long)(row0 + 8 * r) * ld + z_off + 8 * t;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                ry[r][v] = *reinterpret_cast<const uint4*>(yr + 32 * v);
                rz[r][v] = *reinterpret_cast<const uint4*>(zr + 32 * v);
            }
        }
        // LayerNorm statistics: sum and sum of squares in one pass over the raw registers, as packed fp32 pairs (one FADD2 + one
        // FFMA2 per two channels; the kernel is bound by issue slots — 16 rows x 640 bytes per ~700 warp instructions — not by HBM,
        // so the two-pass form of fd_row_stats, which converts every value twice, costs 15 % here).  A row lives in the four
        // lanes of a quad.
        u64 nrm_a[2], nrm_c[2];                                // (y - mean) rstd = y * rstd + (-mean rstd), both rows, as pairs
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            u64 s2 = 0ull, q2 = 0ull;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint32_t wv[4] = {ry[r][v].x, ry[r][v].y, ry[r][v].z, ry[r][v].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const u64 f = lg_unpack2<TI>(wv[e]);
                    s2 = f2_add(s2, f);
                    q2 = f2_fma(f, f, q2);
                }
            }
            float s0, s1, q0, q1;
            f2_unpack(s2, s0, s1);
            f2_unpack(q2, q0, q1);
            float sm = s0 + s1, sq = q0 + q1;
            sm += __shfl_xor_sync(0xffffffffu, sm, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sm += __shfl_xor_sync(0xffffffffu, sm, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            const float mean = sm * (1.f / LG_D);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / LG_D) - mean * mean, 0.f) + eps);
            nrm_a[r] = f2_pack(rstd, rstd);
            nrm_c[r] = f2_pack(-mean * rstd, -mean * rstd);
        }
        // A fragments: k-tiles 2 v, 2 v + 1 = the two halves of the thread's v-th 16-byte piece, both rows; three packed FMAs per pair
        uint32_t afr[LG_KT][4];
        const float* lb = local + (long)b * LG_D + 8 * t;
#pragma unroll
        for (int v = 0; v < 4; ++v) {                          // 8 channels per 16-byte vector = k-tiles 2 v and 2 v + 1
            const ulonglong2 g0 = *reinterpret_cast<const ulonglong2*>(s_gamma + 32 * v + 8 * t), g1 = *reinterpret_cast<const ulonglong2*>(s_gamma + 32 * v + 8 * t + 4);
            const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(s_beta + 32 * v + 8 * t), b1 = *reinterpret_cast<const ulonglong2*>(s_beta + 32 * v + 8 * t + 4);
            const ulonglong2 l0 = __ldg(reinterpret_cast<const ulonglong2*>(lb + 32 * v)), l1 = __ldg(reinterpret_cast<const ulonglong2*>(lb + 32 * v + 4));
            const u64 gm[4] = {g0.x, g0.y, g1.x, g1.y}, bt[4] = {b0.x, b0.y, b1.x, b1.y}, lc[4] = {l0.x, l0.y, l1.x, l1.y};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t wy[4] = {ry[r][v].x, ry[r][v].y, ry[r][v].z, ry[r][v].w};
                const uint32_t wz[4] = {rz[r][v].x, rz[r][v].y, rz[r][v].z, rz[r][v].w};
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const u64 n2 = f2_fma(f2_fma(lg_unpack2<TI>(wy[e]), nrm_a[r], nrm_c[r]), gm[e], bt[e]);
                    float o0, o1;
                    f2_unpack(f2_fma(n2, lg_unpack2<TI>(wz[e]), lc[e]), o0, o1);
                    pk[e] = lg_pack<TI>(o0, o1);
                }
                // m16n8k16 A fragment: {row g: k 2t..2t+1, row g+8: same, row g: k 2t+8..2t+9, row g+8: same}
                afr[2 * v][r] = pk[0]; afr[2 * v][2 + r] = pk[1];
                afr[2 * v + 1][r] = pk[2]; afr[2 * v + 1][2 + r] = pk[3];
            }
        }
        float acc[LG_NT][4];
#pragma unroll
        for (int nn = 0; nn < LG_NT; ++nn) { acc[nn][0] = acc[nn][1] = acc[nn][2] = acc[nn][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < LG_KT; ++j)
#pragma unroll
            for (int nn = 0; nn < LG_NT; ++nn) {
                const unsigned long long bw = s_w[(nn * LG_KT + j) * 32 + lane];
                mma_16816<TI>(acc[nn], afr[j], (uint32_t)bw, (uint32_t)(bw >> 32));
            }
        // Epilogue through the warp's shared-memory tile.  A C fragment holds two adjacent columns per lane: stored from there a
        // warp instruction wrote eight 16-byte pieces and every 32-byte sector arrived in two halves (measured: the kernel at
        // 2.8 TB/s).  gate * acc goes to the tile in fp32 (row pitch 72 floats: the eight rows of a store land in different bank
        // groups), and is read back as full rows: a lane adds 8 addend columns (one LDG.128) and writes 16 contiguous bytes, a warp
        // instruction writes four complete 128-byte rows.
        const float* gp = gate + (long)b * gate_stride + 2 * t;
        float* tile_s = s_out + warp * (16 * LG_OP);
        __syncwarp();                                          // the previous tile's read-back is finished
#pragma unroll
        for (int nn = 0; nn < LG_NT; ++nn) {
            const float2 gv = __ldg(reinterpret_cast<const float2*>(gp + 8 * nn));
            *reinterpret_cast<float2*>(tile_s + g * LG_OP + 8 * nn + 2 * t) = make_float2(gv.x * acc[nn][0], gv.y * acc[nn][1]);
            *reinterpret_cast<float2*>(tile_s + (g + 8) * LG_OP + 8 * nn + 2 * t) = make_float2(gv.x * acc[nn][2], gv.y * acc[nn][3]);
        }
        __syncwarp();
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
            const int rr = ps * 4 + (lane >> 3), cc = (lane & 7) * 8;
            const long o = (long)(tile * 16 + rr) * LG_C + cc;
            const uint4 av = *reinterpret_cast<const uint4*>(addend + o);
            const float4 f0 = *reinterpret_cast<const float4*>(tile_s + rr * LG_OP + cc), f1 = *reinterpret_cast<const float4*>(tile_s + rr * LG_OP + cc + 4);
            const float2 a0 = lg_unpack<TO>(av.x), a1 = lg_unpack<TO>(av.y), a2 = lg_unpack<TO>(av.z), a3 = lg_unpack<TO>(av.w);
            uint4 ov;
            ov.x = lg_pack<TO>(f0.x + a0.x, f0.y + a0.y);
            ov.y = lg_pack<TO>(f0.z + a1.x, f0.w + a1.y);
            ov.z = lg_pack<TO>(f1.x + a2.x, f1.y + a2.y);
            ov.w = lg_pack<TO>(f1.z + a3.x, f1.w + a3.y);
            *reinterpret_cast<uint4*>(out + o) = ov;
        }
    }
}

template <typename TI, typename TO>
int lg_launch(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta, const float* local, const void* w,
              const float* gate, int gate_stride, const void* addend, void* out, int B, int P, float eps, cudaStream_t st) {
    const long tiles = (long)B * P / 16;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long want = (tiles + LG_WARPS - 1) / LG_WARPS;
    const int grid = (int)(want < 2L * sms ? want : 2L * sms);
    const size_t smem = (size_t)LG_WARPS * 16 * LG_OP * sizeof(float);         // + 17 KB static: above the 48 KB default
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ln_gate_out_proj_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    ln_gate_out_proj_kernel<TI, TO><<<grid, LG_WARPS * 32, smem, st>>>((const TI*)y, (const TI*)xz, ld, z_off, gamma, beta, local, (const TI*)w,
                                                                    gate, gate_stride, (const TO*)addend, (TO*)out, (unsigned)tiles,
                                                                    (unsigned)P, eps);
    FD_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int fd_ln_gate_out_proj_supported(int P, int D, int Cout, int ld, int z_off, int io_dtype, int out_dtype) {
    const bool t16 = (io_dtype == FD_BF16 || io_dtype == FD_F16) && (out_dtype == FD_BF16 || out_dtype == FD_F16);
    return t16 && D == LG_D && Cout == LG_C && P > 0 && P % 16 == 0 && ld % 8 == 0 && z_off % 8 == 0 && ld >= z_off + D;
}

extern "C" int fd_ln_gate_out_proj(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta,
                                   const float* local, const void* w, const float* gate, int gate_stride, const void* addend, void* out,
                                   int B, int P, int D, int Cout, float eps, int io_dtype, int out_dtype, cudaStream_t stream) {
    if (!y || !xz || !gamma || !beta || !local || !w || !gate || !addend || !out || B <= 0 || P <= 0) return FD_ERR_BAD_ARGUMENT;
    if (!fd_ln_gate_out_proj_supported(P, D, Cout, ld, z_off, io_dtype, out_dtype)) return FD_ERR_UNSUPPORTED;
    if ((long)B * P >= (1L << 31) || gate_stride % 2 ||
        (((uintptr_t)y | (uintptr_t)xz | (uintptr_t)local | (uintptr_t)w) & 15) || (((uintptr_t)gate) & 7) ||
        (((uintptr_t)addend | (uintptr_t)out) & 15))
        return FD_ERR_UNSUPPORTED;
#define LG_CASE(TIV, TI, TOV, TO) \
    if (io_dtype == TIV && out_dtype == TOV) \
        return lg_launch<TI, TO>(y, xz, ld, z_off, gamma, beta, local, w, gate, gate_stride, addend, out, B, P, eps, stream);
    LG_CASE(FD_BF16, __nv_bfloat16, FD_BF16, __nv_bfloat16) LG_CASE(FD_BF16, __nv_bfloat16, FD_F16, __half)
    LG_CASE(FD_F16, __half, FD_F16, __half) LG_CASE(FD_F16, __half, FD_BF16, __nv_bfloat16)
#undef LG_CASE
    return FD_ERR_UNSUPPORTED;
}
