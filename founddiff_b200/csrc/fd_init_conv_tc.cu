// init_conv (7x7, pad 3, over cat(x_t, x_input); src/DADiff.py:558, 700 + torch.cat :1160) on the tensor cores.
//
// The CUDA-core kernel (fd_sampler.cu) is FMA-bound: 6272 MACs per pixel, 1.2 ms per launch at 16 x 512^2.  Here the
// convolution is an implicit GEMM with M = 128 pixels (an 8 x 16 patch), N = 64 output channels and K = 2*7*7 = 98 taps:
//   * the two fp32 images are NOT rounded to 16 bit: every pixel is split into hi = fp16(x) and lo = fp16(x - hi)
//     (together ~22 mantissa bits) and both parts are multiplied by the same fp16 weights, K = [hi | pad | lo | pad]
//     = 4 blocks of 64 (256), accumulated in fp32 in TMEM;
//   * the A tile is an explicit im2col built by the CTA's threads in shared memory, directly in the canonical K-major
//     SWIZZLE_128B layout (one warp writes one 128-byte row per instruction: conflict-free), from a 14 x 22 x 2 patch of
//     the inputs that is split into fp16 hi / lo halves ONCE when it is loaded.  Tap rows are padded to 8 (K index
//     = ci*56 + ky*8 + kx, zero weight at kx = 7) so that a lane's tap pair (kx, kx+1) is two adjacent patch elements:
//     with a second copy of the patch shifted by one element every pair is one aligned LDS.32, and the im2col inner loop
//     is LDS.32 + STS.32 per part.  Generic-proxy writes are made visible to the tensor core with fence.proxy.async;
//   * weights (64 x 256 fp16 = 32 KB, packed by the host) stay resident in shared memory for the persistent CTA;
//   * 16 tcgen05.mma (M128 N64 K16) per tile, issued by one thread; the 8 warps then read the accumulator
//     (tcgen05.ld), add the bias and store 32 bytes per lane.
// Two CTAs per SM (100 KB of shared memory each) overlap one CTA's im2col with the other's MMA / epilogue.
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int IT_TH = 8, IT_TW = 16;                 // output tile (pixels)
constexpr int IT_PH = IT_TH + 6, IT_PW = IT_TW + 6;  // input patch
constexpr int IT_N = 64;                             // output channels
constexpr int IT_KB = 4;                             // K blocks of 64: hi[0:64), hi[64:128), lo[0:64), lo[64:128)
constexpr int IT_TAPS = 112;                         // 2 channels x 7 rows x 8 (7 taps + 1 zero-weight pad) per part
constexpr int IT_PS = 24;                            // patch row pitch (halfwords): 12 words, so that the 28 + 4 tap pairs a warp
constexpr int IT_CH = IT_PH * IT_PS + 24;            // reads per im2col row fall into 32 different banks (channel pitch 180 words)
constexpr int IT_COPY = 2 * IT_CH;                   // halfwords per patch copy (2 channels)
constexpr int IT_PF = (2 * IT_PH * IT_PW + 255) / 256;   // patch elements per thread

FD_DEVINL uint32_t it_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FD_DEVINL uint64_t it_desc(uint32_t saddr) {         // K-major SWIZZLE_128B, SBO = 1024 B (see fd_conv_tc.cu)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// explicit shared-space accesses: the 1024-byte align-up of the dynamic shared memory base goes through an integer cast,
// after which the compiler only knows a generic pointer (ST.E / LD.E with 64-bit addresses instead of STS / LDS)
FD_DEVINL void it_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
FD_DEVINL uint32_t it_pack_half2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256, 2) init_conv_tc_kernel(const float* __restrict__ x_t, const float* __restrict__ x_in,
                                                              const __half* __restrict__ w16, const float* __restrict__ bias,
                                                              T* __restrict__ out, int H, int W, int tiles_w, int tiles_per_img,
                                                              int total_tiles) {
    extern __shared__ __align__(1024) uint8_t it_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)it_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;                                   // IT_KB x (64 rows x 128 B)
    uint8_t* sA = smem + IT_KB * IT_N * 128;              // IT_KB x (128 rows x 128 B)
    __half* s_patch = (__half*)(sA + IT_KB * 128 * 128);  // [part: hi, lo][copy: 0, shifted by 1][channel][IT_CH] fp16
    uint64_t* bar = (uint64_t*)(s_patch + 4 * IT_COPY + 4);
    uint32_t* tmem_slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sA_u = it_smem_u32(sA), sP_u = it_smem_u32(s_patch);

    // resident weights: 16-byte chunks into the swizzled K-major layout
    for (int c = tid; c < IT_N * 32; c += 256) {
        const int n = c >> 5, j = c & 31, kb = j >> 3, jj = j & 7;
        *reinterpret_cast<uint4*>(sB + kb * (IT_N * 128) + n * 128 + ((jj ^ (n & 7)) << 4)) =
            __ldg(reinterpret_cast<const uint4*>(w16) + c);
    }
    // zero the A tile and the patch once: K positions 112..127 of both halves and the patch pad slots are never written again
    for (int c = tid; c < IT_KB * 128 * 8; c += 256) *reinterpret_cast<uint4*>(sA + c * 16) = make_uint4(0, 0, 0, 0);
    for (int c = tid; c < 4 * IT_COPY + 4; c += 256) s_patch[c] = __float2half(0.f);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(it_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(it_smem_u32(tmem_slot)), "r"(IT_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // per-lane tap-pair offset (halfwords, even) inside a patch copy for the two K blocks of a part
    uint32_t off[2];
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
        const int k = kb * 64 + 2 * lane;
        const int kk = k < IT_TAPS ? k : 0;
        const int ci = kk / 56, r = kk % 56;
        off[kb] = (uint32_t)(ci * IT_CH + (r / 8) * IT_PS + (r % 8));
    }
    // accumulator row of this thread in the epilogue: TMEM lane quarter = warp % 4, 32 columns per warp half
    const int q4 = warp & 3, ch = warp >> 2;
    const int m_epi = q4 * 32 + lane;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(IT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
    float bv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bv[j] = __ldg(bias + ch * 32 + j);

    // the patch of the NEXT tile is fetched into registers right after this tile's MMAs are issued, so the global-load latency
    // (three dependent-looking loads per thread were ~2 us of a 5 us tile) hides behind the MMA wait and the epilogue
    float pv[IT_PF];
    auto fetch_patch = [&](int tl) {
        const int fb = tl / tiles_per_img, ft = tl - fb * tiles_per_img;
        const int fy0 = (ft / tiles_w) * IT_TH, fx0 = (ft % tiles_w) * IT_TW;
#pragma unroll
        for (int u = 0; u < IT_PF; ++u) {
            const int i = tid + u * 256;
            const int ci = i / (IT_PH * IT_PW), r = i % (IT_PH * IT_PW);
            const int yy = fy0 - 3 + r / IT_PW, xx = fx0 - 3 + r % IT_PW;
            const bool ok = i < 2 * IT_PH * IT_PW && yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float* src = (ci ? x_in : x_t) + ((long)fb * H + (ok ? yy : 0)) * W + (ok ? xx : 0);
            const float v = __ldg(ci < 2 ? src : x_t);       // unconditional (clamped) load, masked below
            pv[u] = ok ? v : 0.f;
        }
    };
    if ((int)blockIdx.x < total_tiles) fetch_patch(blockIdx.x);
    uint32_t parity = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, t = tile - b * tiles_per_img;
        const int y0 = (t / tiles_w) * IT_TH, x0 = (t % tiles_w) * IT_TW;
        // 1. input patch (values prefetched during the previous tile), split into fp16 hi / lo, each stored in the plain and
        //    the shifted copy
#pragma unroll
        for (int u = 0; u < IT_PF; ++u) {
            const int i = tid + u * 256;
            if (i >= 2 * IT_PH * IT_PW) break;
            const int ci = i / (IT_PH * IT_PW), r = i % (IT_PH * IT_PW);
            const float v = pv[u];
            const uint32_t hh = it_pack_half2(v, 0.f) & 0xffffu;
            const float hf = __half2float(*reinterpret_cast<const __half*>(&hh));
            const uint32_t ll = it_pack_half2(v - hf, 0.f) & 0xffffu;
            const int pr = (r / IT_PW) * IT_PS + r % IT_PW;                        // pitched index inside the channel
            const uint32_t e0 = sP_u + 2u * (uint32_t)(ci * IT_CH + pr);           // copy 0, element pr
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(e0), "h"((unsigned short)hh) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(e0 + 2u * 2u * IT_COPY), "h"((unsigned short)ll) : "memory");
            if (pr > 0) {                                                           // copy 1 holds element pr at index pr - 1
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(e0 + 2u * IT_COPY - 2u), "h"((unsigned short)hh) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(e0 + 2u * 3u * IT_COPY - 2u), "h"((unsigned short)ll) : "memory");
            }
        }
        __syncthreads();
        // 2. im2col: warp w writes rows w, w+8, ...; lane = pair of taps (4 bytes) of the 128-byte row.  All taps of a row have
        //    the parity of its pixel column, which selects the patch copy in which the pair is 4-byte aligned.
#pragma unroll 4
        for (int m = warp; m < 128; m += 8) {
            const uint32_t px = (uint32_t)(m & 15), cpy = px & 1u;
            const uint32_t src = sP_u + 2u * (cpy * IT_COPY + (uint32_t)(m >> 4) * IT_PS + px - cpy);
            const uint32_t row_off = (uint32_t)m * 128u + ((uint32_t)((lane >> 2) ^ (m & 7)) << 4) + ((uint32_t)(lane & 3) << 2);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                if (kb == 1 && lane >= (IT_TAPS - 64) / 2) continue;       // K 112.. of the second block stays zero
                uint32_t hi, lo;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hi) : "r"(src + 2u * off[kb]));
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lo) : "r"(src + 2u * off[kb] + 2u * 2u * IT_COPY));
                it_sts32(sA_u + (uint32_t)(kb * (128 * 128)) + row_off, hi);
                it_sts32(sA_u + (uint32_t)((kb + 2) * (128 * 128)) + row_off, lo);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor core (async proxy)
        __syncthreads();
        // 3. 16 MMAs: warp 0 stays converged, one elected lane issues (a lone divergent thread needs ~25 instructions per MMA)
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da0 = it_desc(it_smem_u32(sA)), db0 = it_desc(it_smem_u32(sB));
            uint32_t leader;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(leader));
            if (leader) {
#pragma unroll
                for (int kb = 0; kb < IT_KB; ++kb) {
                    const uint64_t da = da0 + (uint64_t)(kb * (128 * 128 / 16)), db = db0 + (uint64_t)(kb * (IT_N * 128 / 16));
                    asm volatile(
                        "{\n\t.reg .pred p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\tadd.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
                        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}" ::"r"(tmem),
                        "l"(da), "l"(db), "r"(idesc), "r"(kb ? 1u : 0u)
                        : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(it_smem_u32(bar)) : "memory");
            }
            __syncwarp();
        }
        if (tile + (int)gridDim.x < total_tiles) fetch_patch(tile + gridDim.x);
        // 4. wait for the accumulator, epilogue
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done)
                             : "r"(it_smem_u32(bar)), "r"(parity)
                             : "memory");
            }
            parity ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int oy = y0 + (m_epi >> 4), ox = x0 + (m_epi & 15);
        T* orow = out + (((long)b * H + oy) * W + ox) * IT_N + ch * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ch * 32 + c)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a0 = __uint_as_float(r[2 * i]) + bv[c + 2 * i], a1 = __uint_as_float(r[2 * i + 1]) + bv[c + 2 * i + 1];
                if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
                    w[i] = *reinterpret_cast<uint32_t*>(&h2);
                } else {
                    w[i] = it_pack_half2(a0, a1);
                }
            }
            if (oy < H && ox < W)
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + c), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                             "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                             : "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                 // accumulator drained, A tile and patch free for the next tile
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(IT_N));
}

}  // namespace

extern "C" int fd_init_conv7x7_tc(const float* x_t, const float* x_input, const void* w16, const float* bias, void* out, int B,
                                  int H, int W, int Cout, int dtype, cudaStream_t stream) {
    if (!x_t || !x_input || !w16 || !bias || !out || B <= 0 || H <= 0 || W <= 0) return FD_ERR_BAD_ARGUMENT;
    if (Cout != IT_N || H % IT_TH || W % IT_TW || (dtype != FD_BF16 && dtype != FD_F16)) return FD_ERR_UNSUPPORTED;
    if ((((uintptr_t)w16 | (uintptr_t)out) & 31)) return FD_ERR_UNSUPPORTED;
    const int tiles_w = W / IT_TW, tiles_per_img = (H / IT_TH) * tiles_w;
    const long total = (long)B * tiles_per_img;
    if (total >= (1L << 31)) return FD_ERR_UNSUPPORTED;
    const size_t smem = 1024 + (size_t)IT_KB * IT_N * 128 + (size_t)IT_KB * 128 * 128 + (4 * IT_COPY + 4) * sizeof(__half) + 64;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(total < 2L * sms ? total : 2L * sms);
    if (dtype == FD_BF16) {
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(init_conv_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr = true;
        }
        init_conv_tc_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(x_t, x_input, (const __half*)w16, bias, (__nv_bfloat16*)out, H, W,
                                                                        tiles_w, tiles_per_img, (int)total);
    } else {
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(init_conv_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr = true;
        }
        init_conv_tc_kernel<__half><<<grid, 256, smem, stream>>>(x_t, x_input, (const __half*)w16, bias, (__half*)out, H, W, tiles_w,
                                                                 tiles_per_img, (int)total);
    }
    FD_LAUNCH_CHECK();
    return 0;
}
