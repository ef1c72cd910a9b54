// Implicit-GEMM convolution / 1x1 GEMM on the 5th-generation tensor cores (sm_100a):
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-stage mbarrier ring,
//   * tcgen05.mma (kind::f16, cta_group::1, M=128, N=64/128/256, K=16) issued by one elected thread,
//   * fp32 accumulators in TMEM, read back with tcgen05.ld by four epilogue warps that apply the fused epilogue of
//     fd_conv_params (bias, SiLU on a channel range, adaLN gate, residual addend, GroupNorm partial sums).
//
// GEMM view (same as fd_conv_simt.cu): M = output pixels, N = Cout, K = taps x Cin.  An M tile is an 8x16 patch of
// output pixels of ONE sample; for tap (kh, kw) and a 64-channel slice the A tile is ONE TMA box
// {64 ch, 16 cols, 8 rows, 1 sample} at element offset (kw - pad, kh - pad): out-of-bounds elements are zero-filled
// by the TMA unit, which is exactly the convolution's zero padding.  The box lands in shared memory as 128 rows of
// 128 bytes with the 128B swizzle = the canonical K-major UMMA operand layout, so no thread ever touches A or B.
//   stride 2 (Downsample 4x4 s2): the same box with elementStrides {1,2,2,1}.
//   nearest x2 upsample + 3x3   : decomposed into 4 output phases, each a 2x2 convolution over the LOW-resolution
//                                 input with pre-summed weights (host-packed `weight_up4`), 2.25x fewer FLOPs.
//   torch.cat inputs            : two tensor maps; K blocks switch map at c0.
//   per-sample weights (W_eff)  : 3-D weight map {K, Cout, batch}.
//
// HALO mode (stride-1 convolutions with more than one tap): re-fetching the shifted A box for every tap made the
// full-resolution 3x3 convolutions L2-bandwidth bound (9 x 16 KB per 64-channel slice and tile).  Instead the M tile
// becomes 16 rows x 8 pixels and ONE TMA box {64 ch, 8+KW-1, 16+KH-1} (the halo tile, 23 KB for 3x3) is loaded per
// 64-channel slice; every tap's A operand is a *shifted window* of it, expressed purely in the UMMA descriptor:
// start = halo + ((kh*HT_W + kw) * 128 B), SBO = HT_W * 128 B (pitch of a halo row), base_offset 0 — the 128B swizzle
// is a function of the absolute shared-memory address, so the XOR phase TMA wrote matches what the MMA unit reads
// (tools/probes/umma_halo_probe.cu verifies this on the device for all nine shifts).  Weights go through their own
// ring, or stay resident in shared memory for the whole persistent CTA when they fit (full-resolution 64-channel
// convolutions: 72-144 KB), which removes the second largest L2 stream.
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#include "fd_common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (pixels per tile)
constexpr int TILE_H = 8, TILE_W = 16;
constexpr int BK = 64;           // K block = one 128-byte swizzle atom of 16-bit elements
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 8;
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_STAT_WARPS = 2;                     // LayerNorm row statistics of the staged A tile (ln_v != NULL), 2 rows per lane
constexpr int NTHREADS = (2 + NUM_EPI_WARPS) * 32;    // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, 16 epilogue warps
constexpr int NTHREADS_LN = NTHREADS + NUM_STAT_WARPS * 32;   // + the statistics warp of the LayerNorm-fold instantiation

// n / d for 0 <= n with n * d < 2^32 as one multiply-high (host-precomputed ceil(2^32 / d); d == 1 handled apart)
struct FastDiv {
    uint32_t mul, d;
};
FD_DEVINL uint32_t fdiv(uint32_t n, const FastDiv& f) { return f.d == 1 ? n : __umulhi(n, f.mul); }

struct TcParams {
    fd_conv_params p;
    int Hout, Wout;
    int tiles_h, tiles_w;
    int BN;                 // N tile (64, 128 or 256)
    int n_tiles;            // Cout / BN
    int phases;             // 4 for upsample, else 1
    int taps_h, taps_w;     // taps per phase
    int kblocks0, kblocks1; // 64-channel blocks of src0 / src1
    int fmt;                // 0 = f16, 1 = bf16 (UMMA a/b format)
    int stages;             // depth of the smem ring (box mode: A+B stages; halo mode: A halo stages)
    int total_tiles;
    int tile_h, tile_w;     // M tile = tile_h x tile_w output pixels (8x16 box mode, 16x8 halo mode)
    int halo;               // 1: halo mode
    int ht_h, ht_w;         // halo tile extent (pixels)
    int a_stage_bytes;      // halo mode: bytes per A stage (1024-aligned)
    int stages_b;           // halo mode: weight ring depth (ignored when b_stationary)
    int b_stationary;       // halo mode: all weight tiles resident in shared memory
    FastDiv dv_n, dv_w, dv_h, dv_p;   // tile index -> (n tile, tile column, tile row, phase, sample)
    int tile_w_shift, cpg_shift;
};

struct TileCoord {
    int nt, tw, th, phase, b;
};
FD_DEVINL TileCoord decode_tile(const TcParams& q, int tile) {
    TileCoord c;
    uint32_t t = (uint32_t)tile, u;
    u = fdiv(t, q.dv_n); c.nt = (int)(t - u * q.dv_n.d); t = u;
    u = fdiv(t, q.dv_w); c.tw = (int)(t - u * q.dv_w.d); t = u;
    u = fdiv(t, q.dv_h); c.th = (int)(t - u * q.dv_h.d); t = u;
    u = fdiv(t, q.dv_p); c.phase = (int)(t - u * q.dv_p.d);
    c.b = (int)u;
    return c;
}

// ---------------------------------------------------------------------------------------------------- PTX helpers
FD_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

FD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FD_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
FD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Waits that are expected to be long (producer on a full ring, epilogue on the next accumulator): back off so that the
// spinning warps do not take issue slots from the working ones.
FD_DEVINL void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(40);
}
// epilogue warps waiting for the next accumulator: 16 warps, long waits, and they share schedulers with the MMA warp
FD_DEVINL void mbar_wait_long(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(200);
}
FD_DEVINL void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
FD_DEVINL void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
FD_DEVINL void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the four K = 16 steps of one 64-channel K block: descriptors advance by 32 B (2 descriptor units) per step
FD_DEVINL void umma_f16_x4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
        "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\tadd.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\t"
        "add.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all taps of one 64-channel slice against resident weights, fully unrolled: the A window offsets are immediates and the
// issue stream is nothing but descriptor adds and MMAs (the issuing thread is the critical path for N = 64 tiles)
template <int TAPS_H, int TAPS_W, int HT_W>
FD_DEVINL void umma_taps_resident(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t tap_stride, uint32_t idesc, uint32_t accumulate) {
#pragma unroll
    for (int kh = 0; kh < TAPS_H; ++kh)
#pragma unroll
        for (int kw = 0; kw < TAPS_W; ++kw) {
            umma_f16_x4(tmem_d, da + (uint64_t)((kh * HT_W + kw) * 8), db, idesc, (kh | kw) ? 1u : accumulate);
            db += tap_stride;
        }
}
FD_DEVINL bool elect_one() {      // one lane of the (converged) warp; the same lane every time
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
FD_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FD_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO (ignored for swizzled K-major) = 1, SBO = 1024 B (8 rows x 128 B) >> 4, version 1 (sm_100), layout 2 (128B).
FD_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes = 1024) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c_format f32 (bit 4), a/b format (bits 7, 10), K-major both, N>>3 (bit 17), M>>4 (bit 24)
FD_DEVINL uint32_t make_idesc(int fmt, int n) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------
// 32-byte (16 x 16-bit) vector global access: one full sector per thread per instruction (LDG/STG.E.ENL2.256)
template <typename T> FD_DEVINL void store16(T* dst, const float (&v)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if constexpr (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        } else {
            __half2 h = fd_floats2half2_sat(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
template <typename T> FD_DEVINL void load16(const T* src, float (&v)[16]) {
    uint32_t w[8];
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(src));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float2 f;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[i]));
        else f = __half22float2(*reinterpret_cast<__half2*>(&w[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
// raw 32-byte addend vector: requested early (before the accumulator is waited for), unpacked when it is added
FD_DEVINL void load16_raw(const void* src, uint32_t (&w)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(src));
}
template <typename T> FD_DEVINL void add16_raw(const uint32_t (&w)[8], float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float2 f;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        else f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] += f.x;
        v[2 * i + 1] += f.y;
    }
}
FD_DEVINL float fast_silu(float x) {      // x * sigmoid(x) = 0.5 x (1 + tanh(x / 2)): ONE MUFU op (tanh.approx, 2^-11 relative)
    float t;                               // instead of EX2 + RCP; the result is stored in a 16-bit type (2^-9 / 2^-12) anyway
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// Persistent, warp-specialised kernel.  Roles: warp 0 = TMA producer, warp 1 = MMA issuer (owns TMEM), warps 2..17 =
// epilogue (lane quarter = warp % 4, column group = (warp - 2) / 4).  Two TMEM accumulator buffers let the epilogue of
// tile i overlap the loads + MMAs of tile i+1; every role walks the same static tile sequence
// tile = blockIdx.x + i * gridDim.x (N tile fastest, so neighbouring CTAs share the A tile in L2).
// LNMODE is a template parameter so that the plain instantiation keeps the register allocation and schedule it was tuned
// with (as a run-time flag the fold cost every other 64-wide GEMM 13 - 15 %).  0: no fold; 1: fold, row statistics by the two
// statistics warps from the staged operand tile; 2: fold, per-pixel rstd read from p.ln_rstd (fd_row_rstd) — no statistics warps,
// the stage protocol of the plain kernel (measured: 64 -> 256 at 16 x 512^2 715 us with the statistics warps, 561 us plain).
template <typename T, int LNMODE>
__global__ void __launch_bounds__(LNMODE == 1 ? NTHREADS_LN : NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const TcParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by POINTER arithmetic on the shared array: rounding the address up through an integer cast made every
    // later access of this memory a generic load / store (long-scoreboard latency: the LayerNorm-fold epilogue's ln_v reads alone
    // cost the 64 -> 256 GEMM 130 us of its 690)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int BN = q.BN;
    const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)BN * BK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int stages = q.stages;
    const int kb_per_tap = q.kblocks0 + q.kblocks1;
    const int num_kb = q.taps_h * q.taps_w * kb_per_tap;
    // halo mode: [weight slots][A halo stages]; box mode: [stages x (A, B)]
    const int b_slots = q.b_stationary ? num_kb : q.stages_b;
    uint8_t* sB_halo = smem;
    uint8_t* sA_halo = smem + (size_t)b_slots * b_bytes;
    const size_t data_bytes = q.halo ? (size_t)b_slots * b_bytes + (size_t)stages * q.a_stage_bytes : (size_t)stages * stage_bytes;
    uint64_t* full_bar = (uint64_t*)(smem + data_bytes);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* bfull_bar = empty_bar + MAX_STAGES;     // halo mode: weight ring
    uint64_t* bempty_bar = bfull_bar + MAX_STAGES;
    uint64_t* wfull_bar = bempty_bar + MAX_STAGES;    // halo mode, stationary weights: all tiles landed
    uint64_t* tfull_bar = wfull_bar + 1;              // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
    uint64_t* sfull_bar = tempty_bar + 2;             // [2] LayerNorm row statistics of the tile written (ln fold)
    uint32_t* tmem_slot = (uint32_t*)(sfull_bar + 2);
    float* s_gn = (float*)(tmem_slot + 2);            // [2 flush parities][16 epilogue warps][2][8] GroupNorm partial sums
    float2* s_ln = (float2*)(s_gn + 2 * NUM_EPI_WARPS * 16);   // [2 accumulator buffers][128 rows] (mean, rstd) of the input rows
    float* s_lnv = (float*)(s_ln + 2 * BM);           // [Cout <= 1024] ln_v of the sample the block is working on (ln fold)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const fd_conv_params& p = q.p;
    constexpr bool ln_fold = LNMODE != 0;             // LayerNorm of the input rows folded into this GEMM (box mode, 1x1)
    constexpr bool ln_stats = LNMODE == 1;            // statistics warps in this block
    const int total_tiles = q.total_tiles;
    const uint32_t tmem_cols = 2 * BN <= 128 ? 128u : 2 * BN <= 256 ? 256u : 512u;       // two accumulators, power-of-two allocation

    if (threadIdx.x == 0) {
        // with the LayerNorm fold a stage is released by the MMA commit AND by the statistics warps that read its A tile
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], ln_stats ? 1 + NUM_STAT_WARPS : 1); }
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&bfull_bar[s], 1); mbar_init(&bempty_bar[s], 1); }
        mbar_init(wfull_bar, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], NUM_EPI_WARPS); mbar_init(&sfull_bar[s], NUM_STAT_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 2 * NUM_EPI_WARPS * 16) s_gn[threadIdx.x] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            int stage = 0, bstage = 0;
            uint32_t ph = 0, bph = 0;
            if (q.halo && q.b_stationary) {            // weights are tile-invariant: load every K block once
                mbar_expect_tx(wfull_bar, (uint32_t)num_kb * b_bytes);
                for (int kb = 0; kb < num_kb; ++kb) tma_load_3d(sB_halo + (size_t)kb * b_bytes, &map_w, wfull_bar, kb * BK, 0, 0);
            }
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(q, tile);
                const int phase = tc.phase, b = tc.b;
                const int ho0 = tc.th * q.tile_h, wo0 = tc.tw * q.tile_w, n0 = tc.nt * BN;
                int hbase, wbase;
                if (p.upsample) { hbase = ho0 - 1 + (phase >> 1); wbase = wo0 - 1 + (phase & 1); }
                else { hbase = ho0 * p.stride - p.pad; wbase = wo0 * p.stride - p.pad; }
                const int wbatch = p.per_batch_weight ? b : phase;
                if (q.halo) {
                    const uint32_t halo_bytes = (uint32_t)(q.ht_h * q.ht_w) * 128u;
                    for (int cb = 0; cb < kb_per_tap; ++cb) {
                        mbar_wait_backoff(&empty_bar[stage], ph ^ 1);
                        mbar_expect_tx(&full_bar[stage], halo_bytes);
                        uint8_t* sa = sA_halo + (size_t)stage * q.a_stage_bytes;
                        if (cb < q.kblocks0) tma_load_4d(sa, &map_a0, &full_bar[stage], cb * BK, wbase, hbase, b);
                        else                 tma_load_4d(sa, &map_a1, &full_bar[stage], (cb - q.kblocks0) * BK, wbase, hbase, b);
                        if (++stage == stages) { stage = 0; ph ^= 1; }
                        if (!q.b_stationary) {
                            for (int tap = 0; tap < q.taps_h * q.taps_w; ++tap) {
                                mbar_wait_backoff(&bempty_bar[bstage], bph ^ 1);
                                mbar_expect_tx(&bfull_bar[bstage], b_bytes);
                                tma_load_3d(sB_halo + (size_t)bstage * b_bytes, &map_w, &bfull_bar[bstage], (tap * kb_per_tap + cb) * BK, n0, wbatch);
                                if (++bstage == q.stages_b) { bstage = 0; bph ^= 1; }
                            }
                        }
                    }
                    continue;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / kb_per_tap, cb = kb % kb_per_tap;
                    const int kh = tap / q.taps_w, kw = tap % q.taps_w;
                    mbar_wait_backoff(&empty_bar[stage], ph ^ 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], stage_bytes);
                    if (cb < q.kblocks0) tma_load_4d(sa, &map_a0, &full_bar[stage], cb * BK, wbase + kw, hbase + kh, b);
                    else                 tma_load_4d(sa, &map_a1, &full_bar[stage], (cb - q.kblocks0) * BK, wbase + kw, hbase + kh, b);
                    tma_load_3d(sa + a_bytes, &map_w, &full_bar[stage], kb * BK, n0, wbatch);
                    if (++stage == stages) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // The whole warp walks the (warp-uniform) loop and polls the barriers; one elected lane issues the MMAs and
        // commits.  Keeping the control flow converged lets the compiler hold descriptors in uniform registers — as a
        // single divergent thread the issue loop cost ~25 instructions per MMA and, for N = 64 tiles, was the
        // critical path of the kernel.
        {
            const uint32_t idesc = make_idesc(q.fmt, BN);
            int stage = 0, bstage = 0;
            uint32_t ph = 0, bph = 0;
            int it = 0;
            if (q.halo && q.b_stationary) mbar_wait(wfull_bar, 0);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
                uint32_t acc = 0;
                if (q.halo) {
                    const uint32_t sbo_a = (uint32_t)q.ht_w * 128u;
                    const uint32_t sB0 = smem_u32(sB_halo);
                    const uint32_t tap_stride = (uint32_t)kb_per_tap * (b_bytes >> 4);        // resident weights: descriptor units
                    const uint32_t row_skip = (uint32_t)(q.ht_w - q.taps_w) * 8u;
                    for (int cb = 0; cb < kb_per_tap; ++cb) {
                        mbar_wait(&full_bar[stage], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const int bstage_in = bstage;
                        const uint32_t bph_in = bph;
                        if (elect_one()) {      // one thread runs the whole tap loop of this 64-channel slice
                            uint64_t da = make_smem_desc(smem_u32(sA_halo) + (uint32_t)stage * (uint32_t)q.a_stage_bytes, sbo_a);
                            uint64_t db = make_smem_desc(sB0 + (uint32_t)cb * b_bytes);       // resident: slot of (tap 0, cb)
                            if (q.b_stationary && q.taps_h == 3 && q.taps_w == 3) {
                                umma_taps_resident<3, 3, 10>(tacc, da, db, tap_stride, idesc, acc);
                            } else
                            for (int kh = 0; kh < q.taps_h; ++kh) {
                                for (int kw = 0; kw < q.taps_w; ++kw) {
                                    if (!q.b_stationary) {
                                        mbar_wait(&bfull_bar[bstage], bph);
                                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                        db = make_smem_desc(sB0 + (uint32_t)bstage * b_bytes);
                                    }
                                    umma_f16_x4(tacc, da, db, idesc, acc);
                                    acc = 1u;
                                    if (!q.b_stationary) {
                                        umma_commit(&bempty_bar[bstage]);
                                        if (++bstage == q.stages_b) { bstage = 0; bph ^= 1; }
                                    }
                                    da += 8u;                    // next halo pixel (128 B >> 4)
                                    db += tap_stride;
                                }
                                da += row_skip;
                            }
                            umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        // the ring position is warp-uniform state: every lane advances it the same way
                        if (!q.b_stationary) {
                            const int adv = bstage_in + q.taps_h * q.taps_w;
                            bph = bph_in ^ (uint32_t)((adv / q.stages_b) & 1);
                            bstage = adv % q.stages_b;
                        }
                        acc = 1u;
                        if (++stage == stages) { stage = 0; ph ^= 1; }
                    }
                } else {
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[stage], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = smem_u32(smem) + (uint32_t)stage * stage_bytes;
                        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + a_bytes);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_f16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc | (uint32_t)k);
                            umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        acc = 1u;
                        if (++stage == stages) { stage = 0; ph ^= 1; }
                    }
                }
                if (elect_one()) umma_commit(&tfull_bar[buf]);
                __syncwarp();
            }
        }
    } else if (LNMODE == 1 && warp >= 2 + NUM_EPI_WARPS) {
        // ================================ LayerNorm row statistics (ln fold): 2 warps ================================
        // rstd over the c0 input channels of each of the tile's 128 pixels, read from the SAME staged A tiles the MMA consumes
        // (no extra global traffic, no separate normalisation pass).  A row of a K block is 128 bytes = 8 swizzled 16-byte
        // chunks; sums do not care about the chunk order, so the lane owning row r reads chunk (j + r) & 7 at step j: the
        // eight rows a quarter-warp touches per LDS.128 phase hit eight different bank groups.  The mean itself is not needed
        // downstream: fd_ln_fold gives every weight row a zero sum, so W'(x - mean 1) = W'x.
        if (ln_stats) {
            const float inv_c = 1.f / (float)p.c0;
            int stage = 0, it = 0;
            uint32_t ph = 0;
            // sums as packed fp32 pairs (one FADD2 + one FFMA2 per two channels): an ablation showed the statistics ARITHMETIC, not
            // their loads or the barrier protocol, bounding this instantiation (64 -> 256: 759 us; 577 us with the accumulation
            // skipped, against 558 us for the plain GEMM), i.e. the two statistics warps are the critical path of the block
            auto acc8 = [&](const uint4& v, u64& sum, u64& sq) {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 f;
                    if (q.fmt) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                    else f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                    const u64 f2 = f2_pack(f.x, f.y);
                    sum = f2_add(sum, f2);
                    sq = f2_fma(f2, f2, sq);
                }
            };
            const int row0 = (warp - 2 - NUM_EPI_WARPS) * 64 + lane;            // this lane's rows: row0 and row0 + 32 (same row & 7)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                u64 sm[2] = {0ull, 0ull}, sq[2] = {0ull, 0ull};
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], ph);
                    const uint8_t* sa = smem + (size_t)stage * stage_bytes + row0 * 128;
                    // The stage is handed back only AFTER its rows have been accumulated.  Releasing it right after the loads
                    // (the values sit in registers) was 15 % faster but NOT reproducible for GEMMs with two K blocks and two N
                    // tiles (tools/probes/ln_fold_repro.py: rows normalised with wrong statistics in most launches at 16x256^2,
                    // 128 -> 512; never with the late release, never for one K block or one N tile).  The cause was not found
                    // in the barrier protocol; until it is, correctness wins.
                    uint4 v4[2][8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int off = ((j + lane) & 7) << 4;
                        v4[0][j] = *reinterpret_cast<const uint4*>(sa + off);
                        v4[1][j] = *reinterpret_cast<const uint4*>(sa + 32 * 128 + off);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc8(v4[0][j], sm[0], sq[0]);
                        acc8(v4[1][j], sm[1], sq[1]);
                    }
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[stage])) : "memory");
                    if (++stage == stages) { stage = 0; ph ^= 1; }
                }
                mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);     // the epilogue has read the statistics of tile it - 2
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float s0, s1, q0, q1;
                    f2_unpack(sm[r], s0, s1);
                    f2_unpack(sq[r], q0, q1);
                    const float mean = (s0 + s1) * inv_c;
                    s_ln[buf * BM + row0 + r * 32] = make_float2(mean, rsqrtf(fmaxf((q0 + q1) * inv_c - mean * mean, 0.f) + p.ln_eps));
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sfull_bar[buf])) : "memory");
            }
        }
    } else {
        // ================================ epilogue: 16 warps ================================
        const int q4 = warp & 3;                       // TMEM lane quarter this warp may access
        const int cg = (warp - 2) >> 2;                // column group 0..3
        const int cols_per_warp = BN / 4;
        const int m = q4 * 32 + lane;                  // accumulator row = pixel of the tile
        const int gh = q.Hout / (p.upsample ? 2 : 1), gw = q.Wout / (p.upsample ? 2 : 1);
        T* out = (T*)p.out;
        const T* addend = (const T*)p.addend;
        // GroupNorm partial sums: every lane accumulates its own pixel's contribution per 8-column half (8 consecutive
        // columns always share a group, cpg >= 8) across tiles; lanes are only reduced when the sample (or, with several
        // N tiles, the column set) changes.
        float gacc_s[4][2], gacc_q[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { gacc_s[i][0] = gacc_s[i][1] = gacc_q[i][0] = gacc_q[i][1] = 0.f; }
        int cur_b = -1, cur_n0 = 0;
        int cur_lnb = -1;
        // Every epilogue warp owns a row of s_gn (no shared-memory atomics: the order of additions is the program order); the rows
        // are double-buffered by flush parity so that ONE barrier per flush is enough — while the first epilogue warp sums and
        // clears the rows of flush k, the other warps may already be adding into the rows of flush k + 1.
        int gn_par = 0;
        auto gn_reduce_to_smem = [&]() {               // warp-reduce the lane accumulators into this warp's row of s_gn
            float* my_gn = s_gn + (gn_par * NUM_EPI_WARPS + (warp - 2)) * 16;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                if (ci * 16 < cols_per_warp) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const float sv = fd_warp_sum(gacc_s[ci][hf]), qv = fd_warp_sum(gacc_q[ci][hf]);
                        if (lane == 0) {
                            const int g = (cur_n0 + (ci * 4 + cg) * 16 + hf * 8) >> q.cpg_shift;
                            my_gn[g] += sv;
                            my_gn[8 + g] += qv;
                        }
                        gacc_s[ci][hf] = gacc_q[ci][hf] = 0.f;
                    }
                }
            }
        };
        // The block's sums of sample cur_b leave the block (all 16 epilogue warps call this; only the first one works).
        //   gn_ws given: stored into the slot of this block's residue class (tile index within the sample mod grid size) — the set
        //   of tiles behind a slot and their order depend on the sample's geometry only; fd_conv2d_tc_run then launches a
        //   32-thread-per-sample kernel that adds the slots in index order into gn_sums.  Reproducible bit for bit.  (A "last
        //   block finishes" variant inside this kernel cost 30 - 200 % on the GroupNorm convolutions: the fence in front of the
        //   arrival counter waits for the warp's own output stores.)
        //   no gn_ws: floating-point atomics straight into gn_sums (order = block timing).
        const int tps = total_tiles / p.B;             // tiles per sample
        const int nslots = (int)gridDim.x;
        auto gn_flush_sample = [&]() {
            asm volatile("bar.sync 1, 512;" ::: "memory");          // every warp's row of this parity is complete
            const int par = gn_par;
            gn_par ^= 1;
            if (warp != 2 || cur_b < 0) return;
            float tot = 0.f;
            if (lane < 16) {
                float* rows = s_gn + par * NUM_EPI_WARPS * 16 + lane;
#pragma unroll
                for (int w = 0; w < NUM_EPI_WARPS; ++w) { tot += rows[w * 16]; rows[w * 16] = 0.f; }
            }
            const int which = lane >> 3, g = lane & 7;
            if (!p.gn_ws) {
                if (lane < 16 && g < p.gn_groups) atomicAdd(&p.gn_sums[((long)cur_b * p.gn_groups + g) * 2 + which], tot);
                return;
            }
            const int slot = ((int)blockIdx.x + nslots - (int)(((long)cur_b * tps) % nslots)) % nslots;
            if (lane < 16) p.gn_ws[((long)cur_b * nslots + slot) * 16 + lane] = tot;
        };
        int it = 0;
        auto rstd_of = [&](int tl) -> float {          // ln_rstd of this thread's pixel in tile tl (fold: 1x1, no upsampling)
            const TileCoord tn = decode_tile(q, tl);
            const int ri = tn.th * q.tile_h + (m >> q.tile_w_shift), rj = tn.tw * q.tile_w + (m & (q.tile_w - 1));
            return (ri < q.Hout && rj < q.Wout) ? __ldg(p.ln_rstd + ((long)tn.b * q.Hout + ri) * q.Wout + rj) : 1.f;
        };
        float rstd_next = 1.f;
        if (LNMODE == 2 && (int)blockIdx.x < total_tiles) rstd_next = rstd_of((int)blockIdx.x);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc = decode_tile(q, tile);
            const int phase = tc.phase, b = tc.b;
            const int n0 = tc.nt * BN;
            const int oi = tc.th * q.tile_h + (m >> q.tile_w_shift), oj = tc.tw * q.tile_w + (m & (q.tile_w - 1));
            const bool row_ok = oi < gh && oj < gw;
            const int oh = p.upsample ? 2 * oi + (phase >> 1) : oi;
            const int ow = p.upsample ? 2 * oj + (phase & 1) : oj;
            const long orow = (((long)b * q.Hout + oh) * q.Wout + ow) * p.Cout;
            if (p.gn_sums && (b != cur_b || n0 != cur_n0)) {       // uniform over the 16 warps
                if (cur_b >= 0) gn_reduce_to_smem();
                if (b != cur_b) { gn_flush_sample(); cur_b = b; }
                cur_n0 = n0;
            }
            if (ln_fold && b != cur_lnb) {              // uniform over the 16 warps: the sample's ln_v row moves to shared memory
                // (read per 16-column chunk straight from global memory it cost 20 % of the kernel's stall samples: every chunk
                // waited out an L1 round trip in front of its FMAs)
                asm volatile("bar.sync 1, 512;" ::: "memory");
                for (int i = threadIdx.x - 64; i < p.Cout; i += NUM_EPI_WARPS * 32) s_lnv[i] = __ldg(p.ln_v + (long)b * p.Cout + i);
                asm volatile("bar.sync 1, 512;" ::: "memory");
                cur_lnb = b;
            }
            const int buf = it & 1;
            // residual addend of this thread's first chunk: in flight while the accumulator is still being produced; the next
            // chunk's vector is requested while the current one is processed (the load latency was exposed once per chunk)
            uint32_t araw[8];
            const bool use_add = addend != nullptr && row_ok;
            if (use_add) load16_raw(addend + orow + n0 + cg * 16, araw);
            // external statistics: this tile's value was requested one tile ago (the accumulator is normally ready when the
            // epilogue gets to it, so a load issued here had its whole latency in front of the first FMA: 15 % of the kernel's
            // stall samples); now the next tile's
            const float ln_rstd_g = rstd_next;
            if (LNMODE == 2 && tile + (int)gridDim.x < total_tiles) rstd_next = rstd_of(tile + (int)gridDim.x);
            mbar_wait_long(&tfull_bar[buf], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float ln_rstd = 1.f;
            if (ln_stats) {                            // statistics of this thread's pixel (written by the statistics warp)
                mbar_wait(&sfull_bar[buf], (it >> 1) & 1);
                ln_rstd = s_ln[buf * BM + m].y;
            } else if (ln_fold) {
                ln_rstd = ln_rstd_g;
            }
            const uint32_t tacc = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(q4 * 32) << 16);
            // (Software-pipelining the TMEM reads — tcgen05.ld of chunk ci + 1 in flight under chunk ci's math — needs a second
            // 16-register set; at the 96-register cap of an 18-warp block it spilled 190 bytes per thread.  Not done.)
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                const int cc = ci * 16;
                if (cc >= cols_per_warp) break;
                const int c = (ci * 4 + cg) * 16;      // 16-column chunks dealt round-robin to the 4 column groups: a SiLU
                                                       // range (upper half of in_proj) is spread evenly over the warps
                uint32_t r[16];
                tmem_ld16(tacc + (uint32_t)c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc + 16 >= cols_per_warp) {        // last TMEM read of this tile: hand the accumulator back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[buf])) : "memory");
                }
                float v[16];
                const int n = n0 + c;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (ln_fold) {                         // W LN(x) = rstd W' x + v   (fd_ln_fold made W' with zero row sums, and v)
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        // (an explicit ld.shared here instead of the generic load the compiler emits — the shared-memory carve-up
                        // goes through an integer round-up — measured SLOWER, 715 -> 750 us at 64 -> 256: the volatile asm pins the
                        // load in front of its FMAs)
                        const float4 vv = *reinterpret_cast<const float4*>(s_lnv + n + j);
                        v[j] = fmaf(ln_rstd, v[j], vv.x);
                        v[j + 1] = fmaf(ln_rstd, v[j + 1], vv.y);
                        v[j + 2] = fmaf(ln_rstd, v[j + 2], vv.z);
                        v[j + 3] = fmaf(ln_rstd, v[j + 3], vv.w);
                    }
                }
                if (p.bias) {                          // warp-uniform branches, 16-byte parameter loads
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                        v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                    }
                }
                if (n >= p.silu_from) {                // silu_from is a multiple of 16 (checked on the host)
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fast_silu(v[j]);
                }
                if (p.gn_sums && row_ok) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        float sv = 0.f, qv = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { const float x = v[hf * 8 + j]; sv += x; qv = fmaf(x, x, qv); }
                        gacc_s[ci][hf] += sv;
                        gacc_q[ci][hf] += qv;
                    }
                }
                if (row_ok) {
                    if (p.gate) {
                        const float* gp = p.gate + (long)b * p.gate_stride + n;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 gv = __ldg(reinterpret_cast<const float4*>(gp + j));
                            v[j] *= gv.x; v[j + 1] *= gv.y; v[j + 2] *= gv.z; v[j + 3] *= gv.w;
                        }
                    }
                    if (use_add) {
                        add16_raw<T>(araw, v);
                        if (cc + 16 < cols_per_warp) load16_raw(addend + orow + n0 + ((ci + 1) * 4 + cg) * 16, araw);
                    }
                    if (p.relu_out) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    store16<T>(out + orow + n, v);
                }
            }
        }
        if (p.gn_sums) {
            if (cur_b >= 0) gn_reduce_to_smem();
            gn_flush_sample();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// gn_sums[b, g, which] = sum over the slots, in slot order (one warp per sample; slots of blocks that own no tile of the sample
// hold the zeros the caller put there)
__global__ void __launch_bounds__(32) gn_finalize_kernel(const float* __restrict__ ws, float* __restrict__ sums, int nslots, int groups) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (lane >= 16) return;
    float t = 0.f;
    for (int r = 0; r < nslots; ++r) t += ws[((long)b * nslots + r) * 16 + lane];
    const int which = lane >> 3, g = lane & 7;
    if (g < groups) sums[((long)b * groups + g) * 2 + which] = t;
}

// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

int encode(CUtensorMap* map, int dtype, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, const cuuint32_t* estr) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return FD_ERR_DRIVER;
    CUresult r = fn(map, dtype == FD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                    const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : FD_ERR_DRIVER;
}

int act_map(CUtensorMap* map, const void* base, int dtype, int B, int H, int W, int C, int ld, int estride, int box_w, int box_h) {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    const cuuint32_t box[4] = {BK, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    return encode(map, dtype, 4, base, dims, strides, box, estr);
}

}  // namespace

struct fd_gemm_plan {
    alignas(64) CUtensorMap map_a0;
    alignas(64) CUtensorMap map_a1;
    alignas(64) CUtensorMap map_w;
    TcParams q;
    size_t smem;
    dim3 grid;
};

static int conv_dims(const fd_conv_params* p, int* Hout, int* Wout) {
    const int Hv = p->upsample ? 2 * p->Hin : p->Hin, Wv = p->upsample ? 2 * p->Win : p->Win;
    *Hout = (Hv + 2 * p->pad - p->KH) / p->stride + 1;
    *Wout = (Wv + 2 * p->pad - p->KW) / p->stride + 1;
    return (*Hout > 0 && *Wout > 0) ? 0 : FD_ERR_BAD_ARGUMENT;
}

extern "C" int fd_conv_check_params(const fd_conv_params* p);

extern "C" long fd_conv_gn_ws_floats(int B) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;   // no device: the B200 count
    return B > 0 ? (long)B * sms * 16 : 0;
}

extern "C" int fd_conv2d_tc_supported(const fd_conv_params* p) {
    if (fd_conv_check_params(p)) return 0;
    if (p->dtype != FD_BF16 && p->dtype != FD_F16) return 0;
    if (p->ab_dtype_p1 && p->ab_dtype_p1 - 1 != FD_BF16 && p->ab_dtype_p1 - 1 != FD_F16) return 0;
    if (p->c0 % BK || p->c1 % BK || p->ld0 % 8) return 0;
    if (p->silu_from < p->Cout && p->silu_from % 16) return 0;
    if (((uintptr_t)p->bias | (uintptr_t)p->gate) & 15 || (p->gate && p->gate_stride % 4)) return 0;
    if (p->Cout % 64) return 0;
    int Hout, Wout;
    if (conv_dims(p, &Hout, &Wout)) return 0;
    const int gh = p->upsample ? Hout / 2 : Hout, gw = p->upsample ? Wout / 2 : Wout;   // per-phase output grid
    const bool box_ok = gh % TILE_H == 0 && gw % TILE_W == 0, halo_ok = gh % 16 == 0 && gw % 8 == 0;
    if (!box_ok && !(halo_ok && p->stride == 1 && (p->upsample || p->KH * p->KW > 1))) return 0;
    if (p->upsample) {
        if (!(p->KH == 3 && p->KW == 3 && p->stride == 1 && p->pad == 1) || !p->weight_up4 || p->per_batch_weight) return 0;
    } else {
        if (p->stride != 1 && p->stride != 2) return 0;
    }
    if (p->gn_sums) {
        const int cpg = p->Cout / p->gn_groups;
        if (cpg < 8 || cpg % 8 || (cpg & (cpg - 1)) || p->gn_groups > 8) return 0;
    }
    if (p->ln_v) {                // LayerNorm fold: plain 1x1 GEMM over one input tensor (box mode), 16-byte aligned vector
        if (p->KH != 1 || p->KW != 1 || p->stride != 1 || p->pad != 0 || p->upsample || p->c1) return 0;
        if (((uintptr_t)p->ln_v & 15) || ((uintptr_t)p->ln_rstd & 3) || p->Cout % 4 || p->Cout > 1024) return 0;
    } else if (p->ln_rstd) {
        return 0;
    }
    const uintptr_t al = (uintptr_t)p->src0 | (uintptr_t)p->src1 | (uintptr_t)p->weight | (uintptr_t)p->out |
                         (uintptr_t)p->addend | (uintptr_t)p->weight_up4;
    if (al & 15) return 0;
    return get_encode() != nullptr;
}

extern "C" int fd_conv2d_tc_plan_create(const fd_conv_params* p, fd_gemm_plan** out_plan) {
    if (!out_plan) return FD_ERR_BAD_ARGUMENT;
    *out_plan = nullptr;
    if (!fd_conv2d_tc_supported(p)) return FD_ERR_UNSUPPORTED;
    fd_gemm_plan* plan = nullptr;
    if (posix_memalign((void**)&plan, 64, sizeof(fd_gemm_plan)) || !plan) return FD_ERR_BAD_ARGUMENT;
    memset(plan, 0, sizeof(*plan));
    TcParams& q = plan->q;
    q.p = *p;
    conv_dims(p, &q.Hout, &q.Wout);
    const int Cin = p->c0 + p->c1;
    q.phases = p->upsample ? 4 : 1;
    q.taps_h = p->upsample ? 2 : p->KH;
    q.taps_w = p->upsample ? 2 : p->KW;
    q.kblocks0 = p->c0 / BK;
    q.kblocks1 = p->c1 / BK;
    // N tile: 256 / 192 / 128 / 64 (tcgen05 accepts any multiple of 16 up to 256 at M = 128); 192 serves the qkv projections
    // (Cout = 3C = 192, 384) in one or two tiles instead of three
    q.BN = (p->Cout % 256 == 0) ? 256 : (p->Cout % 192 == 0 && !getenv("FD_CONV_NO_BN192")) ? 192 : (p->Cout % 128 == 0 ? 128 : 64);
    q.n_tiles = p->Cout / q.BN;
    const int ab_dtype = p->ab_dtype_p1 ? p->ab_dtype_p1 - 1 : p->dtype;      // operand storage (both A and B); T = output storage
    q.fmt = ab_dtype == FD_BF16 ? 1 : 0;
    const int gh = q.Hout / (p->upsample ? 2 : 1), gw = q.Wout / (p->upsample ? 2 : 1);
    const int num_kb = q.taps_h * q.taps_w * (q.kblocks0 + q.kblocks1);
    const size_t b_tile = (size_t)q.BN * BK * 2;
    const size_t kSmemBudget = 212 * 1024;      // data region; + 1 KB alignment slack + 512 B barriers <= the 220 KB opt-in
    q.halo = (p->stride == 1 && q.taps_h * q.taps_w > 1 && gh % 16 == 0 && gw % 8 == 0 && !getenv("FD_CONV_NO_HALO")) ? 1 : 0;
    if (q.halo) {
        q.tile_h = 16; q.tile_w = 8;
        q.ht_h = q.tile_h + q.taps_h - 1; q.ht_w = q.tile_w + q.taps_w - 1;
        q.a_stage_bytes = (q.ht_h * q.ht_w * 128 + 1023) & ~1023;
        q.b_stationary = (q.phases == 1 && !p->per_batch_weight && q.BN == p->Cout &&
                          num_kb * b_tile + 3 * (size_t)q.a_stage_bytes <= kSmemBudget + 6 * 1024) ? 1 : 0;
        if (q.b_stationary) {
            q.stages = (int)((kSmemBudget + 6 * 1024 - num_kb * b_tile) / q.a_stage_bytes);
            q.stages_b = 0;
        } else {
            q.stages = 3;
            q.stages_b = (int)((kSmemBudget - 3 * (size_t)q.a_stage_bytes) / b_tile);
            if (q.stages_b > MAX_STAGES) q.stages_b = MAX_STAGES;
            if (q.stages_b < 2) q.halo = 0;
            // spend what is left on deeper A prefetch
            while (q.halo && q.stages < MAX_STAGES &&
                   (size_t)(q.stages + 1) * q.a_stage_bytes + q.stages_b * b_tile <= kSmemBudget) ++q.stages;
        }
        if (q.stages > MAX_STAGES) q.stages = MAX_STAGES;
    }
    if (!q.halo) {
        q.tile_h = TILE_H; q.tile_w = TILE_W; q.ht_h = q.ht_w = q.a_stage_bytes = q.stages_b = q.b_stationary = 0;
        if (gh % TILE_H || gw % TILE_W) { free(plan); return FD_ERR_UNSUPPORTED; }
    }
    q.tiles_h = gh / q.tile_h;
    q.tiles_w = gw / q.tile_w;
    const int box_w = q.halo ? q.ht_w : TILE_W, box_h = q.halo ? q.ht_h : TILE_H;
    int rc = act_map(&plan->map_a0, p->src0, ab_dtype, p->B, p->Hin, p->Win, p->c0, p->ld0 > 0 ? p->ld0 : p->c0,
                     p->upsample ? 1 : p->stride, box_w, box_h);
    if (!rc && p->c1) rc = act_map(&plan->map_a1, p->src1, ab_dtype, p->B, p->Hin, p->Win, p->c1, p->c1, p->upsample ? 1 : p->stride, box_w, box_h);
    if (!rc && !p->c1) plan->map_a1 = plan->map_a0;
    if (!rc) {
        const long Kp = (long)q.taps_h * q.taps_w * Cin;
        const int G = p->upsample ? 4 : (p->per_batch_weight ? p->B : 1);
        const void* wbase = p->upsample ? p->weight_up4 : p->weight;
        const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)p->Cout, (cuuint64_t)G};
        const cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)Kp * p->Cout * 2};
        const cuuint32_t box[3] = {BK, (cuuint32_t)q.BN, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        rc = encode(&plan->map_w, ab_dtype, 3, wbase, dims, strides, box, estr);
    }
    if (rc) { free(plan); return rc; }
    const size_t stage_bytes = BM * BK * 2 + (size_t)q.BN * BK * 2;
    if (!q.halo) {
        q.stages = (int)((190 * 1024) / stage_bytes);
        if (q.stages > MAX_STAGES) q.stages = MAX_STAGES;
        if (q.stages > 2 * num_kb) q.stages = 2 * num_kb < 2 ? 2 : 2 * num_kb;     // enough to prefetch the next tile
        if (p->ln_v && !p->ln_rstd && q.stages < 3 && 3 * stage_bytes <= 190 * 1024) q.stages = 3;   // a stage also waits for the statistics warps
    }
    q.total_tiles = p->B * q.phases * q.tiles_h * q.tiles_w * q.n_tiles;

    {
        auto mk = [](int d) { FastDiv f; f.d = (uint32_t)d; f.mul = d > 1 ? (uint32_t)((((uint64_t)1 << 32) + d - 1) / d) : 0u; return f; };
        q.dv_n = mk(q.n_tiles); q.dv_w = mk(q.tiles_w); q.dv_h = mk(q.tiles_h); q.dv_p = mk(q.phases);
        int dmax = q.n_tiles;
        if (q.tiles_w > dmax) dmax = q.tiles_w;
        if (q.tiles_h > dmax) dmax = q.tiles_h;
        if ((uint64_t)q.total_tiles * (uint64_t)dmax >= ((uint64_t)1 << 32)) { free(plan); return FD_ERR_UNSUPPORTED; }
        q.tile_w_shift = q.tile_w == 8 ? 3 : 4;
        q.cpg_shift = 3;
        if (p->gn_sums) { int cpg = p->Cout / p->gn_groups; q.cpg_shift = 0; while ((1 << q.cpg_shift) < cpg) ++q.cpg_shift; }
    }
    const size_t data_bytes = q.halo ? (size_t)(q.b_stationary ? num_kb : q.stages_b) * b_tile + (size_t)q.stages * q.a_stage_bytes
                                     : (size_t)q.stages * stage_bytes;
    plan->smem = data_bytes + 1024 /*align slack*/ + 5120 /*barriers, per-warp GroupNorm rows (two parities), LayerNorm row statistics*/
                 + (p->ln_v ? 4096 : 0) /*ln_v of the current sample*/;
    if (plan->smem > 225 * 1024) { free(plan); return FD_ERR_UNSUPPORTED; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // with gn_ws the slot classes are residues modulo the grid size: always one block per SM, whatever the batch
    plan->grid = dim3((unsigned)((q.total_tiles < sms && !(p->gn_sums && p->gn_ws)) ? q.total_tiles : sms));
    *out_plan = plan;
    return 0;
}

template <typename T, int LNMODE>
static int conv_tc_launch(const fd_gemm_plan* plan, cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<T, LNMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    conv_tc_kernel<T, LNMODE><<<plan->grid, LNMODE == 1 ? NTHREADS_LN : NTHREADS, plan->smem, stream>>>(plan->map_a0, plan->map_a1, plan->map_w, plan->q);
    return 0;
}

extern "C" int fd_conv2d_tc_run(const fd_gemm_plan* plan, cudaStream_t stream) {
    if (!plan) return FD_ERR_BAD_ARGUMENT;
    const int mode = plan->q.p.ln_v == nullptr ? 0 : plan->q.p.ln_rstd ? 2 : 1;
    int rc;
    if (plan->q.p.dtype == FD_BF16)
        rc = mode == 2 ? conv_tc_launch<__nv_bfloat16, 2>(plan, stream) : mode ? conv_tc_launch<__nv_bfloat16, 1>(plan, stream) : conv_tc_launch<__nv_bfloat16, 0>(plan, stream);
    else
        rc = mode == 2 ? conv_tc_launch<__half, 2>(plan, stream) : mode ? conv_tc_launch<__half, 1>(plan, stream) : conv_tc_launch<__half, 0>(plan, stream);
    if (rc) return rc;
    FD_LAUNCH_CHECK();
    if (plan->q.p.gn_sums && plan->q.p.gn_ws) {
        gn_finalize_kernel<<<plan->q.p.B, 32, 0, stream>>>(plan->q.p.gn_ws, plan->q.p.gn_sums, (int)plan->grid.x, plan->q.p.gn_groups);
        FD_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" void fd_conv2d_tc_plan_destroy(fd_gemm_plan* plan) { free(plan); }
