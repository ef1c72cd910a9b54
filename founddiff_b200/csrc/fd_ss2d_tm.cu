// TIME-MAJOR SS2D core for the 16-bit storage modes (src/emamba2.py:295-367: EfficientScan -> x_proj / dt_proj ->
// SelectiveScan -> EfficientMerge), laid out for a channel-per-lane scan:
//
//   xs_tm   (B, 4, L, D)   16-bit   scan input u: direction k, step l, D channels contiguous
//   xdbl_tm (B, 4, L, XR)  fp32     per step: [dt low-rank input (R, only when dt_proj is fused into the scan) | B (N) | C (N)]
//   dts_tm  (B, 4, L, D)   16-bit   delta = softplus(dt_proj + bias), only for the levels whose dt_rank is too large to fuse
//
// Why time-major.  In the reference's (B, 4D, L) layout a (batch, channel) row is contiguous in l, which suits a warp that
// scans ONE row cooperatively (round 1: 5 shuffle rounds per state and chunk, ~15 issue slots per state update, 0.27 of the
// HBM roof at the full-resolution level).  With the channel index fastest a LANE owns a channel and walks time sequentially
// with its states in registers (FMUL, MUFU.EX2, FMUL, FFMA, FFMA per state update); a warp's 32 lanes read 64 contiguous
// bytes per step, B / C / dt-input of the step are warp-uniform broadcasts from shared memory, the depthwise convolution
// writes its NHWC rows unchanged (a pixel permutation, no transpose), x_proj becomes a plain K-contiguous GEMM and the
// EfficientMerge store is 64 contiguous bytes of one pixel.
//
// Rows are long (L = 65536 at 512^2) and there are few of them (8192 channel rows at level 0 = 256 warps), so each row is cut
// into S SEGMENTS that run concurrently.  The state entering a segment is obtained EXACTLY by a carry pass (pass 1): for each
// segment but the last, h_end(from h = 0) = sum_t (prod_{s>t} a_s) b_t is accumulated walking BACKWARDS from the segment end,
// together with P = prod a_s = 2^(A2 * sum dt).  The walk stops once the slowest state of every channel of the block has decayed
// below 2^-30 (the remaining terms are below fp32 round-off of the sum; P is then 0 to the same accuracy); a channel that decays
// slower than the segment is simply walked to the segment start, i.e. the result never depends on an assumed memory length.
// Pass 2 folds the carries of the preceding segments (h = P_j h + hloc_j, a few FMAs per lane) and scans forward.
#include <stdlib.h>
#include <type_traits>

#include "fd_common.cuh"

namespace {

// =========================================================================================================
// K1: depthwise 3x3 + bias + SiLU over the x half of xz (NHWC, row pitch ld) -> xs_tm.
// Register sliding window as dwconv3x3_nhwc_v2_kernel (fd_attn.cu): a thread owns a horizontal pixel PAIR (even x, x+1) x 4
// channels and walks down 32 rows with a 3-row ring of raw 8-byte loads.  The two pixels of a pair belong to directions k and
// k + 2 at the SAME step l, rows alternate between the row-major (even y) and column-major (odd y) sub-grids.
constexpr int TM_V = 4;
constexpr int TM_RY = 32;

template <typename T> FD_DEVINL uint2 tm_ld_raw(const T* p) { return *reinterpret_cast<const uint2*>(p); }
template <typename T> FD_DEVINL void tm_cvt(uint2 r, float (&v)[4]) {
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
template <typename T> FD_DEVINL void tm_st(T* p, const float (&v)[4]) {
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

// 16-bit pair -> packed fp32 pair (FFMA2 operand)
template <typename T> FD_DEVINL u64 tm_cvt2(uint32_t w) {
    float2 f;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
    else f = __half22float2(*reinterpret_cast<__half2*>(&w));
    return f2_pack(f.x, f.y);
}

// The 9-tap accumulation runs on packed fp32 pairs (two adjacent channels per FFMA2): the kernel is bound by issue slots (240 of
// the 710 instructions of a 3-row iteration were FFMA; 607 with 108 FFMA2), and a packed FMA takes one slot for two results.
template <typename T, bool EDGE>
FD_DEVINL void dwtm_rows(const T* __restrict__ in_b, T* __restrict__ out_b, const u64 (&wr)[9][TM_V / 2], const u64 (&bs)[TM_V / 2],
                         int H, int W, int ld, int D, int x, int y0, int y1, bool has_l, bool has_2, int pf) {
    constexpr int NP = TM_V / 2;
    const int H2 = H >> 1, W2 = W >> 1, x2 = x >> 1;
    const long L = (long)H2 * W2;
    const int ymax = min(H - 1, y1);
    int yi = y0 - 1;
    yi -= ((yi % 3) + 3) % 3;                            // round down to a multiple of 3
    const long rowe = (long)W * ld;
    const int offl = (!EDGE || has_l) ? ld : 0, off1 = ld, off2 = (!EDGE || has_2) ? 2 * ld : 0;   // W is even: x + 1 < W always
    const T* p = in_b + (long)min(max(yi, 0), ymax) * rowe + (long)x * ld;      // row the next fetch reads (clamped into the image)
    int yf = yi;
    u64 acc[3][2][NP];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < NP; ++e) acc[r][px][e] = 0ull;
    uint2 ring[3][4];
    auto fetch = [&](auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        ring[S][0] = tm_ld_raw<T>(p - offl);
        ring[S][1] = tm_ld_raw<T>(p);
        ring[S][2] = tm_ld_raw<T>(p + off1);
        ring[S][3] = tm_ld_raw<T>(p + off2);
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
        u64 v[4][NP];
        {
            const bool rv = yrow >= 0 && yrow < H;       // block-uniform; straight-line masking keeps the loads in flight
            const bool ok[4] = {rv && (!EDGE || has_l), rv, rv, rv && (!EDGE || has_2)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 r = ring[S][j];
                r.x = ok[j] ? r.x : 0u;
                r.y = ok[j] ? r.y : 0u;
                v[j][0] = tm_cvt2<T>(r.x);
                v[j][1] = tm_cvt2<T>(r.y);
            }
        }
        fetch(slot_c);
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < NP; ++e) {
                u64 a1 = f2_fma(v[px][e], wr[0][e], bs[e]);                        // output row yrow+1: first contribution
                a1 = f2_fma(v[px + 1][e], wr[1][e], a1);
                acc[S1][px][e] = f2_fma(v[px + 2][e], wr[2][e], a1);
                u64 a0 = f2_fma(v[px][e], wr[3][e], acc[S][px][e]);               // output row yrow
                a0 = f2_fma(v[px + 1][e], wr[4][e], a0);
                acc[S][px][e] = f2_fma(v[px + 2][e], wr[5][e], a0);
                u64 a2 = f2_fma(v[px][e], wr[6][e], acc[S2][px][e]);              // output row yrow-1 (complete after this)
                a2 = f2_fma(v[px + 1][e], wr[7][e], a2);
                acc[S2][px][e] = f2_fma(v[px + 2][e], wr[8][e], a2);
            }
        const int yo = yrow - 1;
        if (yo >= y0 && yo < y1) {                       // block-uniform
            const int yh = yo >> 1;
            // even rows: directions 0 / 2, row-major l = (y/2) W2 + x/2; odd rows: directions 1 / 3, column-major l = (x/2) H2 + y/2
            const long off = (yo & 1) ? (L + (long)x2 * H2 + yh) * D : ((long)yh * W2 + x2) * D;
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                float o[TM_V];
#pragma unroll
                for (int e = 0; e < NP; ++e) {
                    float a, b;
                    f2_unpack(acc[S2][px][e], a, b);
                    o[2 * e] = fd_silu16(a);
                    o[2 * e + 1] = fd_silu16(b);
                }
                tm_st<T>(out_b + off + (px ? 2 * L * D : 0), o);
            }
        }
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

template <typename T>
__global__ void __launch_bounds__(256, 2) dwconv_tm_kernel(const T* __restrict__ xz, int ld, const float* __restrict__ w,
                                                           const float* __restrict__ bias, T* __restrict__ xs, int H, int W,
                                                           int D, int pf) {
    const int NV = D / TM_V;
    const int W2 = W >> 1;
    const long f2 = (long)blockIdx.x * 256 + threadIdx.x;   // index over (pixel pair, channel vector)
    const bool live = f2 < (long)W2 * NV;
    const long f2c = live ? f2 : 0;
    const int cv = (int)(f2c % NV), x = 2 * (int)(f2c / NV);
    const int y0 = blockIdx.y * TM_RY, y1 = min(H, y0 + TM_RY);
    u64 wr[9][TM_V / 2], bs[TM_V / 2];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long)t * D + cv * TM_V));
        wr[t][0] = f2_pack(wv.x, wv.y); wr[t][1] = f2_pack(wv.z, wv.w);
    }
    {
        const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias + cv * TM_V)) : make_float4(0.f, 0.f, 0.f, 0.f);
        bs[0] = f2_pack(bv.x, bv.y); bs[1] = f2_pack(bv.z, bv.w);
    }
    const bool has_l = x > 0, has_2 = x + 2 < W;
    const T* in_b = xz + (long)blockIdx.z * H * W * ld + cv * TM_V;
    T* out_b = xs + (long)blockIdx.z * H * W * D + cv * TM_V;          // 4 L D == H W D elements per sample
    if (!live) return;
    if (__any_sync(__activemask(), !(has_l && has_2))) dwtm_rows<T, true>(in_b, out_b, wr, bs, H, W, ld, D, x, y0, y1, has_l, has_2, pf);
    else dwtm_rows<T, false>(in_b, out_b, wr, bs, H, W, ld, D, x, y0, y1, true, true, pf);
}

// =========================================================================================================
// K2: x_proj (+ dt_proj for the levels whose dt_rank is not fused into the scan) on time-major input.
//   stage 1:  X_dbl[l, c] = sum_d xs_tm[l, d] Wx[k][c, d]       M = 128 steps per block (16 per warp), N = CCp, K = D
//   stage 2:  dts[l, d]   = softplus(sum_r X_dbl[l, r] Wdt[k][d, r] + bias[d])   (WITH_DT)   M = steps, N = D, K = Rp
// mma.sync m16n8k16, fp32 accumulate; both operands are K-contiguous, so plain (non-transposing) ldmatrix feeds them.  The
// stage-1 accumulator fragments of the dt columns ARE the stage-2 A fragments (C layout of two adjacent n-tiles = A layout
// of one k-tile), so X_dbl's dt part never leaves the registers.  Bound by the single read of xs_tm.
constexpr int XM_STEPS = 128, XM_KC = 64, XM_LD = XM_KC + 8, XM_STAGES = 3;

FD_DEVINL void tm_cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
FD_DEVINL float tm_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
FD_DEVINL float tm_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log1p(exp(x)), branch-free, 2 MUFU ops + 9 FP32 ops.  w = 1 + y rounds; for y < 1 the lost part d = y - (w - 1) is exact and
// log1p(y) = log(w) + log1p(d / w) ~ log(w) + d (1 - y) (the factor only has to be right to O(1): the term is <= half an ulp of
// w); for y >= 1 the factor (1 - min(y, 1)) switches the correction off.  Large x: log2(1 + 2^(x log2 e)) ln 2 = x to fp32
// precision (selective_scan_fn switches to the identity above 20: same value within 3e-7 relative); the clamp keeps y finite.
FD_DEVINL float tm_softplus(float x) {
    x = fminf(x, 80.f);
    const float y = tm_ex2(x * 1.4426950408889634f);
    const float w = 1.f + y;
    const float corr = (y - (w - 1.f)) * (1.f - fminf(y, 1.f));
    return fmaf(tm_lg2(w), 0.6931471805599453f, corr);
}
// tm_softplus on two values at once: 4 MUFU + 2 FMNMX clamps + 2 FMNMX + 7 packed ops (15 issue slots for two, 11 each before)
FD_DEVINL void tm_softplus2(float x0, float x1, float& d0, float& d1) {
    const u64 xa = f2_mul(f2_pack(fminf(x0, 80.f), fminf(x1, 80.f)), f2_pack(1.4426950408889634f, 1.4426950408889634f));
    float a0, a1;
    f2_unpack(xa, a0, a1);
    const float y0 = tm_ex2(a0), y1 = tm_ex2(a1);
    const u64 y = f2_pack(y0, y1);
    const u64 w = f2_add(y, f2_pack(1.f, 1.f));
    const u64 wm1 = f2_add(w, f2_pack(-1.f, -1.f));
    const u64 lost = f2_fma(wm1, f2_pack(-1.f, -1.f), y);                              // y - (w - 1), exact for y < 1
    const u64 fac = f2_fma(f2_pack(fminf(y0, 1.f), fminf(y1, 1.f)), f2_pack(-1.f, -1.f), f2_pack(1.f, 1.f));
    const u64 corr = f2_mul(lost, fac);
    float w0, w1;
    f2_unpack(w, w0, w1);
    const u64 r = f2_fma(f2_pack(tm_lg2(w0), tm_lg2(w1)), f2_pack(0.6931471805599453f, 0.6931471805599453f), corr);
    f2_unpack(r, d0, d1);
}

template <typename T> FD_DEVINL uint32_t tm_pack2(float a, float b) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = fd_floats2half2_sat(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

template <typename T, int NT, bool WITH_DT>      // NT = CCp / 8 (even)
__global__ void __launch_bounds__(256) x_proj_tm_kernel(const T* __restrict__ xs, const T* __restrict__ xw16, float* __restrict__ xdbl,
                                                        int D, int L, int CC, int col0, int out_ld, const T* __restrict__ dw16,
                                                        T* __restrict__ dts, const float* __restrict__ dt_bias, int Rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CCp = NT * 8;
    T* s_a = reinterpret_cast<T*>(smem_raw);                     // [XM_STAGES][XM_STEPS][XM_LD]
    T* s_w = s_a + XM_STAGES * XM_STEPS * XM_LD;                 // [XM_STAGES][CCp][XM_LD]
    const int bk = blockIdx.y, k = bk & 3;
    const int l0 = blockIdx.x * XM_STEPS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const T* xr = xs + (long)bk * L * D;
    const T* wx = xw16 + (long)k * CCp * D;
    const int nchunks = D / XM_KC;

    auto stage = [&](int ch, int buf) {
        const int d0 = ch * XM_KC;
        T* sa = s_a + buf * XM_STEPS * XM_LD;
        T* sw = s_w + buf * CCp * XM_LD;
        for (int i = tid; i < XM_STEPS * (XM_KC / 8); i += 256) {
            const int r = i / (XM_KC / 8), v = i % (XM_KC / 8);
            const bool ok = l0 + r < L;
            tm_cp_async16(sa + r * XM_LD + v * 8, ok ? xr + (long)(l0 + r) * D + d0 + v * 8 : xr, ok);
        }
        for (int i = tid; i < CCp * (XM_KC / 8); i += 256) {
            const int r = i / (XM_KC / 8), v = i % (XM_KC / 8);
            tm_cp_async16(sw + r * XM_LD + v * 8, wx + (long)r * D + d0 + v * 8, true);
        }
    };

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;

#pragma unroll
    for (int s = 0; s < XM_STAGES - 1; ++s) {
        if (s < nchunks) stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        asm volatile("cp.async.wait_group %0;" ::"n"(XM_STAGES - 2) : "memory");
        __syncthreads();                                          // chunk ch landed; everyone is done with chunk ch-1's buffer
        if (ch + XM_STAGES - 1 < nchunks) stage(ch + XM_STAGES - 1, (ch + XM_STAGES - 1) % XM_STAGES);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const T* sa = s_a + (ch % XM_STAGES) * XM_STEPS * XM_LD;
        const T* sw = s_w + (ch % XM_STAGES) * CCp * XM_LD;
#pragma unroll
        for (int ks = 0; ks < XM_KC / 16; ++ks) {
            uint32_t afr[4];
            ldmatrix_x4(afr, sa + (warp * 16 + (lane & 15)) * XM_LD + ks * 16 + 8 * (lane >> 4));
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t bfr[4];         // (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
                ldmatrix_x4(bfr, sw + (np * 16 + (lane & 7) + 8 * (lane >> 4)) * XM_LD + ks * 16 + 8 * ((lane >> 3) & 1));
                mma_16816<T>(acc[2 * np], afr, bfr[0], bfr[1]);
                mma_16816<T>(acc[2 * np + 1], afr, bfr[2], bfr[3]);
            }
        }
    }
    // stage-1 epilogue: columns [col0, CC) of X_dbl, fp32, time-major rows of out_ld floats
    const int r0 = l0 + warp * 16 + g, r1 = r0 + 8;
    float* o0 = xdbl + ((long)bk * L + r0) * out_ld - col0;
    float* o1 = xdbl + ((long)bk * L + r1) * out_ld - col0;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int col = nt * 8 + 2 * t4;
        if (col >= col0 && col < CC) {
            if (r0 < L) *reinterpret_cast<float2*>(o0 + col) = make_float2(acc[nt][0], acc[nt][1]);
            if (r1 < L) *reinterpret_cast<float2*>(o1 + col) = make_float2(acc[nt][2], acc[nt][3]);
        }
    }
    if constexpr (WITH_DT) {
        // stage 2 straight from the accumulator fragments (dt input rounded to the storage type, as the fp32 reference's
        // consumers see it after a 16-bit x_dbl); columns >= R meet zero rows of the padded dw16
        uint32_t afr[2][4];
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            if (kt * 16 < Rp && 2 * kt + 1 < NT) {
                afr[kt][0] = tm_pack2<T>(acc[2 * kt][0], acc[2 * kt][1]);
                afr[kt][1] = tm_pack2<T>(acc[2 * kt][2], acc[2 * kt][3]);
                afr[kt][2] = tm_pack2<T>(acc[2 * kt + 1][0], acc[2 * kt + 1][1]);
                afr[kt][3] = tm_pack2<T>(acc[2 * kt + 1][2], acc[2 * kt + 1][3]);
            } else {
                afr[kt][0] = afr[kt][1] = afr[kt][2] = afr[kt][3] = 0u;
            }
        }
        const T* wd = dw16 + (long)k * D * Rp;
        const float* bk_bias = dt_bias + (long)k * D;
        // B fragments (dt_proj weights, K-contiguous rows of dw16) come straight from global memory (L1 / L2: the same for every
        // block of a direction).  One 64-channel GROUP at a time: the loads of all 8 n-tiles of the group are issued together,
        // then 8 independent (MMA -> softplus -> tile store) chains follow, fully unrolled — as a one-n-tile loop with a
        // one-deep prefetch a warp sat out an L2 round trip and a 150-cycle dependent softplus chain per n-tile (D = 1024:
        // 128 of them per block, 0.36 instructions per clock and SM).
        const int nk = Rp / 16;
        const int ngroups = D / 64;
        // delta leaves through a per-warp shared-memory tile (16 steps x 64 channels, the pipeline's buffers are free by now): a C
        // fragment holds 2 adjacent channels per lane, so storing from it a warp instruction wrote eight 16-byte pieces in eight
        // cache lines; read back by rows, a lane writes 16 bytes and a warp instruction four complete 128-byte row segments.
        __syncthreads();                                  // every warp is out of the stage-1 loop: s_a is reusable
        constexpr int OT = 64 + 8;                        // tile row pitch (elements): the 8 rows of a fragment store hit 32 banks
        T* ot = s_a + warp * 16 * OT;
        for (int gr = 0; gr < ngroups; ++gr) {
            uint32_t bf[8][2][2];
            float2 bb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int nd = gr * 8 + i;
                const T* wrow = wd + (long)(nd * 8 + g) * Rp + 2 * t4;
#pragma unroll
                for (int kt = 0; kt < 2; ++kt) {
                    if (kt < nk) {
                        bf[i][kt][0] = *reinterpret_cast<const uint32_t*>(wrow + kt * 16);
                        bf[i][kt][1] = *reinterpret_cast<const uint32_t*>(wrow + kt * 16 + 8);
                    } else {
                        bf[i][kt][0] = bf[i][kt][1] = 0u;
                    }
                }
                bb[i] = __ldg(reinterpret_cast<const float2*>(bk_bias + nd * 8 + 2 * t4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kt = 0; kt < 2; ++kt)
                    if (kt < nk) mma_16816<T>(o, afr[kt], bf[i][kt][0], bf[i][kt][1]);
                const int dc = i * 8 + 2 * t4;            // column inside the 64-channel group
                float s0, s1, s2, s3;                     // softplus on packed pairs
                tm_softplus2(o[0] + bb[i].x, o[1] + bb[i].y, s0, s1);
                tm_softplus2(o[2] + bb[i].x, o[3] + bb[i].y, s2, s3);
                *reinterpret_cast<uint32_t*>(ot + g * OT + dc) = tm_pack2<T>(s0, s1);
                *reinterpret_cast<uint32_t*>(ot + (g + 8) * OT + dc) = tm_pack2<T>(s2, s3);
            }
            __syncwarp();
            const int d0 = gr * 64 + (lane & 7) * 8;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int rr = ps * 4 + (lane >> 3);
                const int row = l0 + warp * 16 + rr;
                const uint4 v = *reinterpret_cast<const uint4*>(ot + rr * OT + (lane & 7) * 8);
                if (row < L) *reinterpret_cast<uint4*>(dts + ((long)bk * L + row) * D + d0) = v;
            }
            __syncwarp();
        }
    }
}

template <typename T> FD_DEVINL float tm_ld16(const T* p) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) return __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(p)) << 16);
    else return __half2float(*p);
}

// EfficientMerge (src/emamba2.py:238-262) as a running ELEMENT offset into one sample's (H, W, D) tensor: direction k, step l
// -> pixel (2 (l / W2), 2 (l % W2) + kx) for the row-major directions, (2 (l % H2) + 1, 2 (l / H2) + kx) for the column-major
// ones.  One add per step, one compare; no multiply in the loop.  H * W * D < 2^31 (checked by the launcher).
struct MergeWalk {
    int off, mr, mdiv, inc, wrap;
    FD_DEVINL void init(int k, int l, int H, int W, int D) {
        const int col = k & 1, kx = k >> 1;
        mdiv = col ? (H >> 1) : (W >> 1);
        const int mq = l / mdiv;
        mr = l - mq * mdiv;
        const int hh = col ? 2 * mr + 1 : 2 * mq, ww = col ? 2 * mq + kx : 2 * mr + kx;
        off = (hh * W + ww) * D;
        inc = col ? 2 * W * D : 2 * D;
        wrap = col ? (2 - H * W) * D : W * D;            // extra offset when the fast index wraps
    }
    FD_DEVINL void next() {
        off += inc;
        if (++mr == mdiv) { mr = 0; off += wrap; }
    }
};

// Four consecutive steps of one channel, software-pipelined by hand: the four delta chains (dt_proj FMAs, softplus = 2 MUFU)
// and then the 4 N decay factors are independent of each other and of the recurrence, so they are issued back to back;
// only h = a h + b is sequential.  Left to itself ptxas keeps each step's chain in program order and a warp then sits out
// ~150 clk of dependent latency per step.
template <typename T, int NS, int RDT, int XR, int USTRIDE, bool MASK>
FD_DEVINL void load_steps4(const float* __restrict__ xr, const T* __restrict__ ur, const T* __restrict__ dr, const float (&wdt)[RDT > 0 ? RDT : 1],
                           float bias, int nvalid, float (&dt)[4], float (&u)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if constexpr (RDT > 0) {
            float t = bias;
#pragma unroll
            for (int r = 0; r < RDT; r += 4) {
                const float4 v = *reinterpret_cast<const float4*>(xr + q * XR + r);
                t = fmaf(wdt[r], v.x, t); t = fmaf(wdt[r + 1], v.y, t); t = fmaf(wdt[r + 2], v.z, t); t = fmaf(wdt[r + 3], v.w, t);
            }
            dt[q] = t;
        } else {
            dt[q] = tm_ld16<T>(dr + q * USTRIDE);
        }
        u[q] = tm_ld16<T>(ur + q * USTRIDE);
    }
    if constexpr (RDT > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dt[q] = tm_softplus(dt[q]);
    }
    if constexpr (MASK) {                                // steps past the end of the row / segment: identity (a = 1, b = 0)
#pragma unroll
        for (int q = 0; q < 4; ++q) dt[q] = q < nvalid ? dt[q] : 0.f;
    }
}

// =========================================================================================================
// K3: segmented channel-per-lane selective scan (+ EfficientMerge store).  See the file header.
// Block = WARPS x 32 lanes = CHB consecutive channels of one (sample, direction); grid (segments, D / CHB, B * 4).
// Per 32-step chunk the block stages u (and delta when it is not fused) [32][CHB] 16-bit and the X_dbl rows [32][XR] fp32
// with cp.async, double buffered, one barrier per chunk.  RDT > 0: delta = softplus(dt_w[d, :RDT] . xdbl[l, :RDT] + bias)
// is formed in registers (RDT FMAs + 2 MUFU per step).
constexpr int SC_T = 32;           // steps per staged chunk
constexpr int SC_WARPS = 4;
constexpr int SC_CHB = SC_WARPS * 32;
constexpr float SC_DECAYED = -30.f;            // log2 of the decay below which a carry is dropped


template <typename T, int NS, int RDT, bool CARRY>
__global__ void __launch_bounds__(SC_CHB, (NS >= 16 ? 4 : 6)) scan_tm_kernel(
    const T* __restrict__ u_tm, const T* __restrict__ dts_tm, const float* __restrict__ xdbl, const float* __restrict__ A,
    const float* __restrict__ dt_w, const float* __restrict__ dt_bias, const float* __restrict__ Dskip, float* __restrict__ carry,
    T* __restrict__ y, int D, int L, int H, int W, int S, int seg_len) {
    constexpr int XR = RDT + 2 * NS;                    // floats per step in xdbl
    constexpr bool HAS_DT = RDT == 0;                   // delta comes from dts_tm
    constexpr bool PRE_A = NS <= 8;                     // decay factors of a 4-step block computed ahead of the recurrence
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);                              // [2][SC_T][XR]
    T* s_u = reinterpret_cast<T*>(s_x + 2 * SC_T * XR);                           // [2][SC_T][SC_CHB]
    T* s_d = s_u + 2 * SC_T * SC_CHB;                                             // [2][SC_T][SC_CHB]   (HAS_DT)
    const int tid = threadIdx.x;
    const int seg = blockIdx.x, bk = blockIdx.z, k = bk & 3, b = bk >> 2;
    const int ch0 = blockIdx.y * SC_CHB;
    const int dloc = ch0 + tid;                          // channel within the direction group
    const int d = k * D + dloc;                          // row of A / dt_w / bias / D
    const int t_begin = seg * seg_len, t_end = min(L, t_begin + seg_len);
    const int nch = (t_end - t_begin + SC_T - 1) / SC_T;
    const T* ug = u_tm + (long)bk * L * D + ch0;
    const T* dg = HAS_DT ? dts_tm + (long)bk * L * D + ch0 : nullptr;
    const float* xg = xdbl + (long)bk * L * XR;

    float A2[NS], h[NS];
    float a2max = -INFINITY;
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        A2[n] = A[(long)d * NS + n] * 1.4426950408889634f;
        a2max = fmaxf(a2max, A2[n]);
        h[n] = 0.f;
    }
    float wdt[RDT > 0 ? RDT : 1];
    float bias = 0.f;
    wdt[0] = 0.f;
    if constexpr (RDT > 0) {
#pragma unroll
        for (int r = 0; r < RDT; ++r) wdt[r] = dt_w[(long)d * RDT + r];
        bias = dt_bias[d];
    }
    const float Dd = Dskip[d];

    auto stage = [&](int c, int buf) {                   // chunk c of this segment: steps [t_begin + c*SC_T, +SC_T), zero-filled past t_end
        const int t0 = t_begin + c * SC_T;
        float* sx = s_x + buf * SC_T * XR;
        for (int i = tid; i < SC_T * XR / 4; i += SC_CHB) {           // rows are contiguous in global memory: one flat run
            const bool ok = t0 + (i * 4) / XR < t_end;
            tm_cp_async16(sx + i * 4, ok ? xg + (long)t0 * XR + i * 4 : xg, ok);
        }
        T* su = s_u + buf * SC_T * SC_CHB;
        T* sd = s_d + buf * SC_T * SC_CHB;
        for (int i = tid; i < SC_T * (SC_CHB / 8); i += SC_CHB) {
            const int r = i / (SC_CHB / 8), v = i % (SC_CHB / 8);
            const bool ok = t0 + r < t_end;
            tm_cp_async16(su + r * SC_CHB + v * 8, ok ? ug + (long)(t0 + r) * D + v * 8 : ug, ok);
            if constexpr (HAS_DT) tm_cp_async16(sd + r * SC_CHB + v * 8, ok ? dg + (long)(t0 + r) * D + v * 8 : dg, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if constexpr (CARRY) {
        // ---- pass 1: backward walk from the segment end; acc_n = sum_t 2^(A2_n * cum_t) dt_t u_t B_tn, cum_t = sum_{s > t} dt_s
        float cum = 0.f;
        bool done = false, early = false;
        stage(nch - 1, (nch - 1) & 1);
        for (int c = nch - 1; c >= 0; --c) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (__syncthreads_and(done)) { early = true; break; }       // chunk c visible; every channel of the block has decayed
            if (c > 0) stage(c - 1, (c - 1) & 1);
            const int buf = c & 1;
            const int ns = min(SC_T, t_end - (t_begin + c * SC_T));
            const float* sx = s_x + buf * SC_T * XR;
            const T* su = s_u + buf * SC_T * SC_CHB + tid;
            const T* sd = s_d + buf * SC_T * SC_CHB + tid;
#pragma unroll 2
            for (int s0 = SC_T - 4; s0 >= 0; s0 -= 4) {
                if (s0 >= ns) continue;                  // block-uniform (only the segment's last chunk can be short)
                float dt[4], u[4], cq[4];
                load_steps4<T, NS, RDT, XR, SC_CHB, true>(sx + s0 * XR, su + s0 * SC_CHB, sd + s0 * SC_CHB, wdt, bias, ns - s0, dt, u);
                cq[3] = cum; cq[2] = cq[3] + dt[3]; cq[1] = cq[2] + dt[2]; cq[0] = cq[1] + dt[1];
                cum = cq[0] + dt[0];
#pragma unroll
                for (int q = 3; q >= 0; --q) {
                    const float du = dt[q] * u[q];
                    const float* pb = sx + (s0 + q) * XR + RDT;
#pragma unroll
                    for (int n = 0; n < NS; n += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(pb + n);
                        h[n] = fmaf(tm_ex2(A2[n] * cq[q]), du * b4.x, h[n]);
                        h[n + 1] = fmaf(tm_ex2(A2[n + 1] * cq[q]), du * b4.y, h[n + 1]);
                        h[n + 2] = fmaf(tm_ex2(A2[n + 2] * cq[q]), du * b4.z, h[n + 2]);
                        h[n + 3] = fmaf(tm_ex2(A2[n + 3] * cq[q]), du * b4.w, h[n + 3]);
                    }
                }
            }
            done = a2max * cum < SC_DECAYED;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        float* cb = carry + (((long)bk * S + seg) * 2) * NS * D + dloc;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            cb[(long)n * D] = h[n];
            cb[(long)(NS + n) * D] = early ? 0.f : tm_ex2(A2[n] * cum);
        }
    } else {
        // ---- pass 2: fold the carries of the preceding segments, then scan forward and store through EfficientMerge
        for (int j = 0; j < seg; ++j) {
            const float* cb = carry + (((long)bk * S + j) * 2) * NS * D + dloc;
#pragma unroll
            for (int n = 0; n < NS; ++n) h[n] = fmaf(cb[(long)(NS + n) * D], h[n], cb[(long)n * D]);
        }
        u64 hp[NS / 2], A2p[NS / 2];
#pragma unroll
        for (int j = 0; j < NS / 2; ++j) { hp[j] = f2_pack(h[2 * j], h[2 * j + 1]); A2p[j] = f2_pack(A2[2 * j], A2[2 * j + 1]); }
        MergeWalk mw;
        mw.init(k, t_begin, H, W, D);
        T* ybase = y + (long)b * H * W * D + dloc;
        stage(0, 0);
        for (int c = 0; c < nch; ++c) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                             // chunk c visible; every warp is past chunk c-1
            if (c + 1 < nch) stage(c + 1, (c + 1) & 1);
            const int buf = c & 1;
            const int ns = min(SC_T, t_end - (t_begin + c * SC_T));
            const float* sx = s_x + buf * SC_T * XR;
            const T* su = s_u + buf * SC_T * SC_CHB + tid;
            const T* sd = s_d + buf * SC_T * SC_CHB + tid;
            auto run = [&](auto mask_c) {
                constexpr bool MASK = decltype(mask_c)::value;
#pragma unroll(NS >= 32 ? 1 : 2)
                for (int s0 = 0; s0 < SC_T; s0 += 4) {
                    if (MASK && s0 >= ns) break;
                    float dt[4], u[4];
                    load_steps4<T, NS, RDT, XR, SC_CHB, MASK>(sx + s0 * XR, su + s0 * SC_CHB, sd + s0 * SC_CHB, wdt, bias, ns - s0, dt, u);
                    // states as pairs (n, n + 1): decay argument, input term, recurrence and output dot product are one packed
                    // instruction per pair (3 issue slots per state update instead of 5; the update is then bound by MUFU.EX2
                    // alone: tools/probes/pipe_rates.cu)
                    u64 a[PRE_A ? 4 : 1][PRE_A ? NS / 2 : 1];
                    if constexpr (PRE_A) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const u64 dt2 = f2_pack(dt[q], dt[q]);
#pragma unroll
                            for (int j = 0; j < NS / 2; ++j) {
                                float e0, e1;
                                f2_unpack(f2_mul(dt2, A2p[j]), e0, e1);
                                a[q][j] = f2_pack(tm_ex2(e0), tm_ex2(e1));
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float du = dt[q] * u[q];
                        const u64 du2 = f2_pack(du, du), dt2 = f2_pack(dt[q], dt[q]);
                        const float* pb = sx + (s0 + q) * XR + RDT;
                        const float* pc = pb + NS;
                        u64 ya = 0ull, yb = 0ull;
#pragma unroll
                        for (int n = 0; n < NS; n += 4) {
                            const ulonglong2 b2 = *reinterpret_cast<const ulonglong2*>(pb + n);
                            const ulonglong2 c2 = *reinterpret_cast<const ulonglong2*>(pc + n);
                            u64 a0, a1;
                            if constexpr (PRE_A) {
                                a0 = a[PRE_A ? q : 0][PRE_A ? n / 2 : 0];
                                a1 = a[PRE_A ? q : 0][PRE_A ? n / 2 + 1 : 0];
                            } else {
                                float e0, e1, e2, e3;
                                f2_unpack(f2_mul(dt2, A2p[n / 2]), e0, e1);
                                f2_unpack(f2_mul(dt2, A2p[n / 2 + 1]), e2, e3);
                                a0 = f2_pack(tm_ex2(e0), tm_ex2(e1));
                                a1 = f2_pack(tm_ex2(e2), tm_ex2(e3));
                            }
                            hp[n / 2] = f2_fma(a0, hp[n / 2], f2_mul(du2, b2.x));
                            hp[n / 2 + 1] = f2_fma(a1, hp[n / 2 + 1], f2_mul(du2, b2.y));
                            ya = n == 0 ? f2_mul(hp[0], c2.x) : f2_fma(hp[n / 2], c2.x, ya);
                            yb = n == 0 ? f2_mul(hp[1], c2.y) : f2_fma(hp[n / 2 + 1], c2.y, yb);
                        }
                        if (!MASK || s0 + q < ns) {
                            float y0, y1;
                            f2_unpack(f2_add(ya, yb), y0, y1);
                            fd_st(ybase + mw.off, fmaf(Dd, u[q], y0) + y1);
                            mw.next();
                        }
                    }
                }
            };
            if (ns == SC_T) run(std::false_type{}); else run(std::true_type{});
        }
    }
}

// =========================================================================================================
// K3b: time-sliced cooperative scan for the levels with FEW, LONG rows (full and half resolution: 8 K - 16 K channel rows
// of 16 K - 64 K steps).  The segmented kernel above pays for its parallelism with a second evaluation of exp(dt A) in the
// carry pass — with the reference's dt initialisation (1e-3 .. 1e-1, A = -1 .. -N) the slowest state of some channel of every
// warp remembers longer than a segment, so the carry pass walks whole segments and the MUFU work doubles.  Here every
// exp(dt A) is evaluated ONCE:
//   block = TW warps x 32 lanes; a lane is a channel, warp w owns time slice w of every chunk (chunk = TW x ST steps).
//   phase A   each warp scans its ST steps from h = 0 and keeps, per step, y_local (registers) and the fix-up row
//             g[n] = C[n] * prod(a[n]) (shared memory, one 16-byte vector per step and lane for N = 4);
//             it publishes (P = prod a, h_local) of its slice.
//   barrier   (one per chunk)
//   fold      warp w composes the slices j < w onto the chunk's entry state: h_in = P_j h_in + h_j  (a few FMAs)
//   fix-up    y[t] = y_local[t] + g[t] . h_in  (N FMAs per step, no MUFU), EfficientMerge store
//   the last warp hands the chunk's exit state to the next chunk through shared memory (published by the next barrier).
// u and the X_dbl rows of a warp's slice are staged by the warp itself (cp.async, consumed in phase A only, so the next
// slice's copy is issued right after phase A and lands under the fold / fix-up).  The result equals the sequential
// recurrence up to re-association.
template <typename T, int NS, int RDT, int ST, int TW>
__global__ void __launch_bounds__(TW * 32, (TW == 8 ? 2 : 4)) scan_tw_kernel(
    const T* __restrict__ u_tm, const float* __restrict__ xdbl, const float* __restrict__ A, const float* __restrict__ dt_w,
    const float* __restrict__ dt_bias, const float* __restrict__ Dskip, T* __restrict__ y, int D, int L, int H, int W) {
    constexpr int XR = RDT + 2 * NS;
    constexpr int CH = TW * ST;                          // steps per chunk
    static_assert(ST % 4 == 0 && NS % 4 == 0 && RDT % 4 == 0 && RDT > 0, "4-step blocks, float4 rows");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_g = reinterpret_cast<float*>(smem_raw);     // [TW][ST][32][NS]
    float* s_x = s_g + TW * ST * 32 * NS;                // [TW][ST][XR]
    float* s_ph = s_x + TW * ST * XR;                    // [2][TW][2 NS][32]
    float* s_cy = s_ph + 2 * TW * 2 * NS * 32;           // [2][NS][32]
    T* s_u = reinterpret_cast<T*>(s_cy + 2 * NS * 32);   // [TW][ST][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bk = blockIdx.y, k = bk & 3, b = bk >> 2;
    const int ch0 = blockIdx.x * 32;
    const int dloc = ch0 + lane;
    const int d = k * D + dloc;
    const T* ug = u_tm + (long)bk * L * D + ch0;
    const float* xg = xdbl + (long)bk * L * XR;
    float* gw = s_g + warp * ST * 32 * NS + lane * NS;
    float* xw = s_x + warp * ST * XR;
    T* uw = s_u + warp * ST * 32;

    float A2[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) A2[n] = A[(long)d * NS + n] * 1.4426950408889634f;
    float wdt[RDT];
#pragma unroll
    for (int r = 0; r < RDT; ++r) wdt[r] = dt_w[(long)d * RDT + r];
    const float bias = dt_bias[d], Dd = Dskip[d];
    if (warp == 0) {
#pragma unroll
        for (int n = 0; n < NS; ++n) s_cy[n * 32 + lane] = 0.f;          // entry state of chunk 0 (published by the first barrier)
    }

    auto stage = [&](int t0) {                           // this warp's slice [t0, t0 + ST): rows past L are zero-filled
        for (int i = lane; i < ST * XR / 4; i += 32) {   // X_dbl rows are contiguous in global memory
            const bool ok = t0 + (i * 4) / XR < L;
            tm_cp_async16(xw + i * 4, ok ? xg + (long)t0 * XR + i * 4 : xg, ok);
        }
        for (int i = lane; i < ST * 4; i += 32) {        // u: ST rows of 32 channels = 4 x 16 bytes each
            const int r = i >> 2, v = i & 3;
            const bool ok = t0 + r < L;
            tm_cp_async16(uw + r * 32 + v * 8, ok ? ug + (long)(t0 + r) * D + v * 8 : ug, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    T* ybase = y + (long)b * H * W * D + dloc;
    const int nchunks = (L + CH - 1) / CH;
    stage(warp * ST);
    for (int c = 0; c < nchunks; ++c) {
        const int t0 = c * CH + warp * ST;
        const int ns = min(ST, L - t0);                  // live steps of this slice (<= 0: the slice is past the end of the row)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- phase A: local scan of the slice from h = 0
        float hl[NS], P[NS], yl[ST];
#pragma unroll
        for (int n = 0; n < NS; ++n) { hl[n] = 0.f; P[n] = 1.f; }
        auto phase_a = [&](auto mask_c) {
            constexpr bool MASK = decltype(mask_c)::value;
#pragma unroll
            for (int i0 = 0; i0 < ST; i0 += 4) {
                float dt[4], u[4], a[4][NS];
                load_steps4<T, NS, RDT, XR, 32, MASK>(xw + i0 * XR, uw + i0 * 32 + lane, nullptr, wdt, bias, ns - i0, dt, u);
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int n = 0; n < NS; ++n) a[q][n] = tm_ex2(dt[q] * A2[n]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float* xr = xw + (i0 + q) * XR + RDT;
                    const float du = dt[q] * u[q];
                    float y0 = Dd * u[q], y1 = 0.f;
#pragma unroll
                    for (int n = 0; n < NS; n += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(xr + n);
                        const float4 c4 = *reinterpret_cast<const float4*>(xr + NS + n);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
                        float gq[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            hl[n + j] = fmaf(a[q][n + j], hl[n + j], du * bb[j]);
                            if (j & 1) y1 = fmaf(hl[n + j], cc[j], y1); else y0 = fmaf(hl[n + j], cc[j], y0);
                            P[n + j] *= a[q][n + j];
                            gq[j] = cc[j] * P[n + j];
                        }
                        *reinterpret_cast<float4*>(gw + (i0 + q) * 32 * NS + n) = make_float4(gq[0], gq[1], gq[2], gq[3]);
                    }
                    yl[i0 + q] = y0 + y1;
                }
            }
        };
        if (ns >= ST) phase_a(std::false_type{}); else phase_a(std::true_type{});
        {
            float* ph = s_ph + ((c & 1) * TW + warp) * 2 * NS * 32 + lane;
#pragma unroll
            for (int n = 0; n < NS; ++n) { ph[n * 32] = P[n]; ph[(NS + n) * 32] = hl[n]; }
        }
        __syncwarp();                                    // every lane is done reading the staged slice
        if (c + 1 < nchunks) stage(t0 + CH);
        __syncthreads();                                 // slices (and the entry state) of chunk c are published
        // ---- fold: state entering this warp's slice
        float hin[NS];
        {
            const float* cy = s_cy + (c & 1) * NS * 32 + lane;
#pragma unroll
            for (int n = 0; n < NS; ++n) hin[n] = cy[n * 32];
            for (int j = 0; j < warp; ++j) {
                const float* ph = s_ph + ((c & 1) * TW + j) * 2 * NS * 32 + lane;
#pragma unroll
                for (int n = 0; n < NS; ++n) hin[n] = fmaf(ph[n * 32], hin[n], ph[(NS + n) * 32]);
            }
        }
        if (warp == TW - 1) {                            // exit state of the chunk = entry state of the next one
            float* cy = s_cy + ((c + 1) & 1) * NS * 32 + lane;
#pragma unroll
            for (int n = 0; n < NS; ++n) cy[n * 32] = fmaf(P[n], hin[n], hl[n]);
        }
        // ---- fix-up + EfficientMerge store
        if (ns > 0) {
            MergeWalk mw;
            mw.init(k, t0, H, W, D);
            auto fixup = [&](auto mask_c) {
                constexpr bool MASK = decltype(mask_c)::value;
#pragma unroll
                for (int i = 0; i < ST; ++i) {
                    if (!MASK || i < ns) {
                        float yv = yl[i], y2 = 0.f;
#pragma unroll
                        for (int n = 0; n < NS; n += 4) {
                            const float4 g4 = *reinterpret_cast<const float4*>(gw + i * 32 * NS + n);
                            yv = fmaf(g4.x, hin[n], yv); y2 = fmaf(g4.y, hin[n + 1], y2);
                            yv = fmaf(g4.z, hin[n + 2], yv); y2 = fmaf(g4.w, hin[n + 3], y2);
                        }
                        fd_st(ybase + mw.off, yv + y2);
                        mw.next();
                    }
                }
            };
            if (ns >= ST) fixup(std::false_type{}); else fixup(std::true_type{});
        }
    }
}

template <typename T, int NS, int RDT, int ST, int TW>
int scan_tw_launch(const void* u_tm, const float* xdbl, const float* A, const float* dt_w, const float* dt_bias, const float* Dskip,
                   void* y, int B, int D, int H, int W, cudaStream_t st) {
    const int L = (H / 2) * (W / 2);
    constexpr int XR = RDT + 2 * NS;
    const size_t smem = ((size_t)TW * ST * 32 * NS + (size_t)TW * ST * XR + (size_t)2 * TW * 2 * NS * 32 + 2 * NS * 32) * sizeof(float) +
                        (size_t)TW * ST * 32 * sizeof(T);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(scan_tw_kernel<T, NS, RDT, ST, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    scan_tw_kernel<T, NS, RDT, ST, TW><<<dim3(D / 32, B * 4), TW * 32, smem, st>>>((const T*)u_tm, xdbl, A, dt_w, dt_bias, Dskip, (T*)y, D, L, H, W);
    FD_LAUNCH_CHECK();
    return 0;
}

// =========================================================================================================
// K3c: the time-sliced scan rewritten around PACKED fp32 arithmetic (fma / mul .f32x2 = FFMA2 / FMUL2 on sm_100a).
// Measured on B200 (tools/probes/pipe_rates.cu): a warp-wide FFMA2 occupies the FMA pipe for 2 clk but ONE issue slot,
// MUFU.EX2 runs at 16 lanes / clk / SM (one warp instruction per 8 clk and scheduler) — so a state update
// (FMUL, EX2, FMUL, FFMA, FFMA [+ FMUL, FMUL, FFMA of the time-slice fix-up]) is MUFU-bound as soon as it costs fewer than
// 8 issue slots, and K3b (ncu: 96 instructions per channel-step, 69 % of the issue slots busy) was bound by issue.  Here the
// states of a lane live as pairs (n, n + 1) in 64-bit registers: decay argument, input term, recurrence, output dot product,
// running product and fix-up row are one packed instruction per PAIR, the softplus of two steps shares its non-MUFU
// arithmetic the same way, B / C / dt-input rows arrive as LDS.128 = two ready-made pairs, and (P, h) of a slice travel
// through shared memory as 64-bit words.  Same algorithm, same association order as K3b (results agree to the last
// re-association), and only for geometries where no slice is ragged or straddles an EfficientMerge row: L % (TW ST) == 0
// and the merge row length (W/2 resp. H/2) a multiple of ST — everything else takes K3b.
// VAR bit 0: y_local kept in shared memory instead of ST registers; bit 1: delta of the whole slice formed ahead of the recurrence
// (ST / 2 independent softplus chains in flight instead of SB / 2); bit 2: two half-slices per warp scanned in lockstep
template <typename T, int NS, int RDT, int ST, int TW, int VAR>
__global__ void __launch_bounds__(TW * 32, (TW == 8 ? 2 : 4)) scan_tw2_kernel(
    const T* __restrict__ u_tm, const float* __restrict__ xdbl, const float* __restrict__ A, const float* __restrict__ dt_w,
    const float* __restrict__ dt_bias, const float* __restrict__ Dskip, T* __restrict__ y, int D, int L, int H, int W, int nseg,
    uint32_t* __restrict__ chain_ws) {
    constexpr int XR = RDT + 2 * NS;
    constexpr int CH = TW * ST;                          // steps per chunk
    constexpr int NP = NS / 2, RP = RDT / 2, NQ = NS / 4;
    constexpr int SB = NS <= 4 ? 4 : 2;                  // steps whose decay factors are issued ahead of the recurrence
    constexpr bool YL_SMEM = VAR & 1, DT_AHEAD = (VAR & 2) != 0;
    constexpr int NSL = (VAR & 4) ? 2 : 1, STS = ST / NSL;  // half-slices scanned in lockstep by one warp
    constexpr bool G_TMEM = (VAR & 8) != 0;              // fix-up rows in tensor memory instead of shared memory
    constexpr int TCOLS = (TW / 4) * ST * NS;            // TMEM columns of the block: per lane quarter, TW / 4 warps x ST steps x NS
    static_assert(!G_TMEM || (NSL == 1 && !YL_SMEM && (SB * NS == 16) && TCOLS >= 32 && (TCOLS & (TCOLS - 1)) == 0), "TMEM rows: 16 registers per step block");
    constexpr int GS = G_TMEM ? 0 : NQ * 128 + (YL_SMEM ? 32 : 0);   // floats per step in s_g: NQ planes of one float4 per lane [+ one y_local per lane]
    static_assert(ST % SB == 0 && SB % NSL == 0 && NS % 4 == 0 && RDT % 4 == 0 && RDT > 0, "step blocks, float4 rows");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_g = reinterpret_cast<float*>(smem_raw);     // [TW][ST][GS]   fix-up rows C * prod(a)
    float* s_x = s_g + TW * ST * GS;                // [TW][ST][XR]
    u64* s_ph = reinterpret_cast<u64*>(s_x + TW * ST * XR);   // [2][TW][2 NP][32]  (P pairs | h pairs) of a slice
    u64* s_cy = s_ph + 2 * TW * 2 * NP * 32;             // [2][NP][32]           state entering a chunk
    T* s_u = reinterpret_cast<T*>(s_cy + 2 * NP * 32);   // [TW][ST][32]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_u + TW * ST * 32);   // TMEM base address of the block's allocation
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Chained segments (nseg > 1, grid.z = nseg): a row of L steps is cut into nseg segments that are separate blocks; the state
    // leaving a segment travels to its successor through global memory behind a release / acquire flag.  256 rows on 2 x 148
    // block slots left 40 SMs with one block (done at 0.59 of the kernel's time, ncu: l1tex__cycles_active min / max); with short
    // blocks the hardware scheduler refills the slots and every SM works until the last segments.  Which (segment, row) a block
    // takes comes from a ticket drawn in START order, segment-major — the predecessor of a block holds a smaller ticket, so it is
    // resident or finished whenever its successor waits for it: no deadlock whatever order the blocks are dispatched in.
    // Arithmetic is untouched (the carried state is the fp32 s_cy row of the unchained kernel): results are bit-identical.
    int bx = (int)blockIdx.x, bk = (int)blockIdx.y, seg = 0, link = 0;
    if (nseg > 1) {
        if (threadIdx.x == 0) {
            const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
            const uint32_t t = atomicAdd(chain_ws, 1u);
            if (t == total - 1) atomicExch(chain_ws, 0u);          // every ticket of this launch is drawn: ready for the next launch
            s_tmem[1] = t;
        }
        __syncthreads();
        const uint32_t t = s_tmem[1], units = gridDim.x * gridDim.y;
        seg = (int)(t / units);
        const uint32_t un = t - (uint32_t)seg * units;
        bk = (int)(un / gridDim.x);
        bx = (int)(un - (uint32_t)bk * gridDim.x);
        link = (int)un * (nseg - 1) + seg;                          // link seg -> seg + 1 of this row (seg - 1 -> seg is link - 1)
    }
    uint32_t* const chain_flag = chain_ws + 1;
    u64* const chain_carry = reinterpret_cast<u64*>(chain_ws + ((1 + gridDim.x * gridDim.y * (uint32_t)(nseg - 1) + 3u) & ~3u));
    const int k = bk & 3, b = bk >> 2;
    const int ch0 = bx * 32;
    const int dloc = ch0 + lane;
    const int d = k * D + dloc;
    const T* ug = u_tm + (long)bk * L * D + ch0;
    const float* xg = xdbl + (long)bk * L * XR;
    uint32_t tg = 0;                                     // this warp's TMEM window: its lane quarter, ST * NS columns
    if constexpr (G_TMEM) {
        // Tensor memory as a per-lane scratchpad.  The fix-up rows are lane-private (written in phase A, read back by the same
        // lane after the barrier): through shared memory they were 8 of the kernel's 19 wavefronts per warp-step on the one
        // shared-memory data pipe (profiles/r2_ncu_scan_tw2.txt); tcgen05.st / tcgen05.ld move 16 registers per instruction
        // on TMEM's own path.  A warp may touch the 32 TMEM lanes of its quarter (warp % 4); warps w and w + 4 share a quarter
        // and take different columns.
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_tmem)), "r"(TCOLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tg = *s_tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * ST * NS);
    }
    float* gw = s_g + warp * ST * GS + lane * 4;
    float* ylw = s_g + warp * ST * GS + NQ * 128 + lane;
    float* xw = s_x + warp * ST * XR;
    T* uw = s_u + warp * ST * 32;

    u64 A2[NP], wdt[RP];
#pragma unroll
    for (int j = 0; j < NP; ++j) A2[j] = f2_pack(A[(long)d * NS + 2 * j] * 1.4426950408889634f, A[(long)d * NS + 2 * j + 1] * 1.4426950408889634f);
#pragma unroll
    for (int r = 0; r < RP; ++r) wdt[r] = f2_pack(dt_w[(long)d * RDT + 2 * r], dt_w[(long)d * RDT + 2 * r + 1]);
    const u64 bias2 = f2_pack(dt_bias[d], 0.f);
    const float Dd = Dskip[d];
    if (warp == 0 && seg == 0) {
#pragma unroll
        for (int j = 0; j < NP; ++j) s_cy[j * 32 + lane] = 0ull;          // entry state of chunk 0 (published by the first barrier)
    }

    auto stage = [&](int t0) {                           // this warp's slice [t0, t0 + ST)
        for (int i = lane; i < ST * XR / 4; i += 32) tm_cp_async16(xw + i * 4, xg + (long)t0 * XR + i * 4, true);
        for (int i = lane; i < ST * 4; i += 32) {        // u: ST rows of 32 channels = 4 x 16 bytes each
            const int r = i >> 2, v = i & 3;
            tm_cp_async16(uw + r * 32 + v * 8, ug + (long)(t0 + r) * D + v * 8, true);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto dt_raw = [&](int i) {                           // dt_proj of step i of the staged slice (+ bias), before the softplus
        u64 p = bias2;
#pragma unroll
        for (int r = 0; r < RP; r += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xw + i * XR + 2 * r);
            p = f2_fma(wdt[r], v.x, p);
            p = f2_fma(wdt[r + 1], v.y, p);
        }
        float plo, phi;
        f2_unpack(p, plo, phi);
        return plo + phi;
    };

    // EfficientMerge position of this warp's slice, advanced by CH steps per chunk (no division in the loop): slices never
    // straddle a merge row here, so inside a slice the store address moves by a constant stride.
    const int col = k & 1, kx = k >> 1;
    const int mdiv = col ? (H >> 1) : (W >> 1);
    const int nchunks = (L / CH) / nseg;                 // chunks of this block's segment
    const int step0 = seg * nchunks * CH;                // first step of the segment
    int mq = (step0 + warp * ST) / mdiv, mr = (step0 + warp * ST) - mq * mdiv;
    const int inc = col ? 2 * W * D : 2 * D;
    T* ybase = y + (long)b * H * W * D + dloc;
    stage(step0 + warp * ST);
    for (int c = 0; c < nchunks; ++c) {
        const int t0 = step0 + c * CH + warp * ST;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- phase A: local scan(s) of the slice from h = 0.  NSL = 2: the slice is two half-slices scanned in lockstep (two
        // independent recurrences per lane: twice the instruction-level parallelism for 2 NS more registers)
        u64 hl[NSL][NP], P[NSL][NP];
        float yl[YL_SMEM ? 1 : ST];
        float dta[DT_AHEAD ? ST : 1];
#pragma unroll
        for (int s = 0; s < NSL; ++s)
#pragma unroll
            for (int j = 0; j < NP; ++j) { hl[s][j] = 0ull; P[s][j] = f2_pack(1.f, 1.f); }
        if constexpr (DT_AHEAD) {
#pragma unroll
            for (int i = 0; i < ST; ++i) dta[i] = dt_raw(i);
#pragma unroll
            for (int i = 0; i < ST; i += 2) tm_softplus2(dta[i], dta[i + 1], dta[i], dta[i + 1]);
        }
#pragma unroll
        for (int i0 = 0; i0 < STS; i0 += SB / NSL) {
            float dt[SB], u[SB];
#pragma unroll
            for (int q = 0; q < SB; ++q) {               // in-flight step q: half-slice q % NSL, step i0 + q / NSL of it
                const int st = (q % NSL) * STS + i0 + q / NSL;
                dt[q] = DT_AHEAD ? dta[DT_AHEAD ? st : 0] : dt_raw(st);
                u[q] = tm_ld16<T>(uw + st * 32 + lane);
            }
            if constexpr (!DT_AHEAD) {
#pragma unroll
                for (int q = 0; q < SB; q += 2) tm_softplus2(dt[q], dt[q + 1], dt[q], dt[q + 1]);
            }
            u64 a[SB][NP];
#pragma unroll
            for (int q = 0; q < SB; ++q) {
                const u64 dt2 = f2_pack(dt[q], dt[q]);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    float e0, e1;
                    f2_unpack(f2_mul(dt2, A2[j]), e0, e1);
                    a[q][j] = f2_pack(tm_ex2(e0), tm_ex2(e1));
                }
            }
            float gt[G_TMEM ? 16 : 1];
#pragma unroll
            for (int q = 0; q < SB; ++q) {
                const int sl = q % NSL, st = sl * STS + i0 + q / NSL;
                const float* xr = xw + st * XR + RDT;
                const float du = dt[q] * u[q];
                const u64 du2 = f2_pack(du, du);
                u64 ya = 0ull;
#pragma unroll
                for (int jq = 0; jq < NQ; ++jq) {
                    const ulonglong2 b2 = *reinterpret_cast<const ulonglong2*>(xr + 4 * jq);
                    const ulonglong2 c2 = *reinterpret_cast<const ulonglong2*>(xr + NS + 4 * jq);
                    const int j = 2 * jq;
                    hl[sl][j] = f2_fma(a[q][j], hl[sl][j], f2_mul(du2, b2.x));
                    hl[sl][j + 1] = f2_fma(a[q][j + 1], hl[sl][j + 1], f2_mul(du2, b2.y));
                    ya = jq == 0 ? f2_mul(hl[sl][j], c2.x) : f2_fma(hl[sl][j], c2.x, ya);
                    ya = f2_fma(hl[sl][j + 1], c2.y, ya);
                    P[sl][j] = f2_mul(P[sl][j], a[q][j]);
                    P[sl][j + 1] = f2_mul(P[sl][j + 1], a[q][j + 1]);
                    ulonglong2 g2;
                    g2.x = f2_mul(c2.x, P[sl][j]);
                    g2.y = f2_mul(c2.y, P[sl][j + 1]);
                    if constexpr (G_TMEM) {
                        f2_unpack(g2.x, gt[q * NS + 4 * jq], gt[q * NS + 4 * jq + 1]);
                        f2_unpack(g2.y, gt[q * NS + 4 * jq + 2], gt[q * NS + 4 * jq + 3]);
                    } else {
                        *reinterpret_cast<ulonglong2*>(gw + st * GS + jq * 128) = g2;
                    }
                }
                float ylo, yhi;
                f2_unpack(ya, ylo, yhi);
                const float yv = fmaf(Dd, u[q], ylo) + yhi;
                if constexpr (YL_SMEM) ylw[st * GS] = yv;
                else yl[YL_SMEM ? 0 : st] = yv;
            }
            if constexpr (G_TMEM) {                      // the SB steps' rows (16 registers) in one tcgen05.st: columns [i0 NS, +16)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tg + (uint32_t)(i0 * NS)),
                             "f"(gt[0]), "f"(gt[1]), "f"(gt[2]), "f"(gt[3]), "f"(gt[4]), "f"(gt[5]), "f"(gt[6]), "f"(gt[7]), "f"(gt[8]), "f"(gt[9]),
                             "f"(gt[10]), "f"(gt[11]), "f"(gt[12]), "f"(gt[13]), "f"(gt[14]), "f"(gt[15]) : "memory");
            }
        }
        if constexpr (G_TMEM) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        {                                                // (P, h) of the whole slice: the half-slices composed
            u64* ph = s_ph + ((c & 1) * TW + warp) * 2 * NP * 32 + lane;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                ph[j * 32] = NSL == 2 ? f2_mul(P[0][j], P[NSL - 1][j]) : P[0][j];
                ph[(NP + j) * 32] = NSL == 2 ? f2_fma(P[NSL - 1][j], hl[0][j], hl[NSL - 1][j]) : hl[0][j];
            }
        }
        __syncwarp();                                    // every lane is done reading the staged slice
        if (c + 1 < nchunks) stage(t0 + CH);
        if (seg > 0 && c == 0 && warp == 0) {            // entry state of the segment: wait for the predecessor (phase A is done already)
            uint32_t* flag = chain_flag + (link - 1);
            uint32_t ready, polls = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(ready) : "l"(flag) : "memory");
                if (!ready) {
                    __nanosleep(100);
                    if (++polls > (1u << 24)) __trap();  // > 1.5 s for a hand-over that takes microseconds: a workspace that was not
                }                                        // zero-filled (or is shared by two launches) — fail loudly instead of hanging
            } while (!ready);
            const u64* cg = chain_carry + (size_t)(link - 1) * NP * 32 + lane;
#pragma unroll
            for (int j = 0; j < NP; ++j) s_cy[j * 32 + lane] = __ldcg(cg + j * 32);
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(flag) = 0u;      // one consumer per flag: clean for the next launch
        }
        __syncthreads();                                 // slices (and the entry state) of chunk c are published
        // ---- fold: state entering this warp's slice (and its second half)
        u64 hin[NSL][NP];
        {
            const u64* cy = s_cy + (c & 1) * NP * 32 + lane;
#pragma unroll
            for (int j = 0; j < NP; ++j) hin[0][j] = cy[j * 32];
            for (int jw = 0; jw < warp; ++jw) {
                const u64* ph = s_ph + ((c & 1) * TW + jw) * 2 * NP * 32 + lane;
#pragma unroll
                for (int j = 0; j < NP; ++j) hin[0][j] = f2_fma(ph[j * 32], hin[0][j], ph[(NP + j) * 32]);
            }
            if constexpr (NSL == 2) {
#pragma unroll
                for (int j = 0; j < NP; ++j) hin[NSL - 1][j] = f2_fma(P[0][j], hin[0][j], hl[0][j]);
            }
        }
        if (warp == TW - 1) {                            // exit state of the chunk = entry state of the next one
            u64* cy = s_cy + ((c + 1) & 1) * NP * 32 + lane;
            const bool hand_over = c == nchunks - 1 && seg + 1 < nseg;         // the successor segment is another block
            u64* cg = chain_carry + (size_t)link * NP * 32 + lane;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const u64 hx = f2_fma(P[NSL - 1][j], hin[NSL - 1][j], hl[NSL - 1][j]);
                cy[j * 32] = hx;
                if (hand_over) __stcg(cg + j * 32, hx);
            }
            if (hand_over) {
                __threadfence();
                __syncwarp();
                if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(chain_flag + link), "r"(1u) : "memory");
            }
        }
        // ---- fix-up + EfficientMerge store
        {
            T* yp = ybase + (col ? ((long)(2 * mr + 1) * W + 2 * mq + kx) * D : ((long)(2 * mq) * W + 2 * mr + kx) * D);
            float gl[G_TMEM ? 16 : 1];
#pragma unroll
            for (int i = 0; i < ST; ++i) {
                const int sl = i / STS;
                if constexpr (G_TMEM) {
                    if (i % SB == 0) {
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                     : "=f"(gl[0]), "=f"(gl[1]), "=f"(gl[2]), "=f"(gl[3]), "=f"(gl[4]), "=f"(gl[5]), "=f"(gl[6]), "=f"(gl[7]), "=f"(gl[8]),
                                       "=f"(gl[9]), "=f"(gl[10]), "=f"(gl[11]), "=f"(gl[12]), "=f"(gl[13]), "=f"(gl[14]), "=f"(gl[15])
                                     : "r"(tg + (uint32_t)(i * NS)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                }
                u64 f = 0ull;
#pragma unroll
                for (int jq = 0; jq < NQ; ++jq) {
                    ulonglong2 g2;
                    if constexpr (G_TMEM) {
                        g2.x = f2_pack(gl[(i % SB) * NS + 4 * jq], gl[(i % SB) * NS + 4 * jq + 1]);
                        g2.y = f2_pack(gl[(i % SB) * NS + 4 * jq + 2], gl[(i % SB) * NS + 4 * jq + 3]);
                    } else {
                        g2 = *reinterpret_cast<const ulonglong2*>(gw + i * GS + jq * 128);
                    }
                    f = jq == 0 ? f2_mul(g2.x, hin[sl][0]) : f2_fma(g2.x, hin[sl][2 * jq], f);
                    f = f2_fma(g2.y, hin[sl][2 * jq + 1], f);
                }
                float flo, fhi;
                f2_unpack(f, flo, fhi);
                const float y0 = YL_SMEM ? ylw[i * GS] : yl[YL_SMEM ? 0 : i];
                fd_st(yp + (long)i * inc, (y0 + flo) + fhi);
            }
            mr += CH;
            while (mr >= mdiv) { mr -= mdiv; ++mq; }
        }
    }
    if constexpr (G_TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s_tmem), "r"(TCOLS));
    }
}

template <typename T, int NS, int RDT, int ST, int TW, int VAR>
int scan_tw2_launch_v(const void* u_tm, const float* xdbl, const float* A, const float* dt_w, const float* dt_bias, const float* Dskip,
                      void* y, int B, int D, int H, int W, int nseg, uint32_t* chain_ws, cudaStream_t st) {
    const int L = (H / 2) * (W / 2);
    constexpr int XR = RDT + 2 * NS, GS = (VAR & 8) ? 0 : NS * 32 + ((VAR & 1) ? 32 : 0);
    const size_t smem = ((size_t)TW * ST * GS + (size_t)TW * ST * XR) * sizeof(float) +
                        ((size_t)2 * TW * NS * 32 + NS * 32) * sizeof(u64) + (size_t)TW * ST * 32 * sizeof(T) + 16;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(scan_tw2_kernel<T, NS, RDT, ST, TW, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    scan_tw2_kernel<T, NS, RDT, ST, TW, VAR><<<dim3(D / 32, B * 4, nseg), TW * 32, smem, st>>>((const T*)u_tm, xdbl, A, dt_w, dt_bias, Dskip, (T*)y, D, L, H, W,
                                                                                              nseg, chain_ws);
    FD_LAUNCH_CHECK();
    return 0;
}
template <typename T, int NS, int RDT, int ST, int TW>
int scan_tw2_launch(const void* u_tm, const float* xdbl, const float* A, const float* dt_w, const float* dt_bias, const float* Dskip,
                    void* y, int B, int D, int H, int W, int nseg, uint32_t* chain_ws, cudaStream_t st) {
    // measured at B = 16 (profiles/r2_scan_tw2_notes.txt): delta formed ahead (VAR 2) 1705 us vs 1730 at the level-0 shape; y_local in
    // shared memory (bit 0) 1838 and two half-slices in lockstep (bit 2) 1738 — the kernel is bound by the shared-memory data pipe
    // (19 wavefronts per warp-step), not by registers or by the recurrence's latency, so neither helps
    static const bool smem_rows = getenv("FD_SCAN_TW2_TMEM") && atoi(getenv("FD_SCAN_TW2_TMEM")) == 0;
    if (smem_rows) return scan_tw2_launch_v<T, NS, RDT, ST, TW, 2>(u_tm, xdbl, A, dt_w, dt_bias, Dskip, y, B, D, H, W, nseg, chain_ws, st);
    return scan_tw2_launch_v<T, NS, RDT, ST, TW, 10>(u_tm, xdbl, A, dt_w, dt_bias, Dskip, y, B, D, H, W, nseg, chain_ws, st);
}

// K3c applies when no slice is ragged and no slice straddles an EfficientMerge row
static bool tw2_ok(int H, int W, int ST, int TW) {
    static const int off = getenv("FD_SCAN_TW2") ? atoi(getenv("FD_SCAN_TW2")) == 0 : 0;
    const int L = (H / 2) * (W / 2);
    return !off && L % (TW * ST) == 0 && (W / 2) % ST == 0 && (H / 2) % ST == 0;
}

// Chained segments of the packed time-sliced kernel (see scan_tw2_kernel): how many blocks a row is cut into.  Measured at B = 16
// (gpurun_out/chain_sweep.log -> profiles/r2_scan_chain_sweep.txt): segments of 16 chunks are the optimum at all three time-sliced
// shapes (level 0 N4: 1637 us unchained, 1441 with 32 segments; N8 R4: 714 -> 649 with 16; N8 R8 x4 warps: 1357 -> 1160 with 32);
// 2 segments are slower than none (the waiting successor holds a block slot), 8-chunk segments pay the block prologue too often.
// FD_SCAN_CHAIN = 0 / 1 switches chaining off, any other value is the segment count.  0 = the geometry is not chained.
static int chain_segments(int H, int W, int ST, int TW) {
    const char* env = getenv("FD_SCAN_CHAIN");           // read per call (plan time and launch, never inside a graph replay): sweepable
    if (!tw2_ok(H, W, ST, TW)) return 0;
    const int nchunks = (H / 2) * (W / 2) / (TW * ST);
    int n = env ? atoi(env) : nchunks / 16;
    if (n <= 1) return 0;
    while (n & (n - 1)) n &= n - 1;                      // power of two (nchunks is one for the image sizes of interest; else halve)
    while (n > 1 && (nchunks % n || nchunks / n < 4)) n >>= 1;
    return n >= (env ? 2 : 4) ? n : 0;
}
static long chain_ws_floats(int B, int D, int dstate, int nseg) {       // ticket counter | flags (16-byte padded) | carried states
    const long links = (long)B * 4 * (D / 32) * (nseg - 1);
    return ((1 + links + 3) & ~3L) + links * dstate * 32;
}
static int tw_steps(int dstate) { return dstate == 4 ? 16 : 8; }        // ST of the SCTW2_CASE table below

// Which scan a geometry gets: 0 = segmented channel-per-lane (scan_tm_kernel), 8 / 4 = time-sliced with that many warps.
// The time-sliced kernel needs ALL its blocks resident at once (a block walks a whole row): 2 blocks per SM with 8 warps,
// 4 with 4 warps; it only exists for the fused-dt small-state levels.
int pick_time_warps(int B, int D, int H2, int W2, int NS, int RDT) {
    const int L = H2 * W2;
    static const int forced = getenv("FD_SCAN_TW") ? atoi(getenv("FD_SCAN_TW")) : -1;
    if (RDT == 0 || NS > 8) return 0;
    if (forced >= 0) return forced;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long blocks = (long)B * 4 * (D / 32);
    if (L < 2048) return 0;
    // measured (profiles/r2_scan_tm_variants.json): with >= 512 channel-warps one unsegmented pass is faster than either
    // form of time parallelism (16x256x16384 N8: 1.57 ms vs 1.95 time-sliced x4 vs 1.73 with 4 segments)
    if (blocks <= 2L * sms) return 8;
    // 16x256x16384 N8 R8 (512 blocks): the packed time-sliced kernel with 4 warps per block (4 blocks per SM) 1.49 ms vs 1.57 unsegmented
    if (NS == 8 && blocks <= 4L * sms && tw2_ok(2 * H2, 2 * W2, 8, 4)) return 4;
    return 0;
}

int pick_segments(int B, int D, int L) {
    // enough warps to fill the machine (>= ~16 per SM), segments not shorter than 2048 steps (the carry pass walks ~the
    // memory length of the slowest channel of a block per segment, so short segments pay it proportionally more often)
    static const int forced = getenv("FD_SCAN_SEGMENTS") ? atoi(getenv("FD_SCAN_SEGMENTS")) : 0;
    if (forced > 0) return forced;
    const long base_warps = (long)B * 4 * D / 32;
    if (base_warps >= 512) return 1;                    // ~1 warp per scheduler: the carry pass would cost more than it buys
    int S = 1;
    while (S < 64 && base_warps * S < 148L * 24 && L / (2 * S) >= 2048) S *= 2;
    return S;
}

template <typename T, int NS, int RDT>
int scan_tm_launch(const void* u_tm, const void* dts_tm, const float* xdbl, const float* A, const float* dt_w, const float* dt_bias,
                   const float* Dskip, float* carry, size_t carry_floats, void* y, int B, int D, int H, int W, int S,
                   cudaStream_t st) {
    const int L = (H / 2) * (W / 2);
    constexpr int XR = RDT + 2 * NS;
    const size_t smem = (size_t)2 * SC_T * XR * sizeof(float) + (size_t)(RDT == 0 ? 4 : 2) * SC_T * SC_CHB * sizeof(T);
    int seg_len = ((L + S - 1) / S + SC_T - 1) / SC_T * SC_T;
    S = (L + seg_len - 1) / seg_len;
    if (S > 1 && (!carry || carry_floats < (size_t)B * 4 * S * 2 * NS * D)) return FD_ERR_BAD_ARGUMENT;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(scan_tm_kernel<T, NS, RDT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_tm_kernel<T, NS, RDT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    if (S > 1) {
        scan_tm_kernel<T, NS, RDT, true><<<dim3(S - 1, D / SC_CHB, B * 4), SC_CHB, smem, st>>>(
            (const T*)u_tm, (const T*)dts_tm, xdbl, A, dt_w, dt_bias, Dskip, carry, (T*)y, D, L, H, W, S, seg_len);
        FD_LAUNCH_CHECK();
    }
    scan_tm_kernel<T, NS, RDT, false><<<dim3(S, D / SC_CHB, B * 4), SC_CHB, smem, st>>>(
        (const T*)u_tm, (const T*)dts_tm, xdbl, A, dt_w, dt_bias, Dskip, carry, (T*)y, D, L, H, W, S, seg_len);
    FD_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int fd_dwconv3x3_silu_tm(const void* xz, int ld, const float* w_tap_major, const float* bias, void* xs_tm, int B, int H,
                                    int W, int D, int dtype, cudaStream_t stream) {
    if (!xz || !w_tap_major || !xs_tm || B <= 0 || H <= 0 || W <= 0 || D <= 0 || ld < D) return FD_ERR_BAD_ARGUMENT;
    if ((H & 1) || (W & 1) || D % TM_V || ld % TM_V || ((uintptr_t)xz & 7) || ((uintptr_t)xs_tm & 7) || ((uintptr_t)w_tap_major & 15) ||
        ((uintptr_t)bias & 15))
        return FD_ERR_UNSUPPORTED;
    static const int pf = getenv("FD_DWCONV_PF") ? atoi(getenv("FD_DWCONV_PF")) : 4;
    const long pairs = (long)(W / 2) * (D / TM_V);
    dim3 grid((unsigned)fd_cdiv(pairs, 256), (unsigned)fd_cdiv(H, TM_RY), (unsigned)B);
    if (dtype == FD_BF16) dwconv_tm_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)xz, ld, w_tap_major, bias, (__nv_bfloat16*)xs_tm, H, W, D, pf);
    else if (dtype == FD_F16) dwconv_tm_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)xz, ld, w_tap_major, bias, (__half*)xs_tm, H, W, D, pf);
    else return FD_ERR_UNSUPPORTED;
    FD_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int x_proj_tm_launch(const void* xs_tm, const void* xw16, float* xdbl, const void* dw16, void* dts_tm, const float* dt_bias,
                            int B, int D, int L, int R, int N, int Rp, int fuse_dt, cudaStream_t stream) {
    const int CC = R + 2 * N, CCp = (CC + 15) / 16 * 16, NT = CCp / 8;
    dim3 grid((unsigned)fd_cdiv(L, XM_STEPS), (unsigned)(B * 4));
    const size_t smem = (size_t)XM_STAGES * (XM_STEPS + CCp) * XM_LD * sizeof(T);
    const int col0 = fuse_dt ? 0 : R, out_ld = fuse_dt ? CC : 2 * N;
#define XPTM_CASE(NTV)                                                                                                            \
    if (NT == NTV) {                                                                                                              \
        if (fuse_dt) {                                                                                                            \
            static bool attr = false;                                                                                             \
            if (!attr) {                                                                                                          \
                cudaError_t e = cudaFuncSetAttribute(x_proj_tm_kernel<T, NTV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                if (e != cudaSuccess) return (int)e;                                                                              \
                attr = true;                                                                                                      \
            }                                                                                                                     \
            x_proj_tm_kernel<T, NTV, false><<<grid, 256, smem, stream>>>((const T*)xs_tm, (const T*)xw16, xdbl, D, L, CC, col0, out_ld, \
                                                                          nullptr, nullptr, nullptr, 0);                         \
        } else {                                                                                                                  \
            static bool attr = false;                                                                                             \
            if (!attr) {                                                                                                          \
                cudaError_t e = cudaFuncSetAttribute(x_proj_tm_kernel<T, NTV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                if (e != cudaSuccess) return (int)e;                                                                              \
                attr = true;                                                                                                      \
            }                                                                                                                     \
            x_proj_tm_kernel<T, NTV, true><<<grid, 256, smem, stream>>>((const T*)xs_tm, (const T*)xw16, xdbl, D, L, CC, col0, out_ld, \
                                                                         (const T*)dw16, (T*)dts_tm, dt_bias, Rp);               \
        }                                                                                                                         \
        FD_LAUNCH_CHECK();                                                                                                        \
        return 0;                                                                                                                 \
    }
    XPTM_CASE(2) XPTM_CASE(4) XPTM_CASE(6) XPTM_CASE(8) XPTM_CASE(10) XPTM_CASE(12)
#undef XPTM_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_x_proj_tm(const void* xs_tm, const void* xw16, float* xdbl_tm, const void* dw16, void* dts_tm, const float* dt_bias,
                            int B, int D, int L, int R, int N, int Rp, int fuse_dt, int dtype, cudaStream_t stream) {
    if (!xs_tm || !xw16 || !xdbl_tm || B <= 0 || D <= 0 || L <= 0 || R <= 0 || N <= 0) return FD_ERR_BAD_ARGUMENT;
    if (!fuse_dt && (!dw16 || !dts_tm || !dt_bias)) return FD_ERR_BAD_ARGUMENT;
    if (D % XM_KC || R + 2 * N > 96 || (R & 1) || (N & 1) || (((uintptr_t)xs_tm | (uintptr_t)xw16) & 15) || ((uintptr_t)xdbl_tm & 7))
        return FD_ERR_UNSUPPORTED;
    if (!fuse_dt && ((Rp != 16 && Rp != 32) || R > Rp || Rp > (R + 2 * N + 15) / 16 * 16 || ((uintptr_t)dw16 & 3) || ((uintptr_t)dts_tm & 15) ||
                     ((uintptr_t)dt_bias & 7)))
        return FD_ERR_UNSUPPORTED;
    if (dtype == FD_BF16) return x_proj_tm_launch<__nv_bfloat16>(xs_tm, xw16, xdbl_tm, dw16, dts_tm, dt_bias, B, D, L, R, N, Rp, fuse_dt, stream);
    if (dtype == FD_F16) return x_proj_tm_launch<__half>(xs_tm, xw16, xdbl_tm, dw16, dts_tm, dt_bias, B, D, L, R, N, Rp, fuse_dt, stream);
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_scan_tm_plan(int B, int D, int H, int W, int dstate, int dt_rank_fused) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    const int L = (H / 2) * (W / 2);
    const int tw = pick_time_warps(B, D, H / 2, W / 2, dstate, dt_rank_fused);
    if (tw) return -tw;
    int S = pick_segments(B, D, L);
    const int seg_len = ((L + S - 1) / S + SC_T - 1) / SC_T * SC_T;
    return (L + seg_len - 1) / seg_len;
}

extern "C" int fd_scan_tm_segments(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return FD_ERR_BAD_ARGUMENT;
    const int L = (H / 2) * (W / 2);
    int S = pick_segments(B, D, L);
    const int seg_len = ((L + S - 1) / S + SC_T - 1) / SC_T * SC_T;
    return (L + seg_len - 1) / seg_len;
}

extern "C" int fd_selective_scan_tm(const void* u_tm, const void* dts_tm, const float* xdbl_tm, const float* A, const float* dt_w,
                                    const float* dt_bias, const float* D_skip, float* carry_ws, long carry_floats, void* y_nhwc,
                                    int B, int D, int H, int W, int dstate, int dt_rank_fused, int segments, int io_dtype,
                                    cudaStream_t stream) {
    if (!u_tm || !xdbl_tm || !A || !D_skip || !y_nhwc || B <= 0 || D <= 0 || H <= 0 || W <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((H & 1) || (W & 1)) return FD_ERR_BAD_ARGUMENT;
    if (dt_rank_fused ? (!dt_w || !dt_bias) : !dts_tm) return FD_ERR_BAD_ARGUMENT;
    if (D % SC_CHB || (((uintptr_t)u_tm | (uintptr_t)dts_tm | (uintptr_t)xdbl_tm) & 15)) return FD_ERR_UNSUPPORTED;
    const int L = (H / 2) * (W / 2);
    if (segments <= 0) {                                 // 0 = automatic: time-sliced kernel where rows are few and long; -8 / -4 = forced
        // -8 / -4: time-sliced kernel, packed form (K3c) where the geometry allows; -1008 / -1004: K3b forced
        const bool force_b = segments <= -1000;
        const int tw = segments < 0 ? -(segments % 1000) : pick_time_warps(B, D, H / 2, W / 2, dstate, dt_rank_fused);
        if (segments < 0 && tw != 8 && tw != 4) return FD_ERR_BAD_ARGUMENT;
#define SCTW2_CASE(NSV, RV, STV, TWV)                                                                                              \
    if (!force_b && tw == TWV && dstate == NSV && dt_rank_fused == RV && tw2_ok(H, W, STV, TWV)) {                                 \
        if (io_dtype == FD_BF16) return scan_tw2_launch<__nv_bfloat16, NSV, RV, STV, TWV>(u_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, y_nhwc, B, D, H, W, 1, nullptr, stream); \
        if (io_dtype == FD_F16) return scan_tw2_launch<__half, NSV, RV, STV, TWV>(u_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, y_nhwc, B, D, H, W, 1, nullptr, stream); \
        return FD_ERR_UNSUPPORTED;                                                                                                 \
    }
        SCTW2_CASE(4, 4, 16, 8) SCTW2_CASE(4, 4, 16, 4) SCTW2_CASE(8, 4, 8, 8) SCTW2_CASE(8, 4, 8, 4) SCTW2_CASE(8, 8, 8, 8) SCTW2_CASE(8, 8, 8, 4)
#undef SCTW2_CASE
#define SCTW_CASE(NSV, RV, STV, TWV)                                                                                               \
    if (tw == TWV && dstate == NSV && dt_rank_fused == RV) {                                                                       \
        if (io_dtype == FD_BF16) return scan_tw_launch<__nv_bfloat16, NSV, RV, STV, TWV>(u_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, y_nhwc, B, D, H, W, stream); \
        if (io_dtype == FD_F16) return scan_tw_launch<__half, NSV, RV, STV, TWV>(u_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, y_nhwc, B, D, H, W, stream); \
        return FD_ERR_UNSUPPORTED;                                                                                                 \
    }
        SCTW_CASE(4, 4, 16, 8) SCTW_CASE(4, 4, 16, 4) SCTW_CASE(8, 4, 8, 8) SCTW_CASE(8, 4, 8, 4) SCTW_CASE(8, 8, 8, 8) SCTW_CASE(8, 8, 8, 4)
#undef SCTW_CASE
        if (segments < 0) return FD_ERR_UNSUPPORTED;
    }
    int S = segments > 0 ? segments : pick_segments(B, D, L);
#define SCTM_CASE(NSV, RV)                                                                                                         \
    if (dstate == NSV && dt_rank_fused == RV) {                                                                                    \
        if (io_dtype == FD_BF16)                                                                                                   \
            return scan_tm_launch<__nv_bfloat16, NSV, RV>(u_tm, dts_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, carry_ws, (size_t)carry_floats, \
                                                          y_nhwc, B, D, H, W, S, stream);                                          \
        if (io_dtype == FD_F16)                                                                                                    \
            return scan_tm_launch<__half, NSV, RV>(u_tm, dts_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, carry_ws, (size_t)carry_floats, y_nhwc, \
                                                   B, D, H, W, S, stream);                                                         \
        return FD_ERR_UNSUPPORTED;                                                                                                 \
    }
    SCTM_CASE(4, 4) SCTM_CASE(8, 4) SCTM_CASE(8, 8) SCTM_CASE(16, 8) SCTM_CASE(4, 0) SCTM_CASE(8, 0) SCTM_CASE(16, 0) SCTM_CASE(32, 0)
#undef SCTM_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_scan_tm_chain_plan(int B, int D, int H, int W, int dstate, int dt_rank_fused, int* ws_floats) {
    if (ws_floats) *ws_floats = 0;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || D % 32) return 0;
    const int tw = pick_time_warps(B, D, H / 2, W / 2, dstate, dt_rank_fused);
    if (tw != 8 && tw != 4) return 0;
    if (!((dstate == 4 && dt_rank_fused == 4) || (dstate == 8 && (dt_rank_fused == 4 || dt_rank_fused == 8)))) return 0;
    const int nseg = chain_segments(H, W, tw_steps(dstate), tw);
    if (nseg && ws_floats) *ws_floats = (int)chain_ws_floats(B, D, dstate, nseg);
    return nseg;
}

extern "C" int fd_selective_scan_tm_chained(const void* u_tm, const float* xdbl_tm, const float* A, const float* dt_w, const float* dt_bias,
                                            const float* D_skip, float* chain_ws, long chain_floats, void* y_nhwc, int B, int D, int H,
                                            int W, int dstate, int dt_rank_fused, int io_dtype, cudaStream_t stream) {
    if (!u_tm || !xdbl_tm || !A || !dt_w || !dt_bias || !D_skip || !chain_ws || !y_nhwc) return FD_ERR_BAD_ARGUMENT;
    int need = 0;
    const int nseg = fd_scan_tm_chain_plan(B, D, H, W, dstate, dt_rank_fused, &need);
    if (nseg <= 1) return FD_ERR_UNSUPPORTED;
    if (chain_floats < need || ((uintptr_t)chain_ws & 15) || (((uintptr_t)u_tm | (uintptr_t)xdbl_tm) & 15)) return FD_ERR_BAD_ARGUMENT;
    const int tw = pick_time_warps(B, D, H / 2, W / 2, dstate, dt_rank_fused);
#define SCCH_CASE(NSV, RV, STV, TWV)                                                                                               \
    if (tw == TWV && dstate == NSV && dt_rank_fused == RV) {                                                                       \
        if (io_dtype == FD_BF16) return scan_tw2_launch<__nv_bfloat16, NSV, RV, STV, TWV>(u_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, y_nhwc, B, D, H, W, nseg, (uint32_t*)chain_ws, stream); \
        if (io_dtype == FD_F16) return scan_tw2_launch<__half, NSV, RV, STV, TWV>(u_tm, xdbl_tm, A, dt_w, dt_bias, D_skip, y_nhwc, B, D, H, W, nseg, (uint32_t*)chain_ws, stream); \
        return FD_ERR_UNSUPPORTED;                                                                                                 \
    }
    SCCH_CASE(4, 4, 16, 8) SCCH_CASE(4, 4, 16, 4) SCCH_CASE(8, 4, 8, 8) SCCH_CASE(8, 4, 8, 4) SCCH_CASE(8, 8, 8, 8) SCCH_CASE(8, 8, 8, 4)
#undef SCCH_CASE
    return FD_ERR_UNSUPPORTED;
}
