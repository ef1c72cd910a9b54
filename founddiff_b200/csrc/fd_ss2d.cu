// SS2D producer / consumer kernels around the selective scan (src/emamba2.py:186-262, 295-367, 713-751).
//
//   fd_dwconv3x3_silu_scan : NHWC x-half of xz -> depthwise 3x3 + bias + SiLU -> 4-direction scan layout
//   fd_xdt_proj            : x_proj / dt_proj on the scan layout
//   fd_merge_ln_gate       : scan layout -> NHWC, LayerNorm(D), * z + local
//
// Scan layout: xs[b, k, d, l], L = (H/2)*(W/2); pixel (h, w) belongs to k = (h&1) | ((w&1)<<1);
//   k in {0,2} (even h): l = (h/2)*(W/2) + (w/2)   (row-major sub-grid)
//   k in {1,3} (odd  h): l = (w/2)*(H/2) + (h/2)   (column-major sub-grid)            emamba2.py:207-210
// Both transposing kernels stage a 32x32-pixel x 16-channel tile in shared memory so that global reads are
// 16-byte vectors along channels and global writes are 16-byte vectors along l (and vice versa).
#include <stdlib.h>
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int TS = 32;        // spatial tile edge (pixels)
constexpr int TH2 = TS / 2;   // sub-grid tile edge
constexpr int CH = 16;        // channels per tile

FD_DEVINL int scan_class(int py, int px) { return (py & 1) | ((px & 1) << 1); }
// index inside the 16x16 sub-grid tile, in the direction's own scan order
FD_DEVINL int scan_local(int k, int py, int px) { return (k & 1) ? (px >> 1) * TH2 + (py >> 1) : (py >> 1) * TH2 + (px >> 1); }

// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dwconv_scan_kernel(const T* __restrict__ xz, int ld, const float* __restrict__ w,
                                                          const float* __restrict__ bias, T* __restrict__ xs, int H, int W,
                                                          int D, int vec_ok) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int NVC = CH / VEC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_in = reinterpret_cast<T*>(smem_raw);                      // [(TS+2)*(TS+2)][CH]
    T* s_out = s_in + (TS + 2) * (TS + 2) * CH;                    // [CH][4][TH2*TH2]
    float* s_w = reinterpret_cast<float*>(s_out + CH * 4 * TH2 * TH2);  // [9][CH] + [CH]

    const int c0 = blockIdx.x * CH;
    const int tiles_w = (W + TS - 1) / TS;
    const int ty0 = (blockIdx.y / tiles_w) * TS, tx0 = (blockIdx.y % tiles_w) * TS;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;

    for (int i = tid; i < 10 * CH; i += 256) {
        const int c = i % CH, tap = i / CH;
        s_w[i] = tap < 9 ? w[(long)(c0 + c) * 9 + tap] : (bias ? bias[c0 + c] : 0.f);
    }
    for (int i = tid; i < (TS + 2) * (TS + 2) * NVC; i += 256) {
        const int pix = i / NVC, vc = i % NVC;
        const int h = ty0 + pix / (TS + 2) - 1, ww = tx0 + pix % (TS + 2) - 1;
        float v[VEC];
        if (h >= 0 && h < H && ww >= 0 && ww < W) {
            fd_ldv<T, VEC>(xz + (((long)b * H + h) * W + ww) * ld + c0 + vc * VEC, v);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = 0.f;
        }
        fd_stv<T, VEC>(s_in + pix * CH + vc * VEC, v);
    }
    __syncthreads();
    for (int i = tid; i < TS * TS * NVC; i += 256) {
        const int pix = i / NVC, vc = i % NVC;
        const int py = pix / TS, px = pix % TS;
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = s_w[9 * CH + vc * VEC + e];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                float v[VEC];
                fd_ldv<T, VEC>(s_in + ((py + dy) * (TS + 2) + px + dx) * CH + vc * VEC, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] = fmaf(v[e], s_w[(dy * 3 + dx) * CH + vc * VEC + e], acc[e]);
            }
        const int k = scan_class(py, px);
        const int idx = scan_local(k, py, px);
#pragma unroll
        for (int e = 0; e < VEC; ++e) fd_st(s_out + ((vc * VEC + e) * 4 + k) * (TH2 * TH2) + idx, fd_silu(acc[e]));
    }
    __syncthreads();
    // write runs of TH2 consecutive l
    const int H2 = H / 2, W2 = W / 2, L = H2 * W2;
    const int h2_0 = ty0 / 2, w2_0 = tx0 / 2;
    constexpr int VPR = TH2 / VEC;  // vectors per run
    for (int i = tid; i < CH * 4 * TH2 * VPR; i += 256) {
        const int vv = i % VPR, r = (i / VPR) % TH2, k = (i / (VPR * TH2)) % 4, c = i / (VPR * TH2 * 4);
        long l;
        int nvalid;
        bool run_ok;
        if (k & 1) { run_ok = (w2_0 + r) < W2; l = (long)(w2_0 + r) * H2 + h2_0; nvalid = H2 - h2_0; }
        else       { run_ok = (h2_0 + r) < H2; l = (long)(h2_0 + r) * W2 + w2_0; nvalid = W2 - w2_0; }
        if (!run_ok) continue;
        const T* sp = s_out + (c * 4 + k) * (TH2 * TH2) + r * TH2 + vv * VEC;
        T* gp = xs + (((long)b * 4 + k) * D + c0 + c) * L + l + vv * VEC;
        if (vec_ok && (vv + 1) * VEC <= nvalid) {
            *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                if (vv * VEC + e < nvalid) gp[e] = sp[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 16-bit storage types: register sliding-window formulation of the same op.  A block owns a 32x32-pixel x 32-channel
// tile; a thread owns a horizontal pixel PAIR x 4 channels and walks down the rows with a 3-row prefetch ring of raw
// 8-byte vectors (loads are unconditional on clamped addresses and masked when consumed, so they stay in flight),
// weights in registers.  Finished SiLU outputs go to a shared-memory tile already in scan order
// [channel][k][run][16 l] (run stride 9 words, channel stride 577 words: bank-conflict-free for both the row-major
// and the column-major directions); the tile is then written as 32-byte runs along l.
constexpr int RW_TS = 32, RW_CH = 32, RW_V = 4;
constexpr int RW_RUN = 18;                      // elements per padded run (16 + 2)
constexpr int RW_KST = 16 * RW_RUN;             // elements per direction k
constexpr int RW_CST = 4 * RW_KST + 2;          // elements per channel (577 words)

template <typename T> FD_DEVINL uint2 rw_ld_raw(const T* p) { return *reinterpret_cast<const uint2*>(p); }
template <typename T> FD_DEVINL void rw_cvt(uint2 r, float (&v)[4]) {
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

// write phase of the register-window kernels: 8 lanes x 4 bytes per 16-element run; each k covers the 16 runs of one (channel, k)
template <typename T>
FD_DEVINL void rw_write_phase(const T* s_out, T* __restrict__ xs, int b, int c0, int ty0, int tx0, int H, int W, int D,
                              int word_ok, int tid) {
    const int H2 = H / 2, W2 = W / 2;
    const long L = (long)H2 * W2;
    const int h2_0 = ty0 / 2, w2_0 = tx0 / 2;
    const int j = tid % 8, r = tid / 8;
    const uint32_t* s_words = reinterpret_cast<const uint32_t*>(s_out) + r * (RW_RUN / 2) + j;
    T* gbase = xs + ((long)b * 4 * D + c0) * L + 2 * j;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        long l;
        int nvalid;
        bool run_ok;
        if (k & 1) { run_ok = (w2_0 + r) < W2; l = (long)(w2_0 + r) * H2 + h2_0; nvalid = H2 - h2_0; }
        else       { run_ok = (h2_0 + r) < H2; l = (long)(h2_0 + r) * W2 + w2_0; nvalid = W2 - w2_0; }
        if (!run_ok) continue;
        T* gp = gbase + (long)k * D * L + l;
        const uint32_t* sp = s_words + k * (RW_KST / 2);
        if (word_ok && 2 * j + 1 < nvalid) {
#pragma unroll 8
            for (int c = 0; c < RW_CH; ++c) *reinterpret_cast<uint32_t*>(gp + c * L) = sp[c * (RW_CST / 2)];
        } else {
            for (int c = 0; c < RW_CH; ++c) {
                const uint32_t wv = sp[c * (RW_CST / 2)];
                if (2 * j < nvalid) *reinterpret_cast<unsigned short*>(gp + c * L) = (unsigned short)(wv & 0xffffu);
                if (2 * j + 1 < nvalid) *reinterpret_cast<unsigned short*>(gp + c * L + 1) = (unsigned short)(wv >> 16);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128, 3) dwconv_scan_rw_kernel(const T* __restrict__ xz, int ld, const float* __restrict__ w,
                                                                const float* __restrict__ bias, T* __restrict__ xs, int H,
                                                                int W, int D, int word_ok) {
    static_assert(sizeof(T) == 2, "16-bit storage only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_out = reinterpret_cast<T*>(smem_raw);                              // [RW_CH][RW_CST]
    float* s_w = reinterpret_cast<float*>(s_out + RW_CH * RW_CST);          // [RW_CH*9] + [RW_CH]

    const int c0 = blockIdx.x * RW_CH;
    const int tiles_w = (W + RW_TS - 1) / RW_TS;
    const int ty0 = (blockIdx.y / tiles_w) * RW_TS, tx0 = (blockIdx.y % tiles_w) * RW_TS;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cv = tid % (RW_CH / RW_V), pr = tid / (RW_CH / RW_V);         // channel vector, pixel pair
    const int x = tx0 + 2 * pr;

    for (int i = tid; i < RW_CH * 10; i += 128)
        s_w[i] = i < RW_CH * 9 ? w[(long)c0 * 9 + i] : (bias ? bias[c0 + i - RW_CH * 9] : 0.f);
    __syncthreads();
    float wr[9][RW_V], bs[RW_V];
#pragma unroll
    for (int e = 0; e < RW_V; ++e) {
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[t][e] = s_w[(cv * RW_V + e) * 9 + t];
        bs[e] = s_w[RW_CH * 9 + cv * RW_V + e];
    }

    const int y0 = ty0, y1 = min(H, ty0 + RW_TS);
    const bool has_0 = x < W, has_l = has_0 && x > 0, has_1 = x + 1 < W, has_2 = x + 2 < W;
    const int xc = min(x, W - 1);
    const T* base = xz + (long)b * H * W * ld + c0 + cv * RW_V;
    const long offl = has_l ? (long)ld : 0, off1 = has_1 ? (long)ld : 0, off2 = has_2 ? 2L * ld : 0;
    const int ymax = min(H - 1, y1);
    float acc[3][2][RW_V];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < RW_V; ++e) acc[r][px][e] = bs[e];
    uint2 ring[3][4];
    auto fetch = [&](int yr, auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        const T* rp = base + ((long)min(max(yr, 0), ymax) * W + xc) * ld;
        ring[S][0] = rw_ld_raw<T>(rp - offl);
        ring[S][1] = rw_ld_raw<T>(rp);
        ring[S][2] = rw_ld_raw<T>(rp + off1);
        ring[S][3] = rw_ld_raw<T>(rp + off2);
    };
    T* s_thr = s_out + cv * RW_V * RW_CST;
    auto step = [&](int yi, auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        float v[4][RW_V];
        {
            const bool rv = yi >= 0 && yi < H;
            const bool ok[4] = {rv && has_l, rv && has_0, rv && has_1, rv && has_2};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 r = ring[S][j];
                r.x = ok[j] ? r.x : 0u;
                r.y = ok[j] ? r.y : 0u;
                rw_cvt<T>(r, v[j]);
            }
        }
        fetch(yi + 3, slot_c);
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int e = 0; e < RW_V; ++e) {
                    const float xv = v[px + dx][e];
                    acc[(S + 1) % 3][px][e] = fmaf(xv, wr[0 * 3 + dx][e], acc[(S + 1) % 3][px][e]);
                    acc[S][px][e] = fmaf(xv, wr[1 * 3 + dx][e], acc[S][px][e]);
                    acc[(S + 2) % 3][px][e] = fmaf(xv, wr[2 * 3 + dx][e], acc[(S + 2) % 3][px][e]);
                }
        const int yo = yi - 1;
        if (yo >= y0 && yo < y1) {
            const int py = yo - y0;
            // even rows: run = py/2, position = pair; odd rows (column-major directions): run = pair, position = py/2
            const int off = (py & 1) ? (RW_KST + pr * RW_RUN + (py >> 1)) : ((py >> 1) * RW_RUN + pr);
#pragma unroll
            for (int px = 0; px < 2; ++px)
#pragma unroll
                for (int e = 0; e < RW_V; ++e)
                    fd_st(s_thr + e * RW_CST + px * 2 * RW_KST + off, fd_silu16(acc[(S + 2) % 3][px][e]));
        }
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < RW_V; ++e) acc[(S + 2) % 3][px][e] = bs[e];
    };
    int yi = y0 - 1;
    yi -= ((yi % 3) + 3) % 3;
    fetch(yi, std::integral_constant<int, 0>{});
    fetch(yi + 1, std::integral_constant<int, 1>{});
    fetch(yi + 2, std::integral_constant<int, 2>{});
    for (; yi <= y1; yi += 3) {
        step(yi, std::integral_constant<int, 0>{});
        if (yi + 1 <= y1) step(yi + 1, std::integral_constant<int, 1>{});
        if (yi + 2 <= y1) step(yi + 2, std::integral_constant<int, 2>{});
    }
    __syncthreads();

    rw_write_phase<T>(s_out, xs, b, c0, ty0, tx0, H, W, D, word_ok, tid);
}

// v2 of dwconv_scan_rw_kernel (same changes as dwconv3x3_nhwc_v2_kernel, fd_attn.cu): one running row pointer advanced by a
// block-uniform stride, mask-free variant for warps without an edge column, a finished accumulator slot re-seeded by the
// first FMA of the next output row.  The step stays straight-line (branches around the consumers made ptxas serialise the
// prefetch ring: measured 30 % slower).
template <typename T, bool EDGE>
FD_DEVINL void rw_conv_phase(const T* __restrict__ base, T* s_thr, const float (&wr)[9][RW_V], const float (&bs)[RW_V], int H,
                             int W, int ld, int xc, int y0, int y1, int pr, bool has_0, bool has_l, bool has_1, bool has_2,
                             int pf) {
    const int ymax = min(H - 1, y1);
    int yi = y0 - 1;
    yi -= ((yi % 3) + 3) % 3;
    const long rowe = (long)W * ld;
    const int offl = (!EDGE || has_l) ? ld : 0, off1 = (!EDGE || has_1) ? ld : 0, off2 = (!EDGE || has_2) ? 2 * ld : 0;
    const T* p = base + (long)min(max(yi, 0), ymax) * rowe + (long)xc * ld;
    int yf = yi;
    float acc[3][2][RW_V];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < RW_V; ++e) acc[r][px][e] = 0.f;
    uint2 ring[3][4];
    auto fetch = [&](auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        ring[S][0] = rw_ld_raw<T>(p - offl);
        ring[S][1] = rw_ld_raw<T>(p);
        ring[S][2] = rw_ld_raw<T>(p + off1);
        ring[S][3] = rw_ld_raw<T>(p + off2);
        if (pf > 0 && yf >= 0 && yf + pf <= ymax) {      // block-uniform: pull the row `pf` steps further down into L2
            const T* pp = p + (long)pf * rowe;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + off1));
        }
        p += (yf >= 0 && yf < ymax) ? rowe : 0;          // block-uniform
        ++yf;
    };
    auto step = [&](int yrow, auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        constexpr int S1 = (S + 1) % 3, S2 = (S + 2) % 3;
        float v[4][RW_V];
        {
            const bool rv = yrow >= 0 && yrow < H;
            const bool ok[4] = {rv && (!EDGE || has_l), rv && (!EDGE || has_0), rv && (!EDGE || has_1), rv && (!EDGE || has_2)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 r = ring[S][j];
                r.x = ok[j] ? r.x : 0u;
                r.y = ok[j] ? r.y : 0u;
                rw_cvt<T>(r, v[j]);
            }
        }
        fetch(slot_c);
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int e = 0; e < RW_V; ++e) {
                float a1 = fmaf(v[px][e], wr[0][e], bs[e]);                       // output row yrow+1: first contribution
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
        if (yo >= y0 && yo < y1) {
            const int py = yo - y0;
            // even rows: run = py/2, position = pair; odd rows (column-major directions): run = pair, position = py/2
            const int off = (py & 1) ? (RW_KST + pr * RW_RUN + (py >> 1)) : ((py >> 1) * RW_RUN + pr);
#pragma unroll
            for (int px = 0; px < 2; ++px)
#pragma unroll
                for (int e = 0; e < RW_V; ++e)
                    fd_st(s_thr + e * RW_CST + px * 2 * RW_KST + off, fd_silu16(acc[S2][px][e]));
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
__global__ void __launch_bounds__(128, 3) dwconv_scan_rw2_kernel(const T* __restrict__ xz, int ld, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, T* __restrict__ xs, int H,
                                                                 int W, int D, int word_ok, int pf) {
    static_assert(sizeof(T) == 2, "16-bit storage only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_out = reinterpret_cast<T*>(smem_raw);                              // [RW_CH][RW_CST]
    float* s_w = reinterpret_cast<float*>(s_out + RW_CH * RW_CST);          // [RW_CH*9] + [RW_CH]
    const int c0 = blockIdx.x * RW_CH;
    const int tiles_w = (W + RW_TS - 1) / RW_TS;
    const int ty0 = (blockIdx.y / tiles_w) * RW_TS, tx0 = (blockIdx.y % tiles_w) * RW_TS;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cv = tid % (RW_CH / RW_V), pr = tid / (RW_CH / RW_V);         // channel vector, pixel pair
    const int x = tx0 + 2 * pr;
    for (int i = tid; i < RW_CH * 10; i += 128)
        s_w[i] = i < RW_CH * 9 ? w[(long)c0 * 9 + i] : (bias ? bias[c0 + i - RW_CH * 9] : 0.f);
    __syncthreads();
    float wr[9][RW_V], bs[RW_V];
#pragma unroll
    for (int e = 0; e < RW_V; ++e) {
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[t][e] = s_w[(cv * RW_V + e) * 9 + t];
        bs[e] = s_w[RW_CH * 9 + cv * RW_V + e];
    }
    const int y0 = ty0, y1 = min(H, ty0 + RW_TS);
    const bool has_0 = x < W, has_l = has_0 && x > 0, has_1 = x + 1 < W, has_2 = x + 2 < W;
    const int xc = min(x, W - 1);
    const T* base = xz + (long)b * H * W * ld + c0 + cv * RW_V;
    T* s_thr = s_out + cv * RW_V * RW_CST;
    const bool edge = !(has_0 && has_l && has_1 && has_2);
    if (__any_sync(0xffffffffu, edge)) rw_conv_phase<T, true>(base, s_thr, wr, bs, H, W, ld, xc, y0, y1, pr, has_0, has_l, has_1, has_2, pf);
    else rw_conv_phase<T, false>(base, s_thr, wr, bs, H, W, ld, xc, y0, y1, pr, true, true, true, true, pf);
    __syncthreads();
    rw_write_phase<T>(s_out, xs, b, c0, ty0, tx0, H, W, D, word_ok, tid);
}

// ---------------------------------------------------------------------------------------------------------
// x_proj + dt_proj.  One thread per l (coalesced along l), CCP >= R+2N accumulators in registers.
template <typename T, int CCP>
__global__ void __launch_bounds__(128) xdt_proj_kernel(const T* __restrict__ xs, const float* __restrict__ x_proj_w,
                                                       const float* __restrict__ dt_w, T* __restrict__ dts,
                                                       float* __restrict__ Bs, float* __restrict__ Cs, int D, int L, int R,
                                                       int N) {
    constexpr int DT = 32;  // d-chunk of x_proj_w staged in shared memory
    __shared__ float s_w[CCP][DT + 1];
    const int CC = R + 2 * N;
    const int bk = blockIdx.y;  // b*4 + k
    const int k = bk & 3;
    const int l = blockIdx.x * 128 + threadIdx.x;
    const bool ok = l < L;
    const T* xr = xs + (long)bk * D * L;
    const float* wx = x_proj_w + (long)k * CC * D;
    float acc[CCP];
#pragma unroll
    for (int c = 0; c < CCP; ++c) acc[c] = 0.f;
    for (int d0 = 0; d0 < D; d0 += DT) {
        __syncthreads();
        for (int i = threadIdx.x; i < CCP * DT; i += 128) {
            const int c = i / DT, dd = i % DT;
            s_w[c][dd] = (c < CC && d0 + dd < D) ? wx[(long)c * D + d0 + dd] : 0.f;
        }
        __syncthreads();
        const int dn = min(DT, D - d0);
        for (int dd = 0; dd < dn; ++dd) {
            const float x = ok ? fd_ld(xr + (long)(d0 + dd) * L + l) : 0.f;
#pragma unroll
            for (int c = 0; c < CCP; ++c) acc[c] = fmaf(s_w[c][dd], x, acc[c]);
        }
    }
    if (!ok) return;
    float* Bo = Bs + (long)bk * N * L;
    float* Co = Cs + (long)bk * N * L;
#pragma unroll
    for (int c = 0; c < CCP; ++c) {
        if (c >= R && c < R + N) Bo[(long)(c - R) * L + l] = acc[c];
        if (c >= R + N && c < CC) Co[(long)(c - R - N) * L + l] = acc[c];
    }
    const float* wd = dt_w + (long)k * D * R;
    T* dr = dts + (long)bk * D * L;
    for (int d = 0; d < D; ++d) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < CCP; ++r)
            if (r < R) s = fmaf(__ldg(wd + (long)d * R + r), acc[r], s);
        fd_st(dr + (long)d * L + l, s);
    }
}

// ---------------------------------------------------------------------------------------------------------
// x_proj + dt_proj on the tensor cores (16-bit storage types).  Per (sample, direction) and 128 consecutive l:
//   stage 1:  X_dbl[c, l] = sum_d Wx[c, d] xs[d, l]        M = c (<= 96, padded to 16), N = l (128), K = d
//   stage 2:  dts[d, l]   = sum_r Wdt[d, r] X_dbl[r, l]     M = d, N = l, K = r (padded to 16 / 32)
// mma.sync m16n8k16 with fp32 accumulation; xs chunks are staged in shared memory (ldmatrix.trans, since the scan
// layout is l-contiguous), weights arrive pre-converted / zero-padded from the host (xw16: (4, CCp, D), dw16: (4, D, Rp)).
// The kernel is bound by the xs read + dts write (2 x 2PC elements); the matmuls are ~2 % of the tensor peak.
constexpr int XT_L = 128;          // l tile
constexpr int XT_KC = 32;          // d chunk of stage 1
constexpr int XT_XLD = XT_L + 8;   // padded row of the xs / X_dbl / output tiles (conflict-free ldmatrix)
constexpr int XT_WLD = XT_KC + 8;

template <typename T, int MT>      // MT = CCp / 16 (1..6)
__global__ void __launch_bounds__(256) xdt_proj_mma_kernel(const T* __restrict__ xs, const T* __restrict__ xw16,
                                                           const T* __restrict__ dw16, T* __restrict__ dts,
                                                           float* __restrict__ Bs, float* __restrict__ Cs, int D, int L, int R,
                                                           int N, int Rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_x = reinterpret_cast<T*>(smem_raw);                 // [XT_KC][XT_XLD]
    T* s_w = s_x + XT_KC * XT_XLD;                           // [MT*16][XT_WLD]
    T* s_xd = s_w + MT * 16 * XT_WLD;                        // [32][XT_XLD]   first Rp rows of X_dbl
    T* s_o = s_xd + 32 * XT_XLD;                             // [8 warps][16][XT_XLD] output staging
    const int CC = R + 2 * N, CCp = MT * 16;
    const int bk = blockIdx.y, k = bk & 3;
    const int l0 = blockIdx.x * XT_L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const T* xr = xs + (long)bk * D * L;
    const T* wx = xw16 + (long)k * CCp * D;
    const bool full = (l0 + XT_L <= L) && (L % 8 == 0);

    float acc[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

    for (int d0 = 0; d0 < D; d0 += XT_KC) {
        __syncthreads();
        // xs chunk: 32 rows x 128 l  (16 vectors of 8 per row)
        for (int i = tid; i < XT_KC * (XT_L / 8); i += 256) {
            const int r = i / (XT_L / 8), v = i % (XT_L / 8);
            uint4 val = make_uint4(0, 0, 0, 0);
            if (d0 + r < D) {
                const T* src = xr + (long)(d0 + r) * L + l0 + v * 8;
                if (full) val = *reinterpret_cast<const uint4*>(src);
                else {
                    T tmp[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) tmp[e] = (l0 + v * 8 + e < L) ? src[e] : T(0.f);
                    val = *reinterpret_cast<uint4*>(tmp);
                }
            }
            *reinterpret_cast<uint4*>(s_x + r * XT_XLD + v * 8) = val;
        }
        // Wx chunk: CCp rows x 32 d (4 vectors per row)
        for (int i = tid; i < CCp * (XT_KC / 8); i += 256) {
            const int r = i / (XT_KC / 8), v = i % (XT_KC / 8);
            uint4 val = make_uint4(0, 0, 0, 0);
            if (d0 + v * 8 < D) val = *reinterpret_cast<const uint4*>(wx + (long)r * D + d0 + v * 8);
            *reinterpret_cast<uint4*>(s_w + r * XT_WLD + v * 8) = val;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < XT_KC / 16; ++ks) {
            uint32_t bfr[4];
            ldmatrix_x4_trans(bfr, s_x + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * XT_XLD + warp * 16 + 8 * (lane >> 4));
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint32_t afr[4];
                ldmatrix_x4(afr, s_w + (mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * XT_WLD + ks * 16 + 8 * (lane >> 4));
                mma_16816<T>(acc[mt][0], afr, bfr[0], bfr[1]);
                mma_16816<T>(acc[mt][1], afr, bfr[2], bfr[3]);
            }
        }
    }
    // stage-1 epilogue: Bs / Cs (fp32, global) and the dt rows (16-bit, shared)
    for (int i = tid; i < 32 * XT_XLD / 8; i += 256) *reinterpret_cast<uint4*>(s_xd + i * 8) = make_uint4(0, 0, 0, 0);
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int c = mt * 16 + g + 8 * hh;
                const int col = warp * 16 + nt * 8 + 2 * t4;
                const float v0 = acc[mt][nt][2 * hh], v1 = acc[mt][nt][2 * hh + 1];
                if (c < R) {
                    fd_st(s_xd + c * XT_XLD + col, v0);
                    fd_st(s_xd + c * XT_XLD + col + 1, v1);
                } else if (c < CC) {
                    float* dst = (c < R + N ? Bs + ((long)bk * N + (c - R)) * L : Cs + ((long)bk * N + (c - R - N)) * L) + l0 + col;
                    // odd L: rows of B / C start at odd float offsets, a float2 store would be misaligned (found with the 48x80
                    // fixture: level 3 has L = 15 -> cudaErrorMisalignedAddress)
                    if (l0 + col + 1 < L && !(L & 1)) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                    else {
                        if (l0 + col < L) dst[0] = v0;
                        if (l0 + col + 1 < L) dst[1] = v1;
                    }
                }
            }
    __syncthreads();
    // stage 2: each warp takes m-tiles (16 channels d) round-robin, all 128 l
    const T* wd = dw16 + (long)k * D * Rp;
    T* so = s_o + warp * 16 * XT_XLD;
    T* dr = dts + (long)bk * D * L;
    for (int mtile = warp; mtile * 16 < D; mtile += 8) {
        const int dbase = mtile * 16;
        float o[16][4];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
        for (int ks = 0; ks < Rp / 16; ++ks) {
            uint32_t afr[4];
            const T* w0 = wd + (long)(dbase + g) * Rp + ks * 16 + 2 * t4;
            const T* w1 = wd + (long)(dbase + g + 8) * Rp + ks * 16 + 2 * t4;
            afr[0] = *reinterpret_cast<const uint32_t*>(w0);
            afr[1] = *reinterpret_cast<const uint32_t*>(w1);
            afr[2] = *reinterpret_cast<const uint32_t*>(w0 + 8);
            afr[3] = *reinterpret_cast<const uint32_t*>(w1 + 8);
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                uint32_t bfr[4];
                ldmatrix_x4_trans(bfr, s_xd + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * XT_XLD + np * 16 + 8 * (lane >> 4));
                mma_16816<T>(o[2 * np], afr, bfr[0], bfr[1]);
                mma_16816<T>(o[2 * np + 1], afr, bfr[2], bfr[3]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const int col = nt * 8 + 2 * t4;
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                *reinterpret_cast<__nv_bfloat162*>(so + g * XT_XLD + col) = __floats2bfloat162_rn(o[nt][0], o[nt][1]);
                *reinterpret_cast<__nv_bfloat162*>(so + (g + 8) * XT_XLD + col) = __floats2bfloat162_rn(o[nt][2], o[nt][3]);
            } else {
                *reinterpret_cast<__half2*>(so + g * XT_XLD + col) = fd_floats2half2_sat(o[nt][0], o[nt][1]);
                *reinterpret_cast<__half2*>(so + (g + 8) * XT_XLD + col) = fd_floats2half2_sat(o[nt][2], o[nt][3]);
            }
        }
        __syncwarp();
        // 16 rows x 256 B: lanes write 16-byte chunks, two rows per instruction
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 2 + (lane >> 4), v = lane & 15;
            if (dbase + r < D) {
                T* dst = dr + (long)(dbase + r) * L + l0 + v * 8;
                if (full) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(so + r * XT_XLD + v * 8);
                else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (l0 + v * 8 + e < L) dst[e] = so[r * XT_XLD + v * 8 + e];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// x_proj only (dt_rank <= 8 levels): X_dbl[b, k, c, l] = sum_d Wx[k, c, d] * xs[b, k, d, l], all R + 2N rows written in
// fp32 to ONE tensor (B, 4, R+2N, L).  The scan kernel reads its B / C rows from it and applies the rank-R dt_proj
// itself (R FMAs per step), so the (B, 4, D, L) delta tensor is never written or read: the op becomes a single
// streaming pass over xs.  128-step tiles, 8 warps x 16 columns, 32-channel chunks through a 3-stage cp.async ring.
constexpr int XP_STAGES = 3;
constexpr int XO_LD = 64 + 8;      // padded row of the per-warp dt staging tile (half an l tile): 3 blocks per SM up to MT = 5
FD_DEVINL float xp_softplus(float x) {      // same branch-free 2-MUFU form as the scan kernel's (fd_scan.cu)
    float y, lg;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    const float w = 1.f + y;
    const float corr = (y < 1.f) ? (y - (w - 1.f)) * (1.f - y) : 0.f;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(w));
    const float r = fmaf(lg, 0.6931471805599453f, corr);
    return x > 20.f ? x : r;
}
FD_DEVINL void xp_cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// WITH_DT = true is the pipelined form of xdt_proj_mma_kernel: B / C rows go to Bs / Cs, the R dt rows stay in shared
// memory (16-bit) and the second GEMM (dt_proj, K = Rp) produces the (B, 4, D, L) delta tensor.
template <typename T, int MT, bool WITH_DT>
__global__ void __launch_bounds__(256, 3) x_proj_mma_kernel(const T* __restrict__ xs, const T* __restrict__ xw16,
                                                         float* __restrict__ xdbl, int D, int L, int CC,
                                                         const T* __restrict__ dw16 = nullptr, T* __restrict__ dts = nullptr,
                                                         float* __restrict__ Bs = nullptr, float* __restrict__ Cs = nullptr,
                                                         int R = 0, int N = 0, int Rp = 0, int bc_time_major = 0,
                                                         const float* __restrict__ dt_bias = nullptr, int dt_softplus = 0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_x = reinterpret_cast<T*>(smem_raw);                     // [XP_STAGES][XT_KC][XT_XLD]
    T* s_w = s_x + XP_STAGES * XT_KC * XT_XLD;                   // [XP_STAGES][MT*16][XT_WLD]
    T* s_xd = s_w + XP_STAGES * MT * 16 * XT_WLD;                // WITH_DT: [32][XT_XLD] first Rp rows of X_dbl
    T* s_o = s_xd + 32 * XT_XLD;                                 // WITH_DT: [8 warps][16][XO_LD] output staging (one 64-l half)
    constexpr int CCp = MT * 16;
    const int bk = blockIdx.y, k = bk & 3;
    const int l0 = blockIdx.x * XT_L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const T* xr = xs + (long)bk * D * L;
    const T* wx = xw16 + (long)k * CCp * D;
    const int nchunks = D / XT_KC;

    auto stage = [&](int ch, int buf) {
        const int d0 = ch * XT_KC;
        T* sx = s_x + buf * XT_KC * XT_XLD;
        T* sw = s_w + buf * CCp * XT_WLD;
        for (int i = tid; i < XT_KC * (XT_L / 8); i += 256) {
            const int r = i / (XT_L / 8), v = i % (XT_L / 8);
            const bool ok = l0 + v * 8 < L;                       // L % 8 == 0: vectors are all-in or all-out
            xp_cp_async16(sx + r * XT_XLD + v * 8, ok ? xr + (long)(d0 + r) * L + l0 + v * 8 : xr, ok);
        }
        for (int i = tid; i < CCp * (XT_KC / 8); i += 256) {
            const int r = i / (XT_KC / 8), v = i % (XT_KC / 8);
            xp_cp_async16(sw + r * XT_WLD + v * 8, wx + (long)r * D + d0 + v * 8, true);
        }
    };

    float acc[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

#pragma unroll
    for (int s = 0; s < XP_STAGES - 1; ++s) {
        if (s < nchunks) stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        asm volatile("cp.async.wait_group %0;" ::"n"(XP_STAGES - 2) : "memory");
        __syncthreads();                                          // chunk ch landed; everyone is done with chunk ch-1's buffer
        if (ch + XP_STAGES - 1 < nchunks) stage(ch + XP_STAGES - 1, (ch + XP_STAGES - 1) % XP_STAGES);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const T* sx = s_x + (ch % XP_STAGES) * XT_KC * XT_XLD;
        const T* sw = s_w + (ch % XP_STAGES) * CCp * XT_WLD;
#pragma unroll
        for (int ks = 0; ks < XT_KC / 16; ++ks) {
            uint32_t bfr[4];
            ldmatrix_x4_trans(bfr, sx + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * XT_XLD + warp * 16 + 8 * (lane >> 4));
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint32_t afr[4];
                ldmatrix_x4(afr, sw + (mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * XT_WLD + ks * 16 + 8 * (lane >> 4));
                mma_16816<T>(acc[mt][0], afr, bfr[0], bfr[1]);
                mma_16816<T>(acc[mt][1], afr, bfr[2], bfr[3]);
            }
        }
    }
    if constexpr (!WITH_DT) {
        float* orow = xdbl + (long)bk * CC * L + l0;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int c = mt * 16 + g + 8 * hh;
                    const int col = warp * 16 + nt * 8 + 2 * t4;
                    if (c < CC && l0 + col < L)
                        *reinterpret_cast<float2*>(orow + (long)c * L + col) = make_float2(acc[mt][nt][2 * hh], acc[mt][nt][2 * hh + 1]);
                }
    } else {
        // stage-1 epilogue: Bs / Cs (fp32, global) and the dt rows (16-bit, shared)
        for (int i = tid; i < 32 * XT_XLD / 8; i += 256) *reinterpret_cast<uint4*>(s_xd + i * 8) = make_uint4(0, 0, 0, 0);
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int c = mt * 16 + g + 8 * hh;
                    const int col = warp * 16 + nt * 8 + 2 * t4;
                    const float v0 = acc[mt][nt][2 * hh], v1 = acc[mt][nt][2 * hh + 1];
                    if (c < R) {
                        fd_st(s_xd + c * XT_XLD + col, v0);
                        fd_st(s_xd + c * XT_XLD + col + 1, v1);
                    } else if (c < CC && l0 + col < L) {
                        if (bc_time_major) {       // (B, 4, L, N): the channel-per-lane scan reads all N states of a step at once
                            const int n = c < R + N ? c - R : c - R - N;
                            float* dst = (c < R + N ? Bs : Cs) + ((long)bk * L + l0 + col) * N + n;
                            dst[0] = v0;
                            dst[N] = v1;
                        } else {
                            float* dst = (c < R + N ? Bs + ((long)bk * N + (c - R)) * L : Cs + ((long)bk * N + (c - R - N)) * L) + l0 + col;
                            *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                        }
                    }
                }
        __syncthreads();
        // stage 2: each warp takes m-tiles (16 channels d) round-robin, all 128 l
        const T* wd = dw16 + (long)k * D * Rp;
        T* so = s_o + warp * 16 * XO_LD;
        T* dr = dts + (long)bk * D * L;
        for (int mtile = warp; mtile * 16 < D; mtile += 8) {
            const int dbase = mtile * 16;
            float b0 = 0.f, b1 = 0.f;
            if (dt_bias) {          // delta = softplus(dt_proj(...) + dt_bias) finished here: the scan (issue-bound) gets final values
                b0 = dbase + g < D ? __ldg(dt_bias + (long)k * D + dbase + g) : 0.f;
                b1 = dbase + g + 8 < D ? __ldg(dt_bias + (long)k * D + dbase + g + 8) : 0.f;
            }
            __syncwarp();
            // two halves of 64 l each: 32 accumulators live instead of 64 keeps the kernel at 3 blocks per SM (ncu: 128
            // registers -> 2 blocks, 24 % warps active, issue 40 %, top stalls wait / long_scoreboard / barrier)
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float o[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
                for (int ks = 0; ks < Rp / 16; ++ks) {
                    uint32_t afr[4];
                    const T* w0 = wd + (long)(dbase + g) * Rp + ks * 16 + 2 * t4;
                    const T* w1 = wd + (long)(dbase + g + 8) * Rp + ks * 16 + 2 * t4;
                    afr[0] = *reinterpret_cast<const uint32_t*>(w0);
                    afr[1] = *reinterpret_cast<const uint32_t*>(w1);
                    afr[2] = *reinterpret_cast<const uint32_t*>(w0 + 8);
                    afr[3] = *reinterpret_cast<const uint32_t*>(w1 + 8);
#pragma unroll
                    for (int np = 0; np < 4; ++np) {
                        uint32_t bfr[4];
                        ldmatrix_x4_trans(bfr, s_xd + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * XT_XLD + (half * 4 + np) * 16 + 8 * (lane >> 4));
                        mma_16816<T>(o[2 * np], afr, bfr[0], bfr[1]);
                        mma_16816<T>(o[2 * np + 1], afr, bfr[2], bfr[3]);
                    }
                }
                if (dt_bias) {
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        o[nt][0] += b0; o[nt][1] += b0; o[nt][2] += b1; o[nt][3] += b1;
                        if (dt_softplus) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) o[nt][e] = xp_softplus(o[nt][e]);
                        }
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int col = nt * 8 + 2 * t4;
                    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                        *reinterpret_cast<__nv_bfloat162*>(so + g * XO_LD + col) = __floats2bfloat162_rn(o[nt][0], o[nt][1]);
                        *reinterpret_cast<__nv_bfloat162*>(so + (g + 8) * XO_LD + col) = __floats2bfloat162_rn(o[nt][2], o[nt][3]);
                    } else {
                        *reinterpret_cast<__half2*>(so + g * XO_LD + col) = fd_floats2half2_sat(o[nt][0], o[nt][1]);
                        *reinterpret_cast<__half2*>(so + (g + 8) * XO_LD + col) = fd_floats2half2_sat(o[nt][2], o[nt][3]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {   // 16 rows x 128 B: lanes write 16-byte chunks, four rows per instruction
                    const int r = it * 4 + (lane >> 3), v = lane & 7, lc = l0 + half * 64 + v * 8;
                    if (dbase + r < D && lc < L)
                        *reinterpret_cast<uint4*>(dr + (long)(dbase + r) * L + lc) = *reinterpret_cast<const uint4*>(so + r * XO_LD + v * 8);
                }
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// merge, pass 1: per-pixel LayerNorm statistics straight from the scan layout (thread per l, loop over d).
template <typename T>
__global__ void __launch_bounds__(256) merge_stats_kernel(const T* __restrict__ ys, float* __restrict__ stats, int H, int W,
                                                          int D, float eps) {
    const int H2 = H / 2, W2 = W / 2, L = H2 * W2;
    const int bk = blockIdx.y, b = bk >> 2, k = bk & 3;
    const int l = blockIdx.x * 256 + threadIdx.x;
    if (l >= L) return;
    const T* yr = ys + (long)bk * D * L + l;
    float s = 0.f, q = 0.f;
    // two-pass-free: shifted sums around the first element keep the cancellation benign
    const float x0 = fd_ld(yr);
    for (int d = 0; d < D; ++d) {
        const float v = fd_ld(yr + (long)d * L) - x0;
        s += v;
        q = fmaf(v, v, q);
    }
    const float m = s / (float)D;
    const float var = fmaxf(q / (float)D - m * m, 0.f);
    int h, w;
    if (k & 1) { w = 2 * (l / H2) + (k >> 1); h = 2 * (l % H2) + 1; }
    else       { h = 2 * (l / W2); w = 2 * (l % W2) + (k >> 1); }
    float* o = stats + (((long)b * H + h) * W + w) * 2;
    o[0] = m + x0;
    o[1] = rsqrtf(var + eps);
}

// merge, pass 2: tile transpose + normalise + gate.
template <typename T>
__global__ void __launch_bounds__(256) merge_apply_kernel(const T* __restrict__ ys, const T* __restrict__ xz, int ld, int z_off,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ local, const float* __restrict__ stats,
                                                          T* __restrict__ out, int H, int W, int D, int vec_ok) {
    constexpr int VEC = fd_vec<T>::N;
    constexpr int NVC = CH / VEC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_y = reinterpret_cast<T*>(smem_raw);  // [CH][4][TH2*TH2]
    const int c0 = blockIdx.x * CH;
    const int tiles_w = (W + TS - 1) / TS;
    const int ty0 = (blockIdx.y / tiles_w) * TS, tx0 = (blockIdx.y % tiles_w) * TS;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int H2 = H / 2, W2 = W / 2, L = H2 * W2;
    const int h2_0 = ty0 / 2, w2_0 = tx0 / 2;
    constexpr int VPR = TH2 / VEC;
    for (int i = tid; i < CH * 4 * TH2 * VPR; i += 256) {
        const int vv = i % VPR, r = (i / VPR) % TH2, k = (i / (VPR * TH2)) % 4, c = i / (VPR * TH2 * 4);
        long l;
        int nvalid;
        bool run_ok;
        if (k & 1) { run_ok = (w2_0 + r) < W2; l = (long)(w2_0 + r) * H2 + h2_0; nvalid = H2 - h2_0; }
        else       { run_ok = (h2_0 + r) < H2; l = (long)(h2_0 + r) * W2 + w2_0; nvalid = W2 - w2_0; }
        T* sp = s_y + (c * 4 + k) * (TH2 * TH2) + r * TH2 + vv * VEC;
        const T* gp = ys + (((long)b * 4 + k) * D + c0 + c) * L + l + vv * VEC;
        if (run_ok && vec_ok && (vv + 1) * VEC <= nvalid) {
            *reinterpret_cast<uint4*>(sp) = *reinterpret_cast<const uint4*>(gp);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) fd_st(sp + e, (run_ok && vv * VEC + e < nvalid) ? fd_ld(gp + e) : 0.f);
        }
    }
    __syncthreads();
    for (int i = tid; i < TS * TS * NVC; i += 256) {
        const int pix = i / NVC, vc = i % NVC;
        const int py = pix / TS, px = pix % TS;
        const int h = ty0 + py, w = tx0 + px;
        if (h >= H || w >= W) continue;
        const int k = scan_class(py, px);
        const int idx = scan_local(k, py, px);
        const long gpix = ((long)b * H + h) * W + w;
        const float mean = stats[gpix * 2], rstd = stats[gpix * 2 + 1];
        float z[VEC], o[VEC];
        fd_ldv<T, VEC>(xz + gpix * ld + z_off + c0 + vc * VEC, z);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int c = vc * VEC + e;
            const float y = fd_ld(s_y + (c * 4 + k) * (TH2 * TH2) + idx);
            const float n = (y - mean) * rstd * __ldg(gamma + c0 + c) + __ldg(beta + c0 + c);
            o[e] = n * z[e] + __ldg(local + (long)b * D + c0 + c);
        }
        fd_stv<T, VEC>(out + gpix * D + c0 + vc * VEC, o);
    }
}

}  // namespace

extern "C" int fd_dwconv3x3_silu_scan(const void* xz, int ld, const float* w, const float* bias, void* xs, int B, int H,
                                      int W, int D, int dtype, cudaStream_t stream) {
    if (!xz || !w || !xs || B <= 0 || H <= 0 || W <= 0 || D <= 0 || ld < D) return FD_ERR_BAD_ARGUMENT;
    if ((H & 1) || (W & 1) || D % CH) return FD_ERR_UNSUPPORTED;
    if (dtype != FD_F32 && D % RW_CH == 0 && ld % RW_V == 0 && ((uintptr_t)xz % 8) == 0 && ((uintptr_t)xs % 4) == 0) {
        dim3 grid(D / RW_CH, fd_cdiv(H, RW_TS) * fd_cdiv(W, RW_TS), B);
        const int word_ok = ((H / 2) % 2 == 0) && ((W / 2) % 2 == 0);
        const size_t smem = (size_t)RW_CH * RW_CST * 2 + RW_CH * 10 * sizeof(float);
        static const bool use_v1 = getenv("FD_DWCONV_V1") != nullptr;      // A/B switch for the measurement scripts
        static const int pf = getenv("FD_DWCONV_PF") ? atoi(getenv("FD_DWCONV_PF")) : 4;     // L2 prefetch distance in rows (0 = off)
#define FD_RW_LAUNCH(KERNEL, TT, ...)                                                                                 \
    {                                                                                                                 \
        static bool attr_set = false;                                                                                 \
        if (!attr_set) {                                                                                              \
            cudaError_t e = cudaFuncSetAttribute(KERNEL<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                      \
            attr_set = true;                                                                                          \
        }                                                                                                             \
        KERNEL<TT><<<grid, 128, smem, stream>>>((const TT*)xz, ld, w, bias, (TT*)xs, H, W, D, word_ok __VA_ARGS__);   \
    }
        if (dtype == FD_BF16) {
            if (use_v1) FD_RW_LAUNCH(dwconv_scan_rw_kernel, __nv_bfloat16) else FD_RW_LAUNCH(dwconv_scan_rw2_kernel, __nv_bfloat16, , pf)
        } else {
            if (use_v1) FD_RW_LAUNCH(dwconv_scan_rw_kernel, __half) else FD_RW_LAUNCH(dwconv_scan_rw2_kernel, __half, , pf)
        }
#undef FD_RW_LAUNCH
        FD_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid(D / CH, fd_cdiv(H, TS) * fd_cdiv(W, TS), B);
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        if (ld % VEC) return FD_ERR_UNSUPPORTED;
        const int vec_ok = ((H / 2) % VEC == 0) && ((W / 2) % VEC == 0);
        const size_t smem = ((size_t)(TS + 2) * (TS + 2) * CH + (size_t)CH * 4 * TH2 * TH2) * sizeof(T) + 10 * CH * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(dwconv_scan_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dwconv_scan_kernel<T><<<grid, 256, smem, stream>>>((const T*)xz, ld, w, bias, (T*)xs, H, W, D, vec_ok);
    });
    FD_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int xdt_launch(const void* xs, const float* x_proj_w, const float* dt_w, void* dts, float* Bs, float* Cs, int B, int D,
                      int L, int R, int N, cudaStream_t stream) {
    const int CC = R + 2 * N;
    dim3 grid(fd_cdiv(L, 128), B * 4);
#define XDT_CASE(P)                                                                                               \
    if (CC <= P) {                                                                                                \
        xdt_proj_kernel<T, P><<<grid, 128, 0, stream>>>((const T*)xs, x_proj_w, dt_w, (T*)dts, Bs, Cs, D, L, R, N); \
        FD_LAUNCH_CHECK();                                                                                        \
        return 0;                                                                                                 \
    }
    XDT_CASE(16) XDT_CASE(32) XDT_CASE(48) XDT_CASE(96)
#undef XDT_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_xdt_proj(const void* xs, const float* x_proj_w, const float* dt_w, void* dts, float* Bs, float* Cs, int B,
                           int D, int L, int R, int N, int dtype, cudaStream_t stream) {
    if (!xs || !x_proj_w || !dt_w || !dts || !Bs || !Cs || B <= 0 || D <= 0 || L <= 0 || R <= 0 || N <= 0) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(dtype, T, return xdt_launch<T>(xs, x_proj_w, dt_w, dts, Bs, Cs, B, D, L, R, N, stream));
    return 0;
}

template <typename T>
static int xdt_mma_launch(const void* xs, const void* xw16, const void* dw16, void* dts, float* Bs, float* Cs, int B, int D, int L,
                          int R, int N, int Rp, int bc_layout, const float* dt_bias, int dt_softplus, cudaStream_t stream) {
    const int CC = R + 2 * N;
    const int MT = (CC + 15) / 16;
    dim3 grid(fd_cdiv(L, XT_L), B * 4);
    if (D % XT_KC == 0 && L % 8 == 0 && !getenv("FD_XDT_NO_PIPE")) {       // pipelined stage 1 (cp.async ring)
#define XDT_PIPE_CASE(M)                                                                                                       \
    if (MT == M) {                                                                                                             \
        const size_t smem = ((size_t)XP_STAGES * ((size_t)XT_KC * XT_XLD + (size_t)M * 16 * XT_WLD) + 32 * XT_XLD + 8 * 16 * XO_LD) * sizeof(T); \
        static bool attr_set = false;                                                                                          \
        if (!attr_set) {                                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(x_proj_mma_kernel<T, M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                               \
            attr_set = true;                                                                                                   \
        }                                                                                                                      \
        x_proj_mma_kernel<T, M, true><<<grid, 256, smem, stream>>>((const T*)xs, (const T*)xw16, nullptr, D, L, CC, (const T*)dw16, \
                                                                   (T*)dts, Bs, Cs, R, N, Rp, bc_layout, dt_bias, dt_softplus); \
        FD_LAUNCH_CHECK();                                                                                                     \
        return 0;                                                                                                              \
    }
        XDT_PIPE_CASE(1) XDT_PIPE_CASE(2) XDT_PIPE_CASE(3) XDT_PIPE_CASE(4) XDT_PIPE_CASE(5) XDT_PIPE_CASE(6)
#undef XDT_PIPE_CASE
    }
    if (bc_layout || dt_bias) return FD_ERR_UNSUPPORTED;      // time-major B / C and the fused bias + softplus exist in the pipelined kernel only
#define XDT_MMA_CASE(M)                                                                                                        \
    if (MT == M) {                                                                                                             \
        const size_t smem = ((size_t)XT_KC * XT_XLD + (size_t)M * 16 * XT_WLD + 32 * XT_XLD + 8 * 16 * XT_XLD) * sizeof(T);     \
        static bool attr_set = false;                                                                                          \
        if (!attr_set) {                                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(xdt_proj_mma_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                               \
            attr_set = true;                                                                                                   \
        }                                                                                                                      \
        xdt_proj_mma_kernel<T, M><<<grid, 256, smem, stream>>>((const T*)xs, (const T*)xw16, (const T*)dw16, (T*)dts, Bs, Cs, D, L, \
                                                                R, N, Rp);                                                      \
        FD_LAUNCH_CHECK();                                                                                                     \
        return 0;                                                                                                              \
    }
    XDT_MMA_CASE(1) XDT_MMA_CASE(2) XDT_MMA_CASE(3) XDT_MMA_CASE(4) XDT_MMA_CASE(5) XDT_MMA_CASE(6)
#undef XDT_MMA_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_xdt_proj_tc(const void* xs, const void* xw16, const void* dw16, void* dts, float* Bs, float* Cs, int B, int D,
                              int L, int R, int N, int Rp, int bc_layout, const float* dt_bias, int delta_softplus, int dtype,
                              cudaStream_t stream) {
    if (!xs || !xw16 || !dw16 || !dts || !Bs || !Cs || B <= 0 || D <= 0 || L <= 0 || R <= 0 || N <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((Rp != 16 && Rp != 32) || R > Rp || D % 16 || R + 2 * N > 96) return FD_ERR_UNSUPPORTED;
    if (dtype == FD_BF16) return xdt_mma_launch<__nv_bfloat16>(xs, xw16, dw16, dts, Bs, Cs, B, D, L, R, N, Rp, bc_layout, dt_bias, delta_softplus, stream);
    if (dtype == FD_F16) return xdt_mma_launch<__half>(xs, xw16, dw16, dts, Bs, Cs, B, D, L, R, N, Rp, bc_layout, dt_bias, delta_softplus, stream);
    return FD_ERR_UNSUPPORTED;
}

template <typename T>
static int x_proj_launch(const void* xs, const void* xw16, float* xdbl, int B, int D, int L, int CC, cudaStream_t stream) {
    const int MT = (CC + 15) / 16;
    dim3 grid(fd_cdiv(L, XT_L), B * 4);
#define XP_CASE(M)                                                                                                            \
    if (MT == M) {                                                                                                            \
        const size_t smem = (size_t)XP_STAGES * ((size_t)XT_KC * XT_XLD + (size_t)M * 16 * XT_WLD) * sizeof(T);               \
        static bool attr_set = false;                                                                                         \
        if (!attr_set) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(x_proj_mma_kernel<T, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                              \
            attr_set = true;                                                                                                  \
        }                                                                                                                     \
        x_proj_mma_kernel<T, M, false><<<grid, 256, smem, stream>>>((const T*)xs, (const T*)xw16, xdbl, D, L, CC);                    \
        FD_LAUNCH_CHECK();                                                                                                    \
        return 0;                                                                                                             \
    }
    XP_CASE(1) XP_CASE(2) XP_CASE(3) XP_CASE(4) XP_CASE(5) XP_CASE(6)
#undef XP_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_x_proj_tc(const void* xs, const void* xw16, float* x_dbl, int B, int D, int L, int R, int N, int dtype,
                            cudaStream_t stream) {
    if (!xs || !xw16 || !x_dbl || B <= 0 || D <= 0 || L <= 0 || R <= 0 || N <= 0) return FD_ERR_BAD_ARGUMENT;
    if (D % XT_KC || L % 8 || R + 2 * N > 96 || (((uintptr_t)xs | (uintptr_t)xw16) & 15) || ((uintptr_t)x_dbl & 7)) return FD_ERR_UNSUPPORTED;
    if (dtype == FD_BF16) return x_proj_launch<__nv_bfloat16>(xs, xw16, x_dbl, B, D, L, R + 2 * N, stream);
    if (dtype == FD_F16) return x_proj_launch<__half>(xs, xw16, x_dbl, B, D, L, R + 2 * N, stream);
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_merge_ln_gate(const void* ys, const void* xz, int ld, int z_off, const float* gamma, const float* beta,
                                const float* local, float* stats_ws, void* out, int B, int H, int W, int D, float eps,
                                int dtype, cudaStream_t stream) {
    if (!ys || !xz || !gamma || !beta || !local || !stats_ws || !out || B <= 0 || H <= 0 || W <= 0 || D <= 0) return FD_ERR_BAD_ARGUMENT;
    if ((H & 1) || (W & 1) || D % CH || ld < z_off + D) return FD_ERR_UNSUPPORTED;
    const int L = (H / 2) * (W / 2);
    FD_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = fd_vec<T>::N;
        if (ld % VEC || z_off % VEC) return FD_ERR_UNSUPPORTED;
        const int vec_ok = ((H / 2) % VEC == 0) && ((W / 2) % VEC == 0);
        merge_stats_kernel<T><<<dim3(fd_cdiv(L, 256), B * 4), 256, 0, stream>>>((const T*)ys, stats_ws, H, W, D, eps);
        FD_LAUNCH_CHECK();
        const size_t smem = (size_t)CH * 4 * TH2 * TH2 * sizeof(T);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(merge_apply_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dim3 grid(D / CH, fd_cdiv(H, TS) * fd_cdiv(W, TS), B);
        merge_apply_kernel<T><<<grid, 256, smem, stream>>>((const T*)ys, (const T*)xz, ld, z_off, gamma, beta, local,
                                                           stats_ws, (T*)out, H, W, D, vec_ok);
    });
    FD_LAUNCH_CHECK();
    return 0;
}
