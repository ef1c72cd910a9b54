// S6 selective scan forward (the op behind selective_scan_cuda_core.fwd, src/emamba2.py:152-154).
//
// Mapping: one WARP per (batch, channel) row; the row is consumed in chunks of 32 lanes x 8 consecutive
// time steps.  Per state n the lane composes its 8 (a, b) pairs locally, a 5-round shuffle scan composes the
// lane aggregates, and the running state h[n] (held redundantly in every lane) carries across chunks — i.e. a
// chunked warp-shuffle scan with no shared memory and no block barrier.  Loads/stores are 16-byte vectors,
// consecutive lanes touching consecutive addresses.  exp(dt*A) is one MUFU.EX2 per (step, state): the kernel
// is bound by the SFU/FMA pipes rather than by HBM for dstate >= 4 ("selective scan").
#include <type_traits>

#include "fd_common.cuh"

namespace {

constexpr int kItems = 8;
constexpr int kChunk = 32 * kItems;
constexpr int kWarpsPerBlock = 4;

FD_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T>
FD_DEVINL void load_items(const T* __restrict__ row, int l0, int L, bool vec_ok, float (&v)[kItems]) {
    if (vec_ok && l0 + kItems <= L) {
        if constexpr (sizeof(T) == 2) {
            fd_ldv<T, 8>(row + l0, v);
        } else {
            float a[4], b[4];
            fd_ldv<float, 4>(reinterpret_cast<const float*>(row) + l0, a);
            fd_ldv<float, 4>(reinterpret_cast<const float*>(row) + l0 + 4, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) v[i] = (l0 + i < L) ? fd_ld(row + l0 + i) : 0.f;
    }
}

template <typename T>
FD_DEVINL void store_items(T* __restrict__ row, int l0, int L, bool vec_ok, const float (&v)[kItems]) {
    if (vec_ok && l0 + kItems <= L) {
        if constexpr (sizeof(T) == 2) {
            fd_stv<T, 8>(row + l0, v);
        } else {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = v[i]; b[i] = v[4 + i]; }
            fd_stv<float, 4>(reinterpret_cast<float*>(row) + l0, a);
            fd_stv<float, 4>(reinterpret_cast<float*>(row) + l0 + 4, b);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i)
            if (l0 + i < L) fd_st(row + l0 + i, v[i]);
    }
}

template <typename T, int NS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) selective_scan_kernel(
    const T* __restrict__ u, const T* __restrict__ delta, const float* __restrict__ A, const float* __restrict__ Bm,
    const float* __restrict__ Cm, const float* __restrict__ D, const float* __restrict__ delta_bias, T* __restrict__ y,
    long rows, int dim, int L, int N, int G, int softplus, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / dim), d = (int)(row % dim);
    const int g = d / (dim / G);
    const T* ur = u + row * (long)L;
    const T* dr = delta + row * (long)L;
    T* yr = y + row * (long)L;
    const float* Bg = Bm + ((long)b * G + g) * N * (long)L;
    const float* Cg = Cm + ((long)b * G + g) * N * (long)L;
    const float bias = delta_bias ? delta_bias[d] : 0.f;
    const float Dd = D ? D[d] : 0.f;

    float A2[NS], h[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        A2[n] = (n < N) ? A[(long)d * N + n] * 1.4426950408889634f : 0.f;
        h[n] = 0.f;
    }

    for (int c0 = 0; c0 < L; c0 += kChunk) {
        const int l0 = c0 + lane * kItems;
        float dt[kItems], dtu[kItems], yacc[kItems];
        load_items<T>(dr, l0, L, vec_ok, dt);
        load_items<T>(ur, l0, L, vec_ok, dtu);
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            float t = dt[i] + bias;
            if (softplus) t = fd_softplus20(t);
            if (l0 + i >= L) t = 0.f;  // identity element beyond the end of the row
            yacc[i] = Dd * dtu[i];
            dtu[i] = t * dtu[i];
            dt[i] = t;
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            if (n < N) {
                float Bn[kItems], Cn[kItems];
                load_items<float>(Bg + (long)n * L, l0, L, vec_ok, Bn);
                load_items<float>(Cg + (long)n * L, l0, L, vec_ok, Cn);
                float a[kItems], ap = 1.f, bp = 0.f;
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    a[i] = ex2_approx(dt[i] * A2[n]);
                    Bn[i] = dtu[i] * Bn[i];
                    ap *= a[i];
                    bp = fmaf(a[i], bp, Bn[i]);
                }
                // inclusive scan of the lane aggregates: (a2,b2) o (a1,b1) = (a1*a2, a2*b1 + b2)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float au = __shfl_up_sync(0xffffffffu, ap, o);
                    const float bu = __shfl_up_sync(0xffffffffu, bp, o);
                    if (lane >= o) { bp = fmaf(ap, bu, bp); ap *= au; }
                }
                float ae = __shfl_up_sync(0xffffffffu, ap, 1);
                float be = __shfl_up_sync(0xffffffffu, bp, 1);
                if (lane == 0) { ae = 1.f; be = 0.f; }
                float hh = fmaf(ae, h[n], be);  // state entering this lane's first step
                const float at = __shfl_sync(0xffffffffu, ap, 31);
                const float bt = __shfl_sync(0xffffffffu, bp, 31);
                h[n] = fmaf(at, h[n], bt);
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    hh = fmaf(a[i], hh, Bn[i]);
                    yacc[i] = fmaf(hh, Cn[i], yacc[i]);
                }
            }
        }
        store_items<T>(yr, l0, L, vec_ok, yacc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// v2 (aligned rows): 8 warps = 8 channels of ONE direction group per block.  B_n[l] / C_n[l] are shared by every
// channel of the group, so the block stages each 256-step chunk of them ONCE in shared memory with cp.async (double
// buffered when it fits) instead of every warp re-reading them from L2 — that re-read was 5x (N=4) to 40x (N=32) the
// kernel's HBM traffic and bounded v1 by L2 bandwidth.  The next chunk's u / delta vectors are prefetched into
// registers before the current chunk is processed.
constexpr int kRowsPerBlock = 8;
constexpr int kPadChunk = kChunk + (kChunk >> 5) * 4;     // 4 floats of padding per 32: conflict-free 32-byte lane reads

FD_DEVINL int pad_idx(int l) { return l + ((l >> 5) << 2); }
FD_DEVINL void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
FD_DEVINL float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log1p(exp(x)), branch-free, 2 MUFU ops.  w = 1 + y rounds; for y < 1 the lost part (y - (w - 1)) is exact and
// log1p(y) = log(w) + log1p((y - (w-1)) / w) ~ log(w) + (y - (w-1)) * (1 - y)   (the factor only has to be right to
// O(1) because the term itself is <= half an ulp of w); for y >= 1 the rounding of w is below 1e-7 relative already.
FD_DEVINL float fast_softplus(float x) {
    const float y = ex2_approx(x * 1.4426950408889634f);
    const float w = 1.f + y;
    const float corr = (y < 1.f) ? (y - (w - 1.f)) * (1.f - y) : 0.f;
    const float r = fmaf(lg2_approx(w), 0.6931471805599453f, corr);
    return x > 20.f ? x : r;
}

// One Kogge-Stone round of the (a, b) composition across lanes: lanes whose source lane exists (the shuffle's own
// predicate) fold it in, with predicated FMA/MUL instead of compare + select.
FD_DEVINL void scan_round(float& ap, float& bp, int o) {
    asm("{\n"
        ".reg .pred p;\n"
        ".reg .f32 au, bu;\n"
        "shfl.sync.up.b32 au|p, %0, %2, 0, 0xffffffff;\n"
        "shfl.sync.up.b32 bu, %1, %2, 0, 0xffffffff;\n"
        "@p fma.rn.f32 %1, %0, bu, %1;\n"
        "@p mul.f32 %0, %0, au;\n"
        "}\n"
        : "+f"(ap), "+f"(bp)
        : "r"(o));
}
// exclusive prefix of the lane aggregates: (ae, be) = aggregate of lane-1, identity for lane 0
FD_DEVINL void scan_exclusive(float ap, float bp, float& ae, float& be) {
    asm("{\n"
        ".reg .pred p;\n"
        "shfl.sync.up.b32 %0|p, %2, 1, 0, 0xffffffff;\n"
        "shfl.sync.up.b32 %1, %3, 1, 0, 0xffffffff;\n"
        "@!p mov.f32 %0, 0f3F800000;\n"
        "@!p mov.f32 %1, 0f00000000;\n"
        "}\n"
        : "=&f"(ae), "=&f"(be)
        : "f"(ap), "f"(bp));
}

// raw (unconverted) 8-item vectors: the prefetch for the next chunk is held in this form so that nothing consumes the
// load before the chunk boundary (a float conversion right behind the load made every warp sit out the DRAM latency)
template <typename T> struct RawItems { uint4 v[sizeof(T) * kItems / 16]; };
template <typename T> FD_DEVINL RawItems<T> load_raw(const T* p) {
    RawItems<T> r;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) * kItems / 16); ++i) r.v[i] = __ldg(reinterpret_cast<const uint4*>(p) + i);
    return r;
}
template <typename T> FD_DEVINL void raw_to_float(const RawItems<T>& r, float (&v)[kItems]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            v[4 * i] = __uint_as_float(r.v[i].x); v[4 * i + 1] = __uint_as_float(r.v[i].y);
            v[4 * i + 2] = __uint_as_float(r.v[i].z); v[4 * i + 3] = __uint_as_float(r.v[i].w);
        }
    } else {
        const uint32_t w[4] = {r.v[0].x, r.v[0].y, r.v[0].z, r.v[0].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f;
            if constexpr (std::is_same<T, __nv_bfloat16>::value) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            else f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
}

// MERGE = true fuses EfficientMerge (src/emamba2.py:238-262) into the scan: instead of the (b, KD, L) scan layout the
// block transposes each 256-step chunk of its 8 channels through shared memory and writes 16-byte channel groups
// straight into the channels-last tensor y_nhwc (B, H, W, dim/4) at the pixel each (direction, l) stands for.
// One block barrier per chunk: the barrier that publishes chunk c's B/C also retires chunk c-1 (its staging buffer may
// be refilled, its transposed outputs may be written out while the warps already work on chunk c).
// RDT > 0 fuses dt_proj: `dtr` holds the rank-RDT dt input rows (fp32, same group layout as B / C with `gstride` floats
// between direction groups), dt_w the (dim, RDT) projection; delta is formed per step in registers and `delta` is unused.
template <typename T, int NS, int NBUF, bool MERGE, int RDT = 0, int NPAR = 1>
__global__ void __launch_bounds__(kRowsPerBlock * 32, (NPAR == 1 && NS <= 8 ? 3 : 2)) selective_scan_smem_kernel(
    const T* __restrict__ u, const T* __restrict__ delta, const float* __restrict__ A, const float* __restrict__ Bm,
    const float* __restrict__ Cm, const float* __restrict__ D, const float* __restrict__ delta_bias, T* __restrict__ y,
    int dim, int L, int G, int softplus, int H, int W, const float* __restrict__ dtr = nullptr,
    const float* __restrict__ dt_w = nullptr, long gstride = 0) {
    extern __shared__ __align__(16) float s_bc[];          // [NBUF][2*NS + RDT][kPadChunk] (+ A*log2e for NS >= 16) (+ s_y)
    constexpr bool kA2InSmem = NS >= 16;                   // keeps the register count at two blocks per SM for large d_state
    constexpr int kRows = 2 * NS + RDT;                    // staged rows per chunk: B, C, (dt input)
    float* s_a2 = s_bc + (size_t)NBUF * kRows * kPadChunk;
    T* s_y = reinterpret_cast<T*>(s_a2 + (kA2InSmem ? kRowsPerBlock * NS : 0));     // [2][8][kChunk] (MERGE only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_group = dim / G;
    const int blocks_per_group = per_group / kRowsPerBlock;
    const int bg = blockIdx.x / blocks_per_group;          // b * G + g
    const int dloc0 = (blockIdx.x % blocks_per_group) * kRowsPerBlock;
    const int b = bg / G;
    const long gs = RDT > 0 ? gstride : (long)NS * L;
    const float* Bg = Bm + (long)bg * gs;
    const float* Cg = Cm + (long)bg * gs;
    const float* Tg = RDT > 0 ? dtr + (long)bg * gs : nullptr;

    const int d = (bg % G) * per_group + dloc0 + warp;
    const long row = (long)b * dim + d;
    const T* ur = u + row * (long)L;
    const T* dr = delta + row * (long)L;
    T* yr = y + row * (long)L;
    const float bias = delta_bias ? delta_bias[d] : 0.f;
    const float Dd = D ? D[d] : 0.f;
    float A2[kA2InSmem ? 1 : NS], h[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        if constexpr (!kA2InSmem) A2[n] = A[(long)d * NS + n] * 1.4426950408889634f;
        h[n] = 0.f;
    }
    float wdt[RDT > 0 ? RDT : 1];
    if constexpr (RDT > 0) {
#pragma unroll
        for (int r = 0; r < RDT; ++r) wdt[r] = dt_w[(long)d * RDT + r];
    }
    if constexpr (kA2InSmem) {
        for (int n = lane; n < NS; n += 32) s_a2[warp * NS + n] = A[(long)d * NS + n] * 1.4426950408889634f;
        __syncwarp();
    }

    auto stage = [&](int c0, int buf) {   // cooperative cp.async of B/C[:, c0 : c0+256] (zero-filled past L)
        float* dst = s_bc + (size_t)buf * kRows * kPadChunk;
#pragma unroll
        for (int i = threadIdx.x; i < kRows * (kChunk / 4); i += kRowsPerBlock * 32) {
            const int v = i % (kChunk / 4), rn = i / (kChunk / 4);     // rn: 0..NS-1 = B rows, NS..2NS-1 = C rows, then dt rows
            const int l = c0 + v * 4;
            const float* src = (rn < NS ? Bg + (long)rn * L : rn < 2 * NS ? Cg + (long)(rn - NS) * L : Tg + (long)(rn - 2 * NS) * L) + l;
            cp_async16(dst + rn * kPadChunk + pad_idx(v * 4), l < L ? src : Bg, l < L);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // EfficientMerge scatter of one finished chunk: one pixel per thread, 8 channels = 16 bytes
    const int mk = bg % G;                                  // direction of this block's channel group
    const int mdiv = (mk & 1) ? (H >> 1) : (W >> 1);        // sub-grid extent along the direction's fast axis
    const int msh = (mdiv & (mdiv - 1)) ? -1 : (__ffs(mdiv) - 1);
    auto merge_store = [&](int c) {
        if constexpr (MERGE) {
            const int l = c * kChunk + (int)threadIdx.x;
            if (l < L) {
                const int qd = msh >= 0 ? (l >> msh) : (l / mdiv);
                const int rm = l - qd * mdiv;
                int hh, ww;
                if (mk & 1) { ww = 2 * qd + (mk >> 1); hh = 2 * rm + 1; }     // column-major sub-grids
                else        { hh = 2 * qd; ww = 2 * rm + (mk >> 1); }
                const unsigned short* src = reinterpret_cast<const unsigned short*>(s_y) + (c & 1) * kRowsPerBlock * kChunk + threadIdx.x;
                uint32_t pk[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) pk[r] = (uint32_t)src[(2 * r) * kChunk] | ((uint32_t)src[(2 * r + 1) * kChunk] << 16);
                *reinterpret_cast<uint4*>(y + (((long)b * H + hh) * W + ww) * per_group + dloc0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    };

    const int nchunks = (L + kChunk - 1) / kChunk;
    stage(0, 0);
    // rows are kItems-aligned (launcher), so a lane's 8 steps are all inside or all outside the row; outside lanes read a
    // clamped (valid) address and are masked through dt = 0
    RawItems<T> dt_raw;
    if constexpr (RDT == 0) dt_raw = load_raw<T>(dr + min(lane * kItems, L - kItems));
    RawItems<T> u_raw = load_raw<T>(ur + min(lane * kItems, L - kItems));

    for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * kChunk, l0 = c0 + lane * kItems;
        const int buf = (NBUF == 2) ? (c & 1) : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                  // chunk c's B/C visible to all; every warp is past chunk c-1
        if (NBUF == 2 && c + 1 < nchunks) stage(c0 + kChunk, buf ^ 1);
        if (c > 0) merge_store(c - 1);
        const float* sB = s_bc + (size_t)buf * kRows * kPadChunk + pad_idx(lane * kItems);
        const float* sC = sB + NS * kPadChunk;
        float dt[kItems], dtu[kItems], yacc[kItems];
        {
            float dtv[kItems], uv[kItems];
            if constexpr (RDT == 0) {
                raw_to_float<T>(dt_raw, dtv);
            } else {                                  // dt_proj: delta = dt_w[d, :] . x_dbl[:R, l]
#pragma unroll
                for (int i = 0; i < kItems; ++i) dtv[i] = 0.f;
                const float* sT = sB + 2 * NS * kPadChunk;
#pragma unroll
                for (int r = 0; r < RDT; ++r) {
                    const float4 t0 = *reinterpret_cast<const float4*>(sT + r * kPadChunk);
                    const float4 t1 = *reinterpret_cast<const float4*>(sT + r * kPadChunk + 4);
                    dtv[0] = fmaf(wdt[r], t0.x, dtv[0]); dtv[1] = fmaf(wdt[r], t0.y, dtv[1]);
                    dtv[2] = fmaf(wdt[r], t0.z, dtv[2]); dtv[3] = fmaf(wdt[r], t0.w, dtv[3]);
                    dtv[4] = fmaf(wdt[r], t1.x, dtv[4]); dtv[5] = fmaf(wdt[r], t1.y, dtv[5]);
                    dtv[6] = fmaf(wdt[r], t1.z, dtv[6]); dtv[7] = fmaf(wdt[r], t1.w, dtv[7]);
                }
            }
            raw_to_float<T>(u_raw, uv);
            const bool live = l0 < L;
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                float t = dtv[i] + bias;
                if (softplus) t = fast_softplus(t);
                t = live ? t : 0.f;
                yacc[i] = Dd * uv[i];
                dtu[i] = t * uv[i];
                dt[i] = t;
            }
        }
        if (c + 1 < nchunks) {          // prefetch the next chunk's u / delta (raw: consumed after the next barrier)
            const int ln = min(l0 + kChunk, L - kItems);
            if constexpr (RDT == 0) dt_raw = load_raw<T>(dr + ln);
            u_raw = load_raw<T>(ur + ln);
        }
        float sdt = 0.f;
#pragma unroll
        for (int i = 0; i < kItems; ++i) sdt += dt[i];
        // NPAR states at a time, their five dependent shuffle rounds issued interleaved.  SASS of the NPAR = 1 kernel shows the
        // states fully serialised (80 registers leave ptxas no room to overlap them: ~200 clk of shuffle latency per state and
        // warp); NPAR = 2 (FD_SCAN_NPAR=2) overlaps two chains but needs 128 registers -> 2 blocks per SM instead of 3, and
        // MEASURED SLOWER: 1694 vs 1621 us (d_state 4), 1394 vs 1299 us (d_state 8).  Occupancy hides that latency better.
#pragma unroll
        for (int n0 = 0; n0 < NS; n0 += NPAR) {
            float a[NPAR][kItems], Bn[NPAR][kItems], ap[NPAR], bp[NPAR];
#pragma unroll
            for (int q = 0; q < NPAR; ++q) {
                const int n = n0 + q;
                const float4 b0 = *reinterpret_cast<const float4*>(sB + n * kPadChunk);
                const float4 b1 = *reinterpret_cast<const float4*>(sB + n * kPadChunk + 4);
                const float bv[kItems] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float a2n = kA2InSmem ? s_a2[warp * NS + n] : A2[kA2InSmem ? 0 : n];
                ap[q] = ex2_approx(sdt * a2n);      // product of the lane's 8 decay factors
                bp[q] = 0.f;
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    a[q][i] = ex2_approx(dt[i] * a2n);
                    Bn[q][i] = dtu[i] * bv[i];
                    bp[q] = fmaf(a[q][i], bp[q], Bn[q][i]);
                }
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                for (int q = 0; q < NPAR; ++q) scan_round(ap[q], bp[q], o);
#pragma unroll
            for (int q = 0; q < NPAR; ++q) {
                const int n = n0 + q;
                float ae, be;
                scan_exclusive(ap[q], bp[q], ae, be);
                float hh = fmaf(ae, h[n], be);
                const float at = __shfl_sync(0xffffffffu, ap[q], 31);
                const float bt = __shfl_sync(0xffffffffu, bp[q], 31);
                h[n] = fmaf(at, h[n], bt);
                const float4 c0v = *reinterpret_cast<const float4*>(sC + n * kPadChunk);
                const float4 c1v = *reinterpret_cast<const float4*>(sC + n * kPadChunk + 4);
                const float Cn[kItems] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    hh = fmaf(a[q][i], hh, Bn[q][i]);
                    yacc[i] = fmaf(hh, Cn[i], yacc[i]);
                }
            }
        }
        if constexpr (!MERGE) {
            if (l0 < L) store_items<T>(yr, l0, L, true, yacc);
        } else {
            fd_stv<T, 8>(s_y + ((c & 1) * kRowsPerBlock + warp) * kChunk + lane * kItems, yacc);
        }
        if (NBUF == 1) {
            __syncthreads();            // single staging buffer: everyone is done with it before it is refilled
            if (c + 1 < nchunks) stage(c0 + kChunk, 0);
        }
    }
    if constexpr (MERGE) {
        __syncthreads();
        merge_store(nchunks - 1);
    }
}

template <typename T>
int scan_launch(const void* u, const void* delta, const float* A, const float* Bm, const float* Cm, const float* D,
                const float* delta_bias, void* y, int batch, int dim, int L, int N, int G, int softplus,
                cudaStream_t st, int mergeH = 0, int mergeW = 0) {
    const long rows = (long)batch * dim;
    const int vec_ok = (L % kItems == 0) && ((((uintptr_t)u | (uintptr_t)delta | (uintptr_t)y) & 31) == 0) &&
                       ((((uintptr_t)Bm | (uintptr_t)Cm) & 31) == 0);
    // v2: rows of a block share one direction group and every access is 16/32-byte aligned
    static const int npar_env = getenv("FD_SCAN_NPAR") ? atoi(getenv("FD_SCAN_NPAR")) : 1;
    const int npar = (npar_env == 2 && (N == 4 || N == 8)) ? 2 : 1;       // states scanned with interleaved shuffle rounds
    if (vec_ok && L >= kItems && (N == 4 || N == 8 || N == 16 || N == 32)) {
#define SCAN2_CASE(NSV, NB, RPWV, NPARV)                                                                            \
    if (N == NSV && npar == NPARV && (dim / G) % (kRowsPerBlock * RPWV) == 0) {                                                      \
        const unsigned grid2 = (unsigned)(rows / (kRowsPerBlock * RPWV));                                           \
        const size_t smem = ((size_t)NB * 2 * NSV * kPadChunk + (NSV >= 16 ? kRowsPerBlock * RPWV * NSV : 0)) * sizeof(float); \
        if (mergeH) {                                                                                               \
          if constexpr (sizeof(T) != 2) { return FD_ERR_UNSUPPORTED; } else {                                       \
            static bool attr_m = false;                                                                             \
            const size_t smem_m = smem + (size_t)2 * kRowsPerBlock * RPWV * kChunk * sizeof(T);                         \
            if (!attr_m) {                                                                                          \
                cudaError_t e = cudaFuncSetAttribute(selective_scan_smem_kernel<T, NSV, NB, true, 0, NPARV>,            \
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);     \
                if (e != cudaSuccess) return (int)e;                                                                \
                attr_m = true;                                                                                      \
            }                                                                                                       \
            selective_scan_smem_kernel<T, NSV, NB, true, 0, NPARV><<<grid2, kRowsPerBlock * 32, smem_m, st>>>(          \
                (const T*)u, (const T*)delta, A, Bm, Cm, D, delta_bias, (T*)y, dim, L, G, softplus, mergeH, mergeW); \
          }                                                                                                         \
        } else {                                                                                                    \
            static bool attr_set = false;                                                                           \
            if (!attr_set) {                                                                                        \
                cudaError_t e = cudaFuncSetAttribute(selective_scan_smem_kernel<T, NSV, NB, false, 0, NPARV>,           \
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
                if (e != cudaSuccess) return (int)e;                                                                \
                attr_set = true;                                                                                    \
            }                                                                                                       \
            selective_scan_smem_kernel<T, NSV, NB, false, 0, NPARV><<<grid2, kRowsPerBlock * 32, smem, st>>>(           \
                (const T*)u, (const T*)delta, A, Bm, Cm, D, delta_bias, (T*)y, dim, L, G, softplus, 0, 0);          \
        }                                                                                                           \
        FD_LAUNCH_CHECK();                                                                                          \
        return 0;                                                                                                   \
    }
        // RPW = 2 (16 channels per block, full-sector writes) was measured SLOWER (register spills, occupancy): not used
        SCAN2_CASE(4, 2, 1, 1) SCAN2_CASE(8, 2, 1, 1) SCAN2_CASE(16, 2, 1, 1) SCAN2_CASE(32, 1, 1, 1)
        SCAN2_CASE(4, 2, 1, 2) SCAN2_CASE(8, 2, 1, 2)
#undef SCAN2_CASE
    }
    if (mergeH) return FD_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
#define SCAN_CASE(NSV)                                                                                              \
    if (N <= NSV) {                                                                                                 \
        selective_scan_kernel<T, NSV><<<grid, kWarpsPerBlock * 32, 0, st>>>((const T*)u, (const T*)delta, A, Bm, Cm, D, \
                                                                            delta_bias, (T*)y, rows, dim, L, N, G,   \
                                                                            softplus, vec_ok);                       \
        FD_LAUNCH_CHECK();                                                                                          \
        return 0;                                                                                                   \
    }
    SCAN_CASE(4) SCAN_CASE(8) SCAN_CASE(16) SCAN_CASE(32) SCAN_CASE(64)
#undef SCAN_CASE
    return FD_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int fd_selective_scan_fwd(const void* u, const void* delta, const float* A, const float* Bm,
                                     const float* Cm, const float* D, const float* delta_bias, void* y, int batch,
                                     int dim, int seqlen, int dstate, int ngroups, int delta_softplus, int io_dtype,
                                     cudaStream_t stream) {
    if (!u || !delta || !A || !Bm || !Cm || !y) return FD_ERR_BAD_ARGUMENT;
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0 || ngroups <= 0 || dim % ngroups) return FD_ERR_BAD_ARGUMENT;
    FD_DISPATCH_DTYPE(io_dtype, T,
                      return scan_launch<T>(u, delta, A, Bm, Cm, D, delta_bias, y, batch, dim, seqlen, dstate, ngroups,
                                            delta_softplus, stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Channel-per-lane scan for the deep levels (many short rows: batch*dim >= ~16 K rows, d_state 16 / 32).
// A LANE owns one (batch, channel) row and walks it sequentially with its d_state states in registers — no cross-lane
// scan at all: per (step, state) the work is FMUL, MUFU.EX2, FMUL, FFMA, FFMA (5 issue slots instead of ~15 for the
// warp-shuffle formulation), which puts these launches on the MUFU pipe.  The 32 lanes of a warp are 32 consecutive
// channels of one direction group, so (a) B_n[l] / C_n[l] are warp-uniform: they are staged per 64-step chunk, in the
// time-major layout (B, 4, L, N) the x_proj kernel writes for this path, and read as broadcast LDS.128; (b) the fused
// EfficientMerge store is naturally coalesced: a warp writes 32 consecutive channels (64 B) of one pixel per step.
// u / delta are read straight from global memory, 16 bytes per lane per 8 steps, prefetched one octet ahead.
constexpr int kCLChunk = 64;

template <typename T, int NS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS) selective_scan_cl_kernel(
    const T* __restrict__ u, const T* __restrict__ delta, const float* __restrict__ A, const float* __restrict__ Bt,
    const float* __restrict__ Ct, const float* __restrict__ D, const float* __restrict__ delta_bias, T* __restrict__ y,
    int dim, int L, int softplus, int H, int W, int pf) {
    extern __shared__ __align__(16) float s_cl[];          // [2][2][kCLChunk][NS]: buffer, B/C, step, state
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_group = dim / 4;
    const int blocks_per_group = per_group / (WARPS * 32);
    const int bg = blockIdx.x / blocks_per_group;          // b * 4 + k
    const int dloc = (blockIdx.x % blocks_per_group) * (WARPS * 32) + warp * 32 + lane;
    const int b = bg >> 2, k = bg & 3;
    const int d = k * per_group + dloc;
    const long row = (long)b * dim + d;
    const T* ur = u + row * (long)L;
    const T* dr = delta + row * (long)L;
    const float* Bg = Bt + (long)bg * L * NS;
    const float* Cg = Ct + (long)bg * L * NS;
    const float bias = delta_bias ? delta_bias[d] : 0.f;
    const float Dd = D ? D[d] : 0.f;
    float A2[NS], h[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) { A2[n] = A[(long)d * NS + n] * 1.4426950408889634f; h[n] = 0.f; }

    auto stage = [&](int c0, int buf) {                    // B and C of steps [c0, c0 + 64): two contiguous runs of 64*NS floats
        float* dst = s_cl + (size_t)buf * 2 * kCLChunk * NS;
        constexpr int kVec = kCLChunk * NS / 4;            // 16-byte vectors per tensor
        for (int i = threadIdx.x; i < 2 * kVec; i += WARPS * 32) {
            const int which = i / kVec, v = i % kVec;
            const int l = c0 + (v * 4) / NS;
            const float* src = (which ? Cg : Bg) + (long)c0 * NS + v * 4;
            cp_async16(dst + which * kCLChunk * NS + v * 4, l < L ? src : Bg, l < L);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // EfficientMerge coordinates of step l, advanced incrementally (src/emamba2.py:207-210, 253-256)
    const int mdiv = (k & 1) ? (H >> 1) : (W >> 1);
    int mq = 0, mr = 0;
    T* ybase = y + (long)b * H * W * per_group + dloc;

    const int nchunks = (L + kCLChunk - 1) / kCLChunk;
    stage(0, 0);
    RawItems<T> dt_raw = load_raw<T>(dr), u_raw = load_raw<T>(ur);
    for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * kCLChunk;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // chunk c's B/C visible; every warp is past chunk c-1
        if (c + 1 < nchunks) stage(c0 + kCLChunk, (c + 1) & 1);
        const float* sB = s_cl + (size_t)(c & 1) * 2 * kCLChunk * NS;
        const float* sC = sB + kCLChunk * NS;
        const int steps = min(kCLChunk, L - c0);           // L % 8 == 0
        for (int s0 = 0; s0 < steps; s0 += kItems) {
            float dtv[kItems], uv[kItems];
            raw_to_float<T>(dt_raw, dtv);
            raw_to_float<T>(u_raw, uv);
            {
                const int ln = min(c0 + s0 + kItems, L - kItems);      // next octet (clamped on the last one)
                dt_raw = load_raw<T>(dr + ln);
                u_raw = load_raw<T>(ur + ln);
                if (pf > 0) {                                          // every lane walks its own row: pull its next lines into L2
                    const int lp = min(ln + pf, L - kItems);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(dr + lp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ur + lp));
                }
            }
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                float t = dtv[i] + bias;
                if (softplus) t = fast_softplus(t);
                const float du = t * uv[i];
                float y0 = Dd * uv[i], y1 = 0.f, y2 = 0.f, y3 = 0.f;
                const float* pb = sB + (s0 + i) * NS;
                const float* pc = sC + (s0 + i) * NS;
#pragma unroll
                for (int n = 0; n < NS; n += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(pb + n);
                    const float4 c4 = *reinterpret_cast<const float4*>(pc + n);
                    h[n] = fmaf(ex2_approx(t * A2[n]), h[n], du * b4.x);
                    y0 = fmaf(h[n], c4.x, y0);
                    h[n + 1] = fmaf(ex2_approx(t * A2[n + 1]), h[n + 1], du * b4.y);
                    y1 = fmaf(h[n + 1], c4.y, y1);
                    h[n + 2] = fmaf(ex2_approx(t * A2[n + 2]), h[n + 2], du * b4.z);
                    y2 = fmaf(h[n + 2], c4.z, y2);
                    h[n + 3] = fmaf(ex2_approx(t * A2[n + 3]), h[n + 3], du * b4.w);
                    y3 = fmaf(h[n + 3], c4.w, y3);
                }
                int hh, ww;
                if (k & 1) { ww = 2 * mq + (k >> 1); hh = 2 * mr + 1; }      // column-major sub-grids
                else       { hh = 2 * mq; ww = 2 * mr + (k >> 1); }
                fd_st(ybase + ((long)hh * W + ww) * per_group, (y0 + y1) + (y2 + y3));
                if (++mr == mdiv) { mr = 0; ++mq; }
            }
        }
    }
}

template <typename T, int NS>
static int scan_cl_launch(const void* u, const void* delta, const float* A, const float* Bt, const float* Ct, const float* D,
                          const float* delta_bias, void* y, int batch, int dim, int H, int W, int softplus, cudaStream_t st) {
    const int L = (H / 2) * (W / 2);
    const int per_group = dim / 4;
    const size_t smem = (size_t)2 * 2 * kCLChunk * NS * sizeof(float);
    // Block size (measured, B200): d_state 32 is bounded by occupancy (128 registers, 32 KB of staging per block) and runs
    // best with 4-warp blocks; for d_state <= 16 the launch is one wave of long-running blocks and an uneven block count
    // per SM is lost time, so the largest of 4 / 2 / 1 warps that still gives >= 6 blocks per SM is used.
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long total_warps = (long)batch * dim / 32;
    static const int pf = getenv("FD_SCANCL_PF") ? atoi(getenv("FD_SCANCL_PF")) : 0;     // L2 prefetch distance in steps (0 = off)
#define CL_CASE(WV)                                                                                                    \
    if (per_group % (WV * 32) == 0 && (WV == 1 || NS >= 32 || total_warps / WV >= 6L * sms)) {                         \
        static bool attr = false;                                                                                      \
        if (!attr) {                                                                                                   \
            cudaError_t e = cudaFuncSetAttribute(selective_scan_cl_kernel<T, NS, WV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                       \
            attr = true;                                                                                               \
        }                                                                                                              \
        const unsigned grid = (unsigned)(batch * 4 * (per_group / (WV * 32)));                                         \
        selective_scan_cl_kernel<T, NS, WV><<<grid, WV * 32, smem, st>>>((const T*)u, (const T*)delta, A, Bt, Ct, D, delta_bias, \
                                                                         (T*)y, dim, L, softplus, H, W, pf);           \
        FD_LAUNCH_CHECK();                                                                                             \
        return 0;                                                                                                      \
    }
    // 4-warp blocks: more blocks than SMs already at 16 K rows, and B/C staging is shared by 128 channels
    CL_CASE(4) CL_CASE(2) CL_CASE(1)
#undef CL_CASE
    return FD_ERR_UNSUPPORTED;
}

extern "C" int fd_selective_scan_fwd_merge_cl(const void* u, const void* delta, const float* A, const float* Bt, const float* Ct,
                                              const float* D, const float* delta_bias, void* y_nhwc, int batch, int dim, int H,
                                              int W, int dstate, int delta_softplus, int io_dtype, cudaStream_t stream) {
    if (!u || !delta || !A || !Bt || !Ct || !y_nhwc) return FD_ERR_BAD_ARGUMENT;
    if (batch <= 0 || dim <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || dim % 4) return FD_ERR_BAD_ARGUMENT;
    const int L = (H / 2) * (W / 2);
    if (L % kItems || (dim / 4) % 32 || (((uintptr_t)u | (uintptr_t)delta | (uintptr_t)Bt | (uintptr_t)Ct) & 15)) return FD_ERR_UNSUPPORTED;
#define CL_N(NSV)                                                                                                       \
    if (dstate == NSV) {                                                                                               \
        if (io_dtype == FD_BF16) return scan_cl_launch<__nv_bfloat16, NSV>(u, delta, A, Bt, Ct, D, delta_bias, y_nhwc, batch, dim, H, W, delta_softplus, stream); \
        if (io_dtype == FD_F16) return scan_cl_launch<__half, NSV>(u, delta, A, Bt, Ct, D, delta_bias, y_nhwc, batch, dim, H, W, delta_softplus, stream); \
    }
    CL_N(8) CL_N(16) CL_N(32)
#undef CL_N
    return FD_ERR_UNSUPPORTED;
}

// Scan + EfficientMerge + fused dt_proj (see include/founddiff_b200.h)
template <typename T, int NS, int RDT>
static int scan_xdbl_launch(const void* u, const float* x_dbl, const float* dt_w, const float* A, const float* D,
                            const float* delta_bias, void* y, int batch, int dim, int H, int W, int softplus, cudaStream_t st) {
    const int L = (H / 2) * (W / 2);
    constexpr int kRows = 2 * NS + RDT;
    const size_t smem = (size_t)2 * kRows * kPadChunk * sizeof(float) + (size_t)2 * kRowsPerBlock * kChunk * sizeof(T);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(selective_scan_smem_kernel<T, NS, 2, true, RDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const long rows = (long)batch * dim;
    selective_scan_smem_kernel<T, NS, 2, true, RDT><<<(unsigned)(rows / kRowsPerBlock), kRowsPerBlock * 32, smem, st>>>(
        (const T*)u, nullptr, A, x_dbl + (long)RDT * L, x_dbl + (long)(RDT + NS) * L, D, delta_bias, (T*)y, dim, L, 4, softplus, H, W,
        x_dbl, dt_w, (long)kRows * L);
    FD_LAUNCH_CHECK();
    return 0;
}

extern "C" int fd_selective_scan_fwd_merge_xdbl(const void* u, const float* x_dbl, const float* dt_w, const float* A, const float* D,
                                                const float* delta_bias, void* y_nhwc, int batch, int dim, int H, int W, int dstate,
                                                int dt_rank, int delta_softplus, int io_dtype, cudaStream_t stream) {
    if (!u || !x_dbl || !dt_w || !A || !y_nhwc) return FD_ERR_BAD_ARGUMENT;
    if (batch <= 0 || dim <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || dim % 4) return FD_ERR_BAD_ARGUMENT;
    const int L = (H / 2) * (W / 2);
    if (L % kItems || L < kItems || (dim / 4) % kRowsPerBlock || (((uintptr_t)u | (uintptr_t)y_nhwc | (uintptr_t)x_dbl) & 15))
        return FD_ERR_UNSUPPORTED;
#define XDBL_CASE(NSV, RV)                                                                                              \
    if (dstate == NSV && dt_rank == RV) {                                                                               \
        if (io_dtype == FD_BF16)                                                                                        \
            return scan_xdbl_launch<__nv_bfloat16, NSV, RV>(u, x_dbl, dt_w, A, D, delta_bias, y_nhwc, batch, dim, H, W, delta_softplus, stream); \
        if (io_dtype == FD_F16)                                                                                         \
            return scan_xdbl_launch<__half, NSV, RV>(u, x_dbl, dt_w, A, D, delta_bias, y_nhwc, batch, dim, H, W, delta_softplus, stream); \
    }
    XDBL_CASE(4, 4) XDBL_CASE(8, 8)
#undef XDBL_CASE
    return FD_ERR_UNSUPPORTED;
}

// Scan + EfficientMerge: same inputs as fd_selective_scan_fwd (4 direction groups, L = H/2 * W/2), output written
// channels-last: y_nhwc (batch, H, W, dim/4), pixel of (direction k, step l) per src/emamba2.py:207-210, 253-256.
extern "C" int fd_selective_scan_fwd_merge(const void* u, const void* delta, const float* A, const float* Bm, const float* Cm,
                                           const float* D, const float* delta_bias, void* y_nhwc, int batch, int dim, int H,
                                           int W, int dstate, int delta_softplus, int io_dtype, cudaStream_t stream) {
    if (!u || !delta || !A || !Bm || !Cm || !y_nhwc) return FD_ERR_BAD_ARGUMENT;
    if (batch <= 0 || dim <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || dstate <= 0 || dim % 4) return FD_ERR_BAD_ARGUMENT;
    const int L = (H / 2) * (W / 2);
    if (io_dtype == FD_BF16)
        return scan_launch<__nv_bfloat16>(u, delta, A, Bm, Cm, D, delta_bias, y_nhwc, batch, dim, L, dstate, 4, delta_softplus, stream, H, W);
    if (io_dtype == FD_F16)
        return scan_launch<__half>(u, delta, A, Bm, Cm, D, delta_bias, y_nhwc, batch, dim, L, dstate, 4, delta_softplus, stream, H, W);
    return FD_ERR_UNSUPPORTED;
}
