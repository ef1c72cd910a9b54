// CUDA-core implicit-GEMM convolution with the fused epilogue of include/founddiff_b200.h (fd_conv_params).
// This is the fp32 validation path and the fallback for shapes the tcgen05 path does not take; it is also the
// on-device cross-check for fd_conv2d_tc in tests.  fp32 accumulate for every storage dtype.
//
// GEMM view: M = pixels of ONE sample (tiles never straddle samples), N = Cout, K = KH*KW*(c0+c1), K index
// = (kh*KW + kw)*Cin + ci so that consecutive k are consecutive channels of one input pixel.
#include "fd_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TA, typename T>       // TA: src0 / src1 / weight storage, T: out / addend storage
__global__ void __launch_bounds__(256) conv_simt_kernel(fd_conv_params p, int Hout, int Wout, int tiles_per_sample) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    __shared__ float gsum[2][16];  // up to 16 groups touched by one 64-channel tile

    const int b = blockIdx.x / tiles_per_sample;
    const int m0 = (blockIdx.x % tiles_per_sample) * TM;
    const int n0 = blockIdx.y * TN;
    const int Cin = p.c0 + p.c1;
    const int ld0 = p.ld0 > 0 ? p.ld0 : p.c0;
    const int K = p.KH * p.KW * Cin;
    const int P = Hout * Wout;
    const int Hin = p.Hin, Win = p.Win;
    const int Hv = p.upsample ? 2 * Hin : Hin, Wv = p.upsample ? 2 * Win : Win;  // virtual (upsampled) input size
    const TA* src0 = (const TA*)p.src0;
    const TA* src1 = (const TA*)p.src1;
    const TA* wgt = (const TA*)p.weight + (p.per_batch_weight ? (long)b * p.Cout * K : 0);

    const int tid = threadIdx.x;
    // loader mapping: 4 consecutive k for one row
    const int lrow = tid / 4, lk = (tid % 4) * 4;
    const int am = m0 + lrow;
    const int aho = am / Wout, awo = am % Wout;
    const bool arow_ok = am < P;
    const int bn = n0 + lrow;
    const bool brow_ok = bn < p.Cout;

    const int tx = tid % 16, ty = tid / 16;  // micro-tile: rows ty*4.., cols tx*4..
    float acc[4][4] = {};

    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + lk + j;
            float av = 0.f, bv = 0.f;
            if (k < K) {
                if (arow_ok) {
                    const int tap = k / Cin, ci = k % Cin;
                    const int kh = tap / p.KW, kw = tap % p.KW;
                    int hi = aho * p.stride - p.pad + kh, wi = awo * p.stride - p.pad + kw;
                    if (hi >= 0 && hi < Hv && wi >= 0 && wi < Wv) {
                        if (p.upsample) { hi >>= 1; wi >>= 1; }
                        const long pix = ((long)b * Hin + hi) * Win + wi;
                        av = ci < p.c0 ? fd_ld(src0 + pix * ld0 + ci) : fd_ld(src1 + pix * p.c1 + (ci - p.c0));
                    }
                }
                if (brow_ok) bv = fd_ld(wgt + (long)bn * K + k);
            }
            As[lk + j][lrow] = av;
            Bs[lk + j][lrow] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; w[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue
    const int cpg = p.gn_sums ? p.Cout / p.gn_groups : 1;
    if (p.gn_sums) {
        if (tid < 32) gsum[tid / 16][tid % 16] = 0.f;
        __syncthreads();
    }
    float gs[4] = {}, gq[4] = {};
    T* out = (T*)p.out;
    const T* addend = (const T*)p.addend;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= P) continue;
        const long orow = ((long)b * P + m) * p.Cout;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
            if (n >= p.silu_from) v = fd_silu(v);
            gs[j] += v;
            gq[j] += v * v;
            float o = p.gate ? p.gate[(long)b * p.gate_stride + n] * v : v;
            if (addend) o += fd_ld(addend + orow + n);
            if (p.relu_out) o = fmaxf(o, 0.f);
            fd_st(out + orow + n, o);
        }
    }
    if (p.gn_sums) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < p.Cout) {
                const int gl = (n - n0) / cpg;  // group index local to the tile (cpg >= 4 => < 16)
                atomicAdd(&gsum[0][gl], gs[j]);
                atomicAdd(&gsum[1][gl], gq[j]);
            }
        }
        __syncthreads();
        const int ngl = (min(TN, p.Cout - n0) + cpg - 1) / cpg;
        if (tid < 2 * ngl) {
            const int which = tid / ngl, gl = tid % ngl;
            const int g = n0 / cpg + gl;
            atomicAdd(&p.gn_sums[((long)b * p.gn_groups + g) * 2 + which], gsum[which][gl]);
        }
    }
}

}  // namespace

static int conv_out_dims(const fd_conv_params* p, int* Hout, int* Wout) {
    const int Hv = p->upsample ? 2 * p->Hin : p->Hin, Wv = p->upsample ? 2 * p->Win : p->Win;
    *Hout = (Hv + 2 * p->pad - p->KH) / p->stride + 1;
    *Wout = (Wv + 2 * p->pad - p->KW) / p->stride + 1;
    return (*Hout > 0 && *Wout > 0) ? 0 : FD_ERR_BAD_ARGUMENT;
}

extern "C" int fd_conv_check_params(const fd_conv_params* p) {
    if (!p || !p->src0 || !p->weight || !p->out) return FD_ERR_BAD_ARGUMENT;
    if (p->c0 <= 0 || p->c1 < 0 || (p->c1 > 0 && !p->src1)) return FD_ERR_BAD_ARGUMENT;
    if (p->B <= 0 || p->Hin <= 0 || p->Win <= 0 || p->Cout <= 0 || p->KH <= 0 || p->KW <= 0 || p->stride <= 0 || p->pad < 0)
        return FD_ERR_BAD_ARGUMENT;
    if (p->gate && p->gate_stride < p->Cout) return FD_ERR_BAD_ARGUMENT;
    if (p->ld0 != 0 && p->ld0 < p->c0) return FD_ERR_BAD_ARGUMENT;
    if (p->gn_sums) {
        if (p->gn_groups <= 0 || p->Cout % p->gn_groups) return FD_ERR_BAD_ARGUMENT;
        const int cpg = p->Cout / p->gn_groups;
        if (cpg < 4 || cpg % 4 || (64 % cpg && cpg % 64)) return FD_ERR_UNSUPPORTED;
    }
    if (p->dtype != FD_F32 && p->dtype != FD_BF16 && p->dtype != FD_F16) return FD_ERR_BAD_ARGUMENT;
    if (p->ab_dtype_p1) {
        const int ab = p->ab_dtype_p1 - 1;
        if (ab != FD_F32 && ab != FD_BF16 && ab != FD_F16) return FD_ERR_BAD_ARGUMENT;
        if ((ab == FD_F32) != (p->dtype == FD_F32)) return FD_ERR_UNSUPPORTED;       // fp32 does not mix with 16-bit storage
    }
    return 0;
}

extern "C" int fd_conv2d_simt(const fd_conv_params* p, cudaStream_t stream) {
    int rc = fd_conv_check_params(p);
    if (rc) return rc;
    int Hout, Wout;
    if ((rc = conv_out_dims(p, &Hout, &Wout))) return rc;
    if (p->gn_sums && (p->Cout / p->gn_groups) > 64) return FD_ERR_UNSUPPORTED;
    const int tiles = fd_cdiv((long)Hout * Wout, TM);
    dim3 grid((unsigned)(tiles * p->B), fd_cdiv(p->Cout, TN));
    const int ab = p->ab_dtype_p1 ? p->ab_dtype_p1 - 1 : p->dtype;
    if (ab == p->dtype) {
        FD_DISPATCH_DTYPE(p->dtype, T, (conv_simt_kernel<T, T><<<grid, 256, 0, stream>>>(*p, Hout, Wout, tiles)));
    } else if (ab == FD_BF16 && p->dtype == FD_F16) {
        conv_simt_kernel<__nv_bfloat16, __half><<<grid, 256, 0, stream>>>(*p, Hout, Wout, tiles);
    } else if (ab == FD_F16 && p->dtype == FD_BF16) {
        conv_simt_kernel<__half, __nv_bfloat16><<<grid, 256, 0, stream>>>(*p, Hout, Wout, tiles);
    } else {
        return FD_ERR_UNSUPPORTED;
    }
    FD_LAUNCH_CHECK();
    return 0;
}
