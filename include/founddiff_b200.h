/*
 * founddiff_b200 — C ABI of the B200-native FoundDiff reverse-diffusion sampling path.
 *
 * Every entry point:
 *   - takes raw DEVICE pointers, sizes and a cudaStream_t; the caller (PyTorch host code) owns all memory;
 *   - never allocates, never synchronises, never reads device memory from the host -> stream-ordered and
 *     CUDA-graph capturable (the tcgen05 GEMM additionally needs fd_gemm_plan objects built on the host
 *     beforehand, they hold TMA descriptors only);
 *   - returns 0 on success, a cudaError_t value (>0) for a CUDA failure, or FD_ERR_* (<0) for a bad argument;
 *     nothing throws across the ABI.
 *
 * Activation layout everywhere: channels-last ("NHWC"), i.e. a (B, H, W, C) tensor is B*H*W rows of C contiguous
 * elements.  `dtype` selects the activation storage type (fd_dtype); accumulation is always fp32.
 *
 * Each declaration cites the reference interface (file:line under the reference repo) that it replaces.
 */
#ifndef FOUNDDIFF_B200_H
#define FOUNDDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(__CUDA_RUNTIME_H__) && !defined(__CUDA_RUNTIME_API_H__) && !defined(__DRIVER_TYPES_H__)
typedef struct CUstream_st* cudaStream_t;      /* include the CUDA runtime headers first if the host uses them */
#endif

typedef enum { FD_F32 = 0, FD_BF16 = 1, FD_F16 = 2 } fd_dtype;

#define FD_ERR_BAD_ARGUMENT (-1)
#define FD_ERR_UNSUPPORTED (-2)
#define FD_ERR_DRIVER (-3)

/* Library / build identification ("founddiff_b200 <version> sm_100a"). */
const char* fd_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Selective scan forward — replaces the third-party native op the reference calls:
 *   selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows)  src/emamba2.py:154
 *   selective_scan_cuda.fwd(u, delta, A, B, C, D, None, delta_bias, delta_softplus)        src/emamba2.py:152
 * u, delta, y: (batch, dim, seqlen) in `io_dtype`; A: (dim, dstate) fp32; Bm, Cm: (batch, ngroups, dstate,
 * seqlen) fp32; D, delta_bias: (dim) fp32 or NULL.  Channel d uses group d / (dim / ngroups).
 *   dt = delta + delta_bias; if delta_softplus: dt = dt <= 20 ? log1p(exp(dt)) : dt
 *   h_n <- exp(dt*A[d,n]) * h_n + dt * B[n,l] * u[l];   y[l] = sum_n h_n * C[n,l] + D[d] * u[l]
 * --------------------------------------------------------------------------------------------------------- */
int fd_selective_scan_fwd(const void* u, const void* delta, const float* A, const float* Bm, const float* Cm,
                          const float* D, const float* delta_bias, void* y, int batch, int dim, int seqlen,
                          int dstate, int ngroups, int delta_softplus, int io_dtype, cudaStream_t stream);

/* Scan fused with EfficientMerge (src/emamba2.py:238-262): same inputs with ngroups = 4 and seqlen = (H/2)*(W/2); the
 * output is written channels-last, y_nhwc (batch, H, W, dim/4), at the pixel each (direction, step) stands for
 * (16-bit io types only).  Followed by fd_ln_gate it replaces fd_merge_ln_gate. */
int fd_selective_scan_fwd_merge(const void* u, const void* delta, const float* A, const float* Bm, const float* Cm,
                                const float* D, const float* delta_bias, void* y_nhwc, int batch, int dim, int H, int W,
                                int dstate, int delta_softplus, int io_dtype, cudaStream_t stream);

/* Same as fd_selective_scan_fwd_merge with dt_proj fused (src/emamba2.py:337-338, 353-359): delta[b, d, l] =
 * sum_r dt_w[d, r] * x_dbl[b, k(d), r, l] is formed inside the scan (fp32) instead of being read from a (B, 4D, L) tensor;
 * B / C are rows [R, R+N) / [R+N, R+2N) of the same x_dbl.  dt_w: (dim, R) fp32 = dt_projs_weight flattened over (k, d).
 * dt_rank in {4, 8}; 16-bit io only. */
int fd_selective_scan_fwd_merge_xdbl(const void* u, const float* x_dbl, const float* dt_w, const float* A, const float* D,
                                     const float* delta_bias, void* y_nhwc, int batch, int dim, int H, int W, int dstate,
                                     int dt_rank, int delta_softplus, int io_dtype, cudaStream_t stream);

/* Scan + EfficientMerge for the deep levels (many short rows, d_state 8 / 16 / 32): one LANE per (batch, channel) row, the
 * states in registers, no cross-lane scan.  Same arguments as fd_selective_scan_fwd_merge except that Bt, Ct are
 * TIME-MAJOR (batch, 4, L, dstate) fp32 (fd_xdt_proj_tc with bc_layout = 1).  dim/4 % 32 == 0, L % 8 == 0; 16-bit io. */
int fd_selective_scan_fwd_merge_cl(const void* u, const void* delta, const float* A, const float* Bt, const float* Ct,
                                   const float* D, const float* delta_bias, void* y_nhwc, int batch, int dim, int H, int W,
                                   int dstate, int delta_softplus, int io_dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / 1x1 GEMM with fused epilogue — replaces F.conv2d / nn.Linear call sites:
 *   WeightStandardizedConv2d 3x3 (src/DADiff.py:139-154; standardisation folded into `weight` by the host),
 *   Downsample 4x4 s2 (:135-136), Upsample nearest x2 + 3x3 (:128-133, `upsample`=1), plain 3x3 (:643, 675),
 *   res_conv 1x1 (:407-408), SS2D in_proj/out_proj (src/emamba2.py:474, 519, 717, 748),
 *   TransposedAttention qkv / project_out 1x1 (src/DADiff.py:258, 260), torch.cat inputs (:727, 733).
 *
 *   acc[m, n]  = sum_{kh,kw,ci} in[b, ho*stride-pad+kh, wo*stride-pad+kw, ci] * weight[(b,) n, kh, kw, ci]
 *   v          = acc + bias[n];  if (n >= silu_from) v = silu(v)
 *   out[m, n]  = (addend ? addend[m, n] : 0) + (gate ? gate[b*gate_stride + n] : 1) * v
 *   gn_sums[b, n / (Cout/gn_groups), 0:2]  = (sum v, sum v^2)          (optional GroupNorm statistics of the output)
 * GroupNorm statistics are REPRODUCIBLE when `gn_ws` is given (tcgen05 path): every thread block stores the partial sums of the
 * tiles it owns of a sample into its own slot — the slot is the residue class of the sample's tile index modulo the (fixed)
 * grid size, so the grouping and the order of additions depend on the sample only, not on the batch it sits in or on block
 * timing — and a second, tiny launch adds the slots in index order.  Without gn_ws the sums are accumulated with
 * floating-point atomics (gn_sums must then be zero on entry, and the last bits depend on the block schedule).
 * `in` is the channel concatenation of src0 (c0 channels) and src1 (c1 channels, may be NULL/0).
 * weight: (Cout, KH, KW, c0+c1) in `dtype`, or (B, Cout, KH, KW, c0+c1) when per_batch_weight.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* src0;
    const void* src1;
    const void* weight;
    const float* bias;
    const float* gate;
    const void* addend;
    void* out;
    float* gn_sums;
    const void* weight_up4; /* upsample only (tcgen05 path): (4, Cout, 2, 2, c0) phase-summed weights, see fd_conv_tc.cu */
    float* gn_ws;         /* optional: fd_conv_gn_ws_floats(B) floats, ZERO on entry (one 16-float slot per sample and thread block) */
    const float* ln_v;    /* optional (tcgen05 path, 1x1 stride 1, c1 == 0): LayerNorm over the c0 input channels of every pixel FOLDED */
                          /*   into this GEMM: out = rstd_p * acc + ln_v[b, n] with acc = W'_b x_p on the RAW input.  W'_b (zero row  */
                          /*   sums) and ln_v (B, Cout) fp32 come from fd_ln_fold; rstd_p is computed inside the kernel from the staged */
                          /*   operand tile. */
    const float* ln_rstd; /* optional with ln_v: (B, Hin*Win) fp32 per-pixel 1/sqrt(var + eps) of the input rows from fd_row_rstd (or from */
                          /*   the kernel that produced the rows).  The GEMM then runs without its statistics warps, at the speed of */
                          /*   the plain 1x1 kernel; ln_eps is not used. */
    int c0, c1;
    int ld0;              /* row pitch (elements) of src0; 0 = dense (c0).  Lets a GEMM read a channel slice of a wider tensor */
    int B, Hin, Win, Cout;
    int KH, KW, stride, pad, upsample;
    int silu_from;        /* >= Cout: no activation */
    int gate_stride;      /* row stride (floats) of gate, e.g. 6*C for the adaLN modulation tensor */
    int gn_groups;        /* used when gn_sums != NULL */
    int per_batch_weight;
    int dtype;
    int relu_out;         /* 1: out = max(out, 0) after bias / gate / addend (the DA-CLIP ResNet blocks, src/DACLIP.py) */
    float ln_eps;         /* LayerNorm epsilon of the folded norm (used when ln_u != NULL) */
    int ab_dtype_p1;      /* 0: src0 / src1 / weight have type `dtype`; else 1 + their fd_dtype while out / addend keep `dtype`
                             (mixed 16-bit storage: fp16 residual stream, bf16 block-internal tensors; the tensor cores
                             need both operands in ONE 16-bit format, the output type is free) */
} fd_conv_params;

/* Size (floats) of fd_conv_params.gn_ws for a batch of B samples on the current device. */
long fd_conv_gn_ws_floats(int B);

/* CUDA-core fp32-accumulate path (any dtype; the fp32 validation path and the fallback for odd shapes). */
int fd_conv2d_simt(const fd_conv_params* p, cudaStream_t stream);

/* tcgen05 / TMEM / TMA path (dtype bf16 or fp16, c0 and c1 multiples of 64 for k>1, ...).  A plan owns the TMA
 * descriptors of one call site (pointers and shapes are baked in); create once, run many times. */
typedef struct fd_gemm_plan fd_gemm_plan;
int fd_conv2d_tc_supported(const fd_conv_params* p);
int fd_conv2d_tc_plan_create(const fd_conv_params* p, fd_gemm_plan** plan);
int fd_conv2d_tc_run(const fd_gemm_plan* plan, cudaStream_t stream);
void fd_conv2d_tc_plan_destroy(fd_gemm_plan* plan);

/* 2x2 average pooling, stride 2, channels-last (B,H,W,C) -> (B,H/2,W/2,C); H, W even, C % 8 == 0 for 16-bit types.
 * nn.AvgPool2d(2) of the DA-CLIP ModifiedResNet (stem, strided bottlenecks, their downsample branch). */
int fd_avgpool2x2_nhwc(const void* in, void* out, int B, int H, int W, int C, int dtype, cudaStream_t stream);

/* init_conv: 7x7, pad 3, over cat(x_t, x_input) (two fp32 single-channel images) -> (B,H,W,Cout) in `dtype`.
 * Replaces Unet.init_conv (src/DADiff.py:558, 700) + torch.cat (:1160).  weight: (Cout, 2, 7, 7) fp32. */
int fd_init_conv7x7(const float* x_t, const float* x_input, const float* weight, const float* bias, void* out,
                    int B, int H, int W, int Cout, int dtype, cudaStream_t stream);

/* Tensor-core form of the same op (16-bit output types, Cout == 64, H % 8 == 0, W % 16 == 0).  The fp32 images are split
 * into fp16 hi + lo parts inside the kernel (no input rounding); w16: host-packed (64, 256) fp16, K index
 * ci*56 + ky*8 + kx with a zero weight at kx = 7 (112 of 128 used), the block stored twice for the hi and lo parts
 * (ops.pack_init_conv_weights). */
int fd_init_conv7x7_tc(const float* x_t, const float* x_input, const void* w16, const float* bias, void* out, int B, int H,
                       int W, int Cout, int dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * LayerNorm over C + adaLN modulate:  out = (LN(x) * gamma + beta) * (1 + scale[b]) + shift[b]
 * Replaces nn.LayerNorm + modulate() (src/DADiff.py:450-451, 459, 461, 486-487).  gamma/beta may be NULL
 * (norm2: elementwise_affine=False, eps=1e-6).  shift/scale: fp32, element (b, c) at [b*mod_stride + c].
 * --------------------------------------------------------------------------------------------------------- */
int fd_ln_modulate(const void* x, void* out, const float* gamma, const float* beta, const float* shift,
                   const float* scale, int mod_stride, int B, int P, int C, float eps, int dtype,
                   cudaStream_t stream);

/* LayerNorm + adaLN modulate folded into the 1x1 GEMM that follows it (src/DADiff.py:486-487 -> src/emamba2.py:717 in_proj,
 * src/DADiff.py:258 qkv).  With xhat = (x - mean) rstd,  g = gamma (1 + scale_b),  h = beta (1 + scale_b) + shift_b
 * (gamma / beta NULL: 1 / 0):   W ((xhat gamma + beta)(1 + scale) + shift) = rstd (W g)(x - mean 1) + W h.
 *   Wf[b, o, c] = W[o, c] g[c] - rowmean_c(W[o, :] g), rounded to `dtype`: every row sums to zero, so Wf (x - mean 1) = Wf x
 *   and the GEMM runs on the RAW activations;  v[b, o] = sum_c W[o, c] h[c].
 *   W: (Cout, C) fp32; shift / scale: rows of mod_stride floats per sample.
 * The consumer is fd_conv2d_tc with per_batch_weight = 1, weight = Wf, ln_v = v: out = rstd_p (Wf x_p) + v. */
int fd_ln_fold(const float* W, const float* gamma, const float* beta, const float* shift, const float* scale, int mod_stride,
               void* Wf, float* v, int B, int Cout, int C, int dtype, cudaStream_t stream);

/* Same with separate storage types for x (in_dtype) and out (out_dtype); both 16-bit or both fp32. */
int fd_ln_modulate_io(const void* x, void* out, const float* gamma, const float* beta, const float* shift,
                      const float* scale, int mod_stride, int B, int P, int C, float eps, int in_dtype, int out_dtype,
                      cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * SS2D producer: depthwise 3x3 (+bias) + SiLU over the x half of `xz`, written in the 4-direction scan layout.
 * Replaces SS2D.conv2d + act (src/emamba2.py:480-488, 722) + EfficientScan.forward (:186-213).
 * xz: (B,H,W,ld) rows, the D conv channels are columns [0, D).  w: (D,3,3) fp32.  xs: (B,4,D,L), L = H/2*W/2:
 *   k=0: (even h, even w) row-major; k=1: (odd h, even w) column-major; k=2: (even h, odd w) row-major;
 *   k=3: (odd h, odd w) column-major.                                   H and W must be even.
 * --------------------------------------------------------------------------------------------------------- */
int fd_dwconv3x3_silu_scan(const void* xz, int ld, const float* w, const float* bias, void* xs, int B, int H,
                           int W, int D, int dtype, cudaStream_t stream);

/* x_proj + dt_proj (src/emamba2.py:335-340): per direction k,
 *   x_dbl = x_proj_w[k] (R+2N, D) @ xs[b,k] (D, L);  dts = dt_w[k] (D, R) @ x_dbl[:R];  Bs = x_dbl[R:R+N];  Cs = rest.
 * dts: (B,4,D,L) in `dtype`; Bs, Cs: (B,4,N,L) fp32.  Weights fp32. */
int fd_xdt_proj(const void* xs, const float* x_proj_w, const float* dt_w, void* dts, float* Bs, float* Cs, int B,
                int D, int L, int R, int N, int dtype, cudaStream_t stream);

/* Tensor-core variant (dtype bf16 / fp16): xw16 = x_proj_w in `dtype`, rows zero-padded to a multiple of 16:
 * (4, CCp, D); dw16 = dt_w in `dtype`, columns zero-padded to Rp in {16, 32}: (4, D, Rp).  Same outputs. */
int fd_xdt_proj_tc(const void* xs, const void* xw16, const void* dw16, void* dts, float* Bs, float* Cs, int B, int D,
                   int L, int R, int N, int Rp, int bc_layout, const float* dt_bias, int delta_softplus, int dtype,
                   cudaStream_t stream);
/* bc_layout: 0 = Bs, Cs as (B,4,N,L); 1 = time-major (B,4,L,N), the layout fd_selective_scan_fwd_merge_cl reads.
 * dt_bias != NULL: dts = dt_proj(...) + dt_bias[k*D + d], then softplus if delta_softplus (src/emamba2.py:344-359 applies
 * both inside selective_scan_fn); the scan is then called with delta_bias = NULL, delta_softplus = 0.  Both options need
 * D % 32 == 0 and L % 8 == 0 (FD_ERR_UNSUPPORTED otherwise). */

/* x_proj alone on the tensor cores (levels with dt_rank <= 8): x_dbl (B, 4, R+2N, L) fp32 = einsum(xs, x_proj_weight)
 * (src/emamba2.py:334-336).  Rows [0,R) are the low-rank dt input, [R,R+N) B, [R+N,R+2N) C; fd_selective_scan_fwd_merge_xdbl
 * consumes the tensor as is.  xw16: the zero-padded 16-bit (4, ceil16(R+2N), D) copy made by the host.  Needs D % 32 == 0,
 * L % 8 == 0. */
int fd_x_proj_tc(const void* xs, const void* xw16, float* x_dbl, int B, int D, int L, int R, int N, int dtype,
                 cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * TIME-MAJOR SS2D core (16-bit storage): the same four reference stages — SS2D.conv2d + act + EfficientScan
 * (src/emamba2.py:480-488, 722, 186-213), x_proj / dt_proj (:335-340), SelectiveScan.forward -> selective_scan_cuda_core.fwd
 * (:124-157, 342-355), EfficientMerge (:238-262) — on tensors whose CHANNEL index is fastest:
 *   xs_tm (B,4,L,D) 16-bit, xdbl_tm (B,4,L,XR) fp32, dts_tm (B,4,L,D) 16-bit; direction / step numbering as fd_dwconv3x3_silu_scan.
 * A lane of the scan owns a channel and walks time with its states in registers; rows are cut into segments whose entry states
 * come from an exact carry pass (founddiff_b200/csrc/fd_ss2d_tm.cu).
 * --------------------------------------------------------------------------------------------------------- */
/* xz: (B,H,W,ld) rows, conv channels = columns [0, D); w_tap_major: (9, D) fp32 (tap = 3*kh + kw); H, W even; D % 4 == 0. */
int fd_dwconv3x3_silu_tm(const void* xz, int ld, const float* w_tap_major, const float* bias, void* xs_tm, int B, int H,
                         int W, int D, int dtype, cudaStream_t stream);
/* x_proj (+ dt_proj) on time-major rows.  xw16 / dw16: the padded 16-bit weights of fd_xdt_proj_tc ((4, ceil16(R+2N), D), (4, D, Rp)).
 * fuse_dt = 1: xdbl_tm rows are [dt input (R) | B (N) | C (N)] (the scan applies dt_proj itself), dw16 / dts_tm / dt_bias unused.
 * fuse_dt = 0: xdbl_tm rows are [B (N) | C (N)] and dts_tm = softplus(dt_proj(x_dbl[:R]) + dt_bias[k*D + d]) (softplus threshold
 * 20, as selective_scan_fn with delta_softplus=True).  D % 64 == 0, R and N even, R + 2N <= 96. */
int fd_x_proj_tm(const void* xs_tm, const void* xw16, float* xdbl_tm, const void* dw16, void* dts_tm, const float* dt_bias,
                 int B, int D, int L, int R, int N, int Rp, int fuse_dt, int dtype, cudaStream_t stream);
/* Number of segments fd_selective_scan_tm cuts a row into for this geometry (sizes the carry workspace: B*4*S*2*dstate*D floats). */
int fd_scan_tm_segments(int B, int D, int H, int W);
/* The kernel `segments = 0` selects: > 0 = segmented channel-per-lane scan with that many segments; -8 / -4 = the time-sliced
 * cooperative scan with 8 / 4 warps per 32-channel block (few, long rows; fused dt_proj, dstate <= 8; no carry workspace). */
int fd_scan_tm_plan(int B, int D, int H, int W, int dstate, int dt_rank_fused);
/* S6 scan + EfficientMerge on time-major inputs: y_nhwc (B,H,W,D) = merge(scan(u, delta, A, B, C) + D_skip * u).
 * A: (4D, dstate) fp32 (= -exp(A_logs)); D_skip: (4D,).  dt_rank_fused > 0: delta = softplus(dt_w[d, :] . xdbl[l, :R] + dt_bias[d])
 * with dt_w (4D, R) fp32, dts_tm unused; dt_rank_fused == 0: delta is read from dts_tm as is.  segments: 0 = automatic
 * (fd_scan_tm_plan), > 0 = that many segments, -8 / -4 = time-sliced kernel.
 * carry_ws: fp32 scratch of carry_floats elements (may be NULL when one segment is used).  D % 128 == 0;
 * (dstate, dt_rank_fused) in {(4,4), (8,4), (8,8), (16,8), (4,0), (8,0), (16,0), (32,0)}. */
int fd_selective_scan_tm(const void* u_tm, const void* dts_tm, const float* xdbl_tm, const float* A, const float* dt_w,
                         const float* dt_bias, const float* D_skip, float* carry_ws, long carry_floats, void* y_nhwc, int B,
                         int D, int H, int W, int dstate, int dt_rank_fused, int segments, int io_dtype, cudaStream_t stream);
/* The time-sliced scan with CHAINED SEGMENTS (same op, same bits: src/emamba2.py:124-157 + EfficientMerge :238-262).  A row is cut
 * into fd_scan_tm_chain_plan(...) segments that run as separate short blocks; the state leaving a segment reaches its successor
 * through chain_ws behind a release / acquire flag, and blocks draw (segment, row) tickets in start order (a predecessor is
 * always resident or finished: no deadlock).  No exp(dt A) is evaluated twice.  What it buys is balance: 64 x D/32 whole-row
 * blocks on 2 x 148 slots leave the SMs that hold one block idle for the second half of the launch; short blocks refill the
 * slots (level 0 at B = 16: 1.57 -> 1.48 ms inside the step).  fd_scan_tm_chain_plan returns the segment count (0: not chained for
 * this geometry — call fd_selective_scan_tm) and the workspace size in floats.  chain_ws must be ZERO-FILLED ONCE by the caller
 * before its first use and left alone afterwards (every launch leaves its counter and flags zeroed again); one workspace per
 * concurrently running launch.  Other arguments as fd_selective_scan_tm with dt_rank_fused > 0. */
int fd_scan_tm_chain_plan(int B, int D, int H, int W, int dstate, int dt_rank_fused, int* ws_floats);
int fd_selective_scan_tm_chained(const void* u_tm, const float* xdbl_tm, const float* A, const float* dt_w, const float* dt_bias,
                                 const float* D_skip, float* chain_ws, long chain_floats, void* y_nhwc, int B, int D, int H,
                                 int W, int dstate, int dt_rank_fused, int io_dtype, cudaStream_t stream);

/* SS2D consumer: EfficientMerge (src/emamba2.py:238-262) + out_norm LayerNorm(D) (:365) + y*z + local (:747-748).
 * ys: (B,4,D,L); z = columns [z_off, z_off+D) of xz rows (already SiLU'd); local: (B, D) fp32; out: (B,H,W,D).
 * stats_ws: caller-provided workspace of B*H*W*2 floats (per-pixel mean / rstd). */
int fd_merge_ln_gate(const void* ys, const void* xz, int ld, int z_off, const float* gamma, const float* beta,
                     const float* local, float* stats_ws, void* out, int B, int H, int W, int D, float eps, int dtype,
                     cudaStream_t stream);

/* Row-wise tail of SS2D on channels-last data: out = (LN_C(y) * gamma + beta) * z + local[b]  (src/emamba2.py:365,
 * 747-748); z = columns [z_off, z_off+C) of xz rows of pitch ld; local: (B, C) fp32. */
/* rstd[b, p] = 1 / sqrt(var_C(x[b, p, :]) + eps) (biased variance, two-pass in registers), x: (B*P, C) rows of a 16-bit type.
 * The LayerNorm statistics a folded GEMM needs (fd_conv_params.ln_rstd): one read of the rows, 4 bytes written per row.
 * C in {64, 128, 256, 512}. */
int fd_row_rstd(const void* x, float* rstd, long rows, int C, float eps, int dtype, cudaStream_t stream);

int fd_ln_gate(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta, const float* local,
               void* out, int B, int P, int C, float eps, int dtype, cudaStream_t stream);

/* fd_ln_gate fused with out_proj and the Mamba block's gated residual (src/emamba2.py:365, 747-748 -> src/DADiff.py:486):
 *   out[p, :] = addend[p, :] + gate[b, :] * (w . ((LN_D(y[p, :]) * gamma + beta) * z[p, :] + local[b, :]))
 * one pass over the pixels instead of fd_ln_gate + a 1x1 convolution: the gated row never touches HBM.  w: (Cout, D) row-major in
 * io_dtype (= storage of y / xz and operand type of the product); gate: (B, Cout) fp32 rows of pitch gate_stride; addend / out:
 * (B, P, Cout) in out_dtype.  Built for the full-resolution level, D = 128, Cout = 64, P % 16 == 0, 16-bit types
 * (fd_ln_gate_out_proj_supported says whether a geometry is taken; otherwise use the two calls). */
int fd_ln_gate_out_proj_supported(int P, int D, int Cout, int ld, int z_off, int io_dtype, int out_dtype);
int fd_ln_gate_out_proj(const void* y, const void* xz, int ld, int z_off, const float* gamma, const float* beta, const float* local,
                        const void* w, const float* gate, int gate_stride, const void* addend, void* out, int B, int P, int D, int Cout,
                        float eps, int io_dtype, int out_dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * TransposedAttention (src/DADiff.py:263-285), C/32 heads of 32 channels.
 * fd_dwconv3x3_qkv_gram: depthwise 3x3 over qkv (B,H,W,3C); writes v (B,H,W,C), gram[b, head, i, j] = sum_p q_i k_j and
 *   qk_sq[b, 0:2, c] = sum_p q_c^2, k_c^2.  REPRODUCIBLE: every block stores the partial of its pixel chunk into `ws`
 *   (fd_gram_ws_floats floats, ZERO on entry) and the last block of a (sample, head) adds the chunks in order — no
 *   floating-point atomics, the same bits for a slice in any batch.
 * fd_attn_weff: attn = softmax(gram / (max(|q_i|,1e-12) max(|k_j|,1e-12)) * temperature[head]) and folds it
 *   into the output projection: weff[b, o, h*32+j] = sum_i proj_w[o, h*32+i] * attn[b,h,i,j]   (dtype),
 *   so that project_out(attn @ v) == v @ weff[b]^T, a per-sample 1x1 GEMM (fd_conv2d_*, per_batch_weight).
 * --------------------------------------------------------------------------------------------------------- */
long fd_gram_ws_floats(int B, int H, int W, int C, int dtype);
int fd_dwconv3x3_qkv_gram(const void* qkv, const float* w, void* v, float* gram, float* qk_sq, float* ws, int B, int H,
                          int W, int C, int dtype, cudaStream_t stream);
/* 16-bit storage types (bf16 / fp16) split the first step into two streaming kernels:
 *   fd_dwconv3x3_nhwc: depthwise 3x3 (+bias, +SiLU) over a channels-last (B,H,W,C) tensor, register sliding window;
 *                      NB its weights are TAP-MAJOR: w is (9, C) fp32 (the other dwconv entry points take (C, 9));
 *   fd_gram_qk:        gram / qk_sq (same meaning and the same `ws` contract as above) from q = columns [0,C), k = columns [C,2C) of
 *                      rows of pitch `ld` — q.k^T and the norms run on the tensor cores (mma.sync, fp32 accumulate).
 * v is then read in place (columns [2C,3C), ld0 = 3C) by the per-sample W_eff GEMM. */
int fd_dwconv3x3_nhwc(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int C, int silu,
                      int dtype, cudaStream_t stream);
int fd_gram_qk(const void* qkv, int ld, float* gram, float* qk_sq, float* ws, int B, int P, int C, int dtype, cudaStream_t stream);
int fd_attn_weff(const float* gram, const float* qk_sq, const float* temperature, const float* proj_w,
                 void* weff, int B, int C, int dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * GroupNorm(G) + SiLU + skip add (src/DADiff.py:213-229, 426-430):
 *   fd_gn_stats:     sums[b, g, 0:2] = (sum, sum of squares) over the group.  Reproducible: block partials go to `ws`
 *                    (fd_gn_stats_ws_floats(B, P, G) floats, ZERO on entry) and are added in block order by the last block.
 *   fd_gn_silu_add:  out = silu((y - mean) * rstd * gamma + beta) + (skip ? skip : 0)
 * --------------------------------------------------------------------------------------------------------- */
long fd_gn_stats_ws_floats(int B, int P, int G);
int fd_gn_stats(const void* y, float* sums, float* ws, int B, int P, int C, int G, int dtype, cudaStream_t stream);
int fd_gn_silu_add(const void* y, const float* sums, const float* gamma, const float* beta, const void* skip,
                   void* out, int B, int P, int C, int G, float eps, int dtype, cudaStream_t stream);

/* Secondary path (lucidrains Unet, src/denoising_diffusion_pytorch.py): Block with the time-embedding scale/shift
 * (:183-199, 201-225): out = silu(GN(y) * (scale[b]+1) + shift[b]) + skip; scale/shift fp32 at [b*ss_stride + c]. */
int fd_gn_scale_shift_silu(const void* y, const float* sums, const float* gamma, const float* beta, const float* scale,
                           const float* shift, int ss_stride, const void* skip, void* out, int B, int P, int C, int G,
                           float eps, int dtype, cudaStream_t stream);
/* Secondary path: bottleneck Attention (:257-279), heads of 32 channels over N = H*W tokens, flash-style:
 * out[b, i, h*32 + d] = sum_j softmax_j(scale * q_i . k_j) v_j[d];  qkv (B, N, 3*heads*32) = [q | k | v] channels-last
 * ('b (h c)' order, as produced by the to_qkv 1x1 GEMM); out (B, N, heads*32).  dtype bf16 / fp16. */
int fd_flash_attn_d32(const void* qkv, void* out, int B, int N, int heads, float scale, int dtype, cudaStream_t stream);
/* The same op on tcgen05 / tensor memory (fd_flash_attn_tc.cu): 128-query CTAs, S = Q K^T and O_j = P_j V_j as tcgen05.mma with
 * TMEM accumulators, softmax by the thread that owns the TMEM lane of its query row.  Any N; 16-byte aligned pointers. */
int fd_flash_attn_d32_tc(const void* qkv, void* out, int B, int N, int heads, float scale, int dtype, cudaStream_t stream);

/* Secondary path: LinearAttention (:227-255) between to_qkv and to_out, same qkv layout as above.
 *   fd_linattn_context: ctx_raw[b,h,d,e] += sum_n exp(k[n,d] - kmax[d]) v[n,e];  ksum[b,hd] += sum_n exp(k - kmax)
 *                       (kmax pre-filled with -inf and ksum / ctx_raw with 0 by the caller; two launches: column max, then
 *                       the tensor-core accumulation);
 *   fd_linattn_weff:    weff[b,o,h*32+d] = scale/(N*ksum[d]) * sum_e wout[o,h*32+e] ctx_raw[b,h,d,e], so that
 *                       to_out.0(out) == softmax_d(q) @ weff[b]^T + bias (a per-sample 1x1 GEMM, fd_conv2d_*);
 *   fd_softmax_d32:     qhat (B,N,heads*32) = softmax over each head's 32 channels of q.
 * The channel LayerNorm of to_out.1 is fd_ln_modulate with zero shift/scale. */
int fd_linattn_context(const void* qkv, float* kmax, float* ksum, float* ctx_raw, int B, int N, int heads, int dtype,
                       cudaStream_t stream);
int fd_linattn_weff(const float* ctx_raw, const float* ksum, const float* wout, void* weff, int B, int N, int heads, int dim,
                    float scale, int dtype, cudaStream_t stream);
int fd_softmax_d32(const void* qkv, void* qhat, int B, int N, int heads, int dtype, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Small dense layers on (B, K) fp32 vectors: time_mlp, adaLN_modulation (src/DADiff.py:173-185, 463-466,
 * 580-585).  out[b, n] = act_out( sum_k act_in(x[b,k]) * W[n,k] + bias[n] ) + add[b,n]
 * act: 0 none, 1 SiLU, 2 GELU(erf), 3 ReLU.
 * --------------------------------------------------------------------------------------------------------- */
int fd_linear_small(const float* x, const float* W, const float* bias, const float* add, float* out, int B, int K,
                    int N, int act_in, int act_out, cudaStream_t stream);
/* SinusoidalPosEmb (src/DADiff.py:173-185): out[b] = cat(sin(t_b f_k), cos(t_b f_k)), f_k = exp(-k ln(1e4)/(dim/2-1)) */
int fd_time_sinusoid(const float* time, float* out, int B, int dim, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Sampler (src/DADiff.py:1153-1209 model_predictions 'pred_res'; :1221-1230 p_sample; :1142-1151 q_posterior;
 * :1317-1344 ddim update).  Images are fp32 (B, H*W).
 * fd_sampler_init:  x_input = 2*ldct - 1;  x_t = x_input + noise_scale * noise;  first = (x_t + 1)/2
 * fd_final_conv_update: final_conv 1x1 (C -> 1, src/DADiff.py:683, 740) fused with the whole update:
 *   pred_res  = clamp(feat . w + bias, -1, 1)
 *   x_start   = clamp(x_input - pred_res, -1, 1)
 *   pred_noise= (x_t - x_input - (acs - 1) * pred_res) / bcs                       (written if non-NULL)
 *   x_next    = c_xt * x_t + c_res * pred_res + c_x0 * x_start + c_noise * noise     (noise may be NULL)
 *   with coef (DEVICE, fp32[6]) = {c_xt, c_res, c_x0, c_noise, acs = alphas_cumsum[t], bcs = betas_cumsum[t]}.
 *   DDIM (eta=0): {1, -(acs_t - acs_next), 0, 0} and {0, 0, 1, 0} for the last pair;  ancestral:
 *   {coef1[t], coef2[t], coef3[t], exp(0.5*logvar[t]) or 0 at t=0}.
 * fd_final_conv_update_obj: the same kernel for every objective branch of model_predictions (src/DADiff.py:1168-1207),
 *   `objective` = FD_OBJ_*; feat1 / w1 / bias1 = the second Unet's final_conv operands (num_unet = 2, UnetRes.forward
 *   src/DADiff.py:817-820), read only by the two-output objectives (NULL otherwise); coef is fp32[8] with
 *   coef[6] = one_minus_alphas_cumsum[t]:
 *     FD_OBJ_PRED_RES        o0 = pred_res: as above (also 'pred_res_noise' with test_res_or_noise = "res", :1176-1182)
 *     FD_OBJ_PRED_NOISE      o0 = pred_noise; x_start = clamp((x_t - acs x_input - bcs o0) / coef[6]);
 *                            pred_res = clamp(x_input - x_start)          (:1194-1201; test_res_or_noise = "noise" :1183-1189)
 *     FD_OBJ_PRED_RES_NOISE  pred_res = clamp(o0), pred_noise = o1, x_start = clamp(x_t - acs pred_res - bcs o1)  (:1169-1175)
 *     FD_OBJ_PRED_X0_NOISE   pred_res = clamp(x_input - o0), pred_noise = o1, x_start = clamp(o0)                 (:1188-1192)
 *   the update line (x_next) is the same for all of them: neither q_posterior (:1142-1151) nor the 'use_pred_noise'
 *   DDIM step (:1344) reads pred_noise.
 * fd_unnormalize: out = (x + 1) / 2.
 * --------------------------------------------------------------------------------------------------------- */
#define FD_OBJ_PRED_RES 0
#define FD_OBJ_PRED_NOISE 1
#define FD_OBJ_PRED_RES_NOISE 2
#define FD_OBJ_PRED_X0_NOISE 3
int fd_sampler_init(const float* ldct, const float* noise, float noise_scale, float* x_input, float* x_t,
                    float* first, long n, cudaStream_t stream);
int fd_final_conv_update(const void* feat, const float* w, const float* bias, const float* x_input,
                         const float* x_t, const float* noise, const float* coef, float* x_next, float* pred_res,
                         float* pred_noise, float* x_start, long npix, int C, int dtype, cudaStream_t stream);
int fd_final_conv_update_obj(const void* feat, const float* w, const float* bias, const void* feat1, const float* w1,
                             const float* bias1, const float* x_input, const float* x_t, const float* noise,
                             const float* coef, float* x_next, float* pred_res, float* pred_noise, float* x_start,
                             long npix, int C, int dtype, int objective, cudaStream_t stream);
int fd_unnormalize(const float* x, float* out, long n, cudaStream_t stream);
/* Evaluation metrics of Trainer.test (src/DADiff.py:1883-1888; src/util.py:188-236 compute_psnr / compute_ssim /
 * compute_rmse) on the device: out[2*b] = sum over the slice of (pred - target)^2, out[2*b + 1] = sum of the SSIM map
 * (11x11 Gaussian window sigma 1.5, reflect border, C1 = (0.01 max_val)^2, C2 = (0.03 max_val)^2, map clamped to [0, 1]).
 * pred, target: (B, H, W) fp32.  PSNR = 10 log10(max_val^2 / (sse / HW)), RMSE = sqrt(sse / HW), SSIM = sum / HW. */
int fd_slice_metrics(const float* pred, const float* target, float* out, int B, int H, int W, float max_val,
                     cudaStream_t stream);

/* Secondary path: epsilon-prediction GaussianDiffusion update (src/denoising_diffusion_pytorch.py:547-576, 588-595,
 * 612-646), fused: x0 = sr*x_t - srm1*eps (clipped to [-1,1] if coef[6]); x_next = a0*x0 + a1*x_t + a2*eps + a3*noise;
 * coef (DEVICE fp32[8]) = {sr, srm1, a0, a1, a2, a3, clip, 0}. */
int fd_ddpm_update(const float* x_t, const float* eps, const float* noise, const float* coef, float* x_next, float* x_start,
                   long n, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Step programs — `unet_step` of SURVEY 8(b): one C call launches a whole timestep on caller-owned memory.
 * Replaces, per timestep, `Unet.forward` (src/DADiff.py:685-740; fd_unet_step) and additionally `model_predictions` + the
 * DDIM / posterior update (src/DADiff.py:1153-1209, 1221-1230, 1323-1344; fd_sample_step).  A plan file is recorded once by the
 * host module for a fixed (batch, H, W, storage type): founddiff_b200/program.py::export_step_program.  It holds the packed
 * weights, the launch list and the buffer layout; loading it resolves every pointer into the caller's arena and builds the
 * TMA descriptors of the convolutions.  Steps are stream-ordered, allocation-free and CUDA-graph capturable.
 *   per step the caller writes (device memory, fp32): "time" (B) = alphas_cumsum[t] * 1000 (src/DADiff.py:1161-1163),
 *   "coef" (8) = {c_xt, c_res, c_x0, c_noise, alphas_cumsum[t], betas_cumsum[t], one_minus_alphas_cumsum[t], 0}, "noise" (B, H*W)
 *   when c_noise != 0; per slice batch: "x_input" and "x_t" (B, H*W, in [-1, 1]) and the conditioning vectors "prompt_emb" /
 *   "local.<block>" (from the DA-CLIP embeddings; the plan file carries the values of the slices it was recorded with).
 *   After fd_sample_step "x_t" holds x_{t-1}.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct fd_program fd_program;
long fd_program_arena_bytes(const char* path);                        /* < 0: not a plan file */
int fd_program_load(const char* path, void* arena, long arena_bytes, fd_program** prog);
void* fd_program_buffer(const fd_program* prog, const char* name, long* nbytes);   /* NULL: no such buffer */
int fd_program_num_launches(const fd_program* prog);
int fd_unet_step(const fd_program* prog, cudaStream_t stream);
int fd_sample_step(const fd_program* prog, cudaStream_t stream);
void fd_program_destroy(fd_program* prog);

#ifdef __cplusplus
}
#endif
#endif /* FOUNDDIFF_B200_H */
