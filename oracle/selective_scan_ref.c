/*
 * CPU restatement of the S6 selective scan forward that FoundDiff calls through the third-party
 * extension `selective_scan_cuda_core.fwd` / `selective_scan_cuda.fwd` (call sites
 * /root/reference/src/emamba2.py:152,154; argument conventions :124-157; shape/flop model :38-59,63-110).
 *
 * The extension itself is NOT vendored in the reference and no version is pinned anywhere
 * (SURVEY.md section 8c), so this file restates the published algorithm of state-spaces/mamba
 * `selective_scan_ref` (the semantics both VMamba's `selective_scan_cuda_core` and mamba_ssm's
 * `selective_scan_cuda` implement):
 *
 *     dt      = delta[b,d,l] + delta_bias[d];  if (softplus) dt = dt <= 20 ? log1p(exp(dt)) : dt
 *     h[n]    = exp(dt * A[d,n]) * h[n] + dt * B[b,g,n,l] * u[b,d,l]          (h[.] = 0 at l = 0)
 *     y[b,d,l]= sum_n h[n] * C[b,g,n,l] + D[d] * u[b,d,l]
 *
 * with g = d / (Dtot / G): B and C are shared by the channels of one direction group.
 * All tensors fp32, contiguous, layouts: u,delta,y (Bt, Dtot, L); A (Dtot, N); B,C (Bt, G, N, L).
 *
 * Pinned (tests/test_oracle_golden.py::test_scan_restatement_vs_published_kernel) against the outputs of a build of the
 * published kernel itself — vLLM 0.22's `torch.ops._C.selective_scan_fwd`, "adapted from state-spaces/mamba", run on a
 * B200 by oracle/gen_golden_scan_vllm.py (tests/golden/scan_vllm.npz) — and against an fp64 sequential evaluation.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, smoke() and bench.py's cpu_baseline; never by the product.
 * `acc64 != 0` carries the state in double (used to bound the fp32 restatement's own rounding).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define FD_MAX_N 256
#define FD_MAX_THREADS 256

typedef struct {
    const float *u, *delta, *A, *Bm, *Cm, *D, *delta_bias;
    float *y;
    int64_t Bt, Dtot, L, N, G;
    int delta_softplus, acc64;
    int64_t row_begin, row_end;
} fd_scan_job;

static void *fd_scan_rows(void *arg)
{
    const fd_scan_job *j = (const fd_scan_job *)arg;
    const float *u = j->u, *delta = j->delta, *A = j->A, *Bm = j->Bm, *Cm = j->Cm, *D = j->D;
    const float *delta_bias = j->delta_bias;
    float *y = j->y;
    const int64_t Dtot = j->Dtot, L = j->L, N = j->N, G = j->G;
    const int delta_softplus = j->delta_softplus, acc64 = j->acc64;
    const int64_t per_group = Dtot / G;
    for (int64_t row = j->row_begin; row < j->row_end; ++row) {
        const int64_t b = row / Dtot, d = row % Dtot, g = d / per_group;
        const float *ur = u + row * L, *dr = delta + row * L;
        const float *Br = Bm + (b * G + g) * N * L, *Cr = Cm + (b * G + g) * N * L;
        const float *Ar = A + d * N;
        const float bias = delta_bias ? delta_bias[d] : 0.f;
        const float Dd = D ? D[d] : 0.f;
        float *yr = y + row * L;
        if (!acc64) {
            float h[FD_MAX_N];
            for (int64_t n = 0; n < N; ++n) h[n] = 0.f;
            for (int64_t l = 0; l < L; ++l) {
                float dt = dr[l] + bias;
                if (delta_softplus && dt <= 20.f) dt = log1pf(expf(dt));
                const float uu = ur[l];
                float acc = 0.f;
                for (int64_t n = 0; n < N; ++n) {
                    h[n] = expf(dt * Ar[n]) * h[n] + dt * Br[n * L + l] * uu;
                    acc += h[n] * Cr[n * L + l];
                }
                yr[l] = acc + Dd * uu;
            }
        } else {
            double h[FD_MAX_N];
            for (int64_t n = 0; n < N; ++n) h[n] = 0.0;
            for (int64_t l = 0; l < L; ++l) {
                double dt = (double)dr[l] + (double)bias;
                if (delta_softplus && dt <= 20.0) dt = log1p(exp(dt));
                const double uu = ur[l];
                double acc = 0.0;
                for (int64_t n = 0; n < N; ++n) {
                    h[n] = exp(dt * (double)Ar[n]) * h[n] + dt * (double)Br[n * L + l] * uu;
                    acc += h[n] * (double)Cr[n * L + l];
                }
                yr[l] = (float)(acc + (double)Dd * uu);
            }
        }
    }
    return NULL;
}

/* nthreads <= 0: use every online core. Rows (b, d) are independent, so they are split evenly. */
int fd_oracle_selective_scan_fwd(const float *u, const float *delta, const float *A, const float *Bm,
                                 const float *Cm, const float *D, const float *delta_bias, float *y,
                                 int64_t Bt, int64_t Dtot, int64_t L, int64_t N, int64_t G,
                                 int delta_softplus, int acc64, int nthreads)
{
    if (N > FD_MAX_N || G <= 0 || Dtot % G != 0) return -1;
    const int64_t rows = Bt * Dtot;
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > FD_MAX_THREADS) nthreads = FD_MAX_THREADS;
    if (nthreads > rows) nthreads = (int)(rows > 0 ? rows : 1);
    fd_scan_job jobs[FD_MAX_THREADS];
    pthread_t tids[FD_MAX_THREADS];
    for (int t = 0; t < nthreads; ++t) {
        fd_scan_job jb = {u, delta, A, Bm, Cm, D, delta_bias, y, Bt, Dtot, L, N, G, delta_softplus, acc64,
                          rows * t / nthreads, rows * (t + 1) / nthreads};
        jobs[t] = jb;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&tids[t], NULL, fd_scan_rows, &jobs[t]);
    fd_scan_rows(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(tids[t], NULL);
    return 0;
}
