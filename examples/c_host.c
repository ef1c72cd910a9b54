/*
 * C host for the founddiff_b200 C ABI (include/founddiff_b200.h): the boundary is plain C — no torch, no C++ types.
 *
 *   c_host                       without arguments (works on a box with no GPU): loads the library, prints its version and
 *                                exercises the argument validation that precedes every CUDA call.
 *   c_host plan.fdp in.bin out.bin
 *                                runs a whole reverse-diffusion chain from a recorded STEP PROGRAM on cudaMalloc'ed memory:
 *                                the per-timestep work of ResidualDiffusion.ddim_sample / p_sample (src/DADiff.py:1275-1365,
 *                                1221-1230) — Unet.forward + model_predictions + update — is ONE call, fd_sample_step.
 *                                in.bin : int32 n_steps, B, P; float x_input[B*P]; float x_t[B*P];
 *                                         per step: float time, float coef[8], int32 has_noise, float noise[B*P] if has_noise
 *                                out.bin: float x_t[B*P] after the last step (in [-1, 1]; (x + 1) / 2 is the denoised slice)
 *
 *   gcc -std=c99 -DFD_HOST_WITH_CUDA -Iinclude -I/usr/local/cuda/include examples/c_host.c -o /tmp/c_host -Lfounddiff_b200 -lfounddiff_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/founddiff_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef FD_HOST_WITH_CUDA          /* the plan runner needs cudaMalloc / cudaMemcpy; the validation demo does not */
#include <cuda_runtime_api.h>
#endif
#include "founddiff_b200.h"

#ifdef FD_HOST_WITH_CUDA
static int fail(const char* what, int rc) {
    fprintf(stderr, "c_host: %s failed (%d)\n", what, rc);
    return 10;
}

static int run_plan(const char* plan, const char* in_path, const char* out_path) {
    long arena_bytes = fd_program_arena_bytes(plan);
    if (arena_bytes < 0) return fail("fd_program_arena_bytes", (int)arena_bytes);
    void* arena = NULL;
    if (cudaMalloc(&arena, (size_t)arena_bytes) != cudaSuccess) return fail("cudaMalloc", 0);
    fd_program* prog = NULL;
    int rc = fd_program_load(plan, arena, arena_bytes, &prog);
    if (rc) return fail("fd_program_load", rc);
    printf("plan %s: %ld MiB arena, %d launches per step\n", plan, arena_bytes >> 20, fd_program_num_launches(prog));

    FILE* f = fopen(in_path, "rb");
    if (!f) return fail("open input", 0);
    int hdr[3];
    if (fread(hdr, sizeof(int), 3, f) != 3) return fail("read header", 0);
    const int n_steps = hdr[0], B = hdr[1], P = hdr[2];
    const size_t n = (size_t)B * P;
    float* host = (float*)malloc(n * sizeof(float));
    float* tvec = (float*)malloc((size_t)B * sizeof(float));
    long nb = 0;
    float* d_xin = (float*)fd_program_buffer(prog, "x_input", &nb);
    float* d_xt = (float*)fd_program_buffer(prog, "x_t", &nb);
    float* d_time = (float*)fd_program_buffer(prog, "time", &nb);
    float* d_coef = (float*)fd_program_buffer(prog, "coef", &nb);
    float* d_noise = (float*)fd_program_buffer(prog, "noise", &nb);
    if (!host || !tvec || !d_xin || !d_xt || !d_time || !d_coef || !d_noise) return fail("buffers", 0);
    if (fread(host, sizeof(float), n, f) != n) return fail("read x_input", 0);
    cudaMemcpy(d_xin, host, n * sizeof(float), cudaMemcpyHostToDevice);
    if (fread(host, sizeof(float), n, f) != n) return fail("read x_t", 0);
    cudaMemcpy(d_xt, host, n * sizeof(float), cudaMemcpyHostToDevice);

    cudaStream_t stream;
    cudaStreamCreate(&stream);
    for (int s = 0; s < n_steps; ++s) {
        float time, coef[8];
        int has_noise = 0;
        if (fread(&time, sizeof(float), 1, f) != 1 || fread(coef, sizeof(float), 8, f) != 8 || fread(&has_noise, sizeof(int), 1, f) != 1)
            return fail("read step", s);
        for (int b = 0; b < B; ++b) tvec[b] = time;
        cudaMemcpyAsync(d_time, tvec, (size_t)B * sizeof(float), cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_coef, coef, sizeof(coef), cudaMemcpyHostToDevice, stream);
        if (has_noise) {
            if (fread(host, sizeof(float), n, f) != n) return fail("read noise", s);
            cudaMemcpyAsync(d_noise, host, n * sizeof(float), cudaMemcpyHostToDevice, stream);
        }
        cudaStreamSynchronize(stream);                    /* the staging vectors above are reused by the next step */
        rc = fd_sample_step(prog, stream);                /* Unet.forward + model_predictions + update: one call */
        if (rc) return fail("fd_sample_step", rc);
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail("cudaStreamSynchronize", (int)cudaGetLastError());
    fclose(f);
    cudaMemcpy(host, d_xt, n * sizeof(float), cudaMemcpyDeviceToHost);
    f = fopen(out_path, "wb");
    if (!f || fwrite(host, sizeof(float), n, f) != n) return fail("write output", 0);
    fclose(f);
    fd_program_destroy(prog);
    cudaFree(arena);
    free(host);
    free(tvec);
    printf("ran %d step(s) on %d slice(s) of %d pixels\n", n_steps, B, P);
    return 0;
}
#endif

int main(int argc, char** argv) {
    const char* v = fd_version();
    printf("%s\n", v);
    if (!strstr(v, "sm_100a")) return 1;
#ifdef FD_HOST_WITH_CUDA
    if (argc == 4) return run_plan(argv[1], argv[2], argv[3]);
#else
    (void)argc;
    (void)argv;
#endif
    /* selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows)  (src/emamba2.py:154) */
    int rc = fd_selective_scan_fwd(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, /*batch*/ 1, /*dim*/ 512, /*seqlen*/ 65536,
                                   /*dstate*/ 4, /*ngroups*/ 4, /*delta_softplus*/ 1, FD_BF16, (cudaStream_t)0);
    printf("fd_selective_scan_fwd(NULL...) -> %d (FD_ERR_BAD_ARGUMENT = %d)\n", rc, FD_ERR_BAD_ARGUMENT);
    if (rc != FD_ERR_BAD_ARGUMENT) return 2;
    if (fd_program_arena_bytes("/nonexistent.fdp") >= 0 || fd_unet_step(NULL, (cudaStream_t)0) != FD_ERR_BAD_ARGUMENT) return 4;
    rc = fd_final_conv_update_obj(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, 64, FD_BF16,
                                  FD_OBJ_PRED_RES_NOISE, (cudaStream_t)0);
    printf("fd_final_conv_update_obj(NULL...) -> %d\n", rc);
    return rc == FD_ERR_BAD_ARGUMENT ? 0 : 3;
}
