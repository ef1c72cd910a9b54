/*
 * Minimal C host for the founddiff_b200 C ABI (include/founddiff_b200.h): shows that the boundary is plain C — no torch, no C++
 * types — by linking the library from a C99 program.  Without a GPU it can still load the library, print its version and
 * exercise the argument validation (which precedes every CUDA call); with a GPU, replace the NULL pointers by cudaMalloc'ed
 * buffers laid out as the header documents.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -o /tmp/c_host -Lfounddiff_b200 -lfounddiff_b200 -Wl,-rpath,$PWD/founddiff_b200
 */
#include <stdio.h>
#include <string.h>

#include "founddiff_b200.h"

int main(void) {
    const char* v = fd_version();
    printf("%s\n", v);
    if (!strstr(v, "sm_100a")) return 1;
    /* selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows)  (src/emamba2.py:154) */
    int rc = fd_selective_scan_fwd(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, /*batch*/ 1, /*dim*/ 512, /*seqlen*/ 65536,
                                   /*dstate*/ 4, /*ngroups*/ 4, /*delta_softplus*/ 1, FD_BF16, (cudaStream_t)0);
    printf("fd_selective_scan_fwd(NULL...) -> %d (FD_ERR_BAD_ARGUMENT = %d)\n", rc, FD_ERR_BAD_ARGUMENT);
    if (rc != FD_ERR_BAD_ARGUMENT) return 2;
    rc = fd_final_conv_update_obj(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, 64, FD_BF16,
                                  FD_OBJ_PRED_RES_NOISE, (cudaStream_t)0);
    printf("fd_final_conv_update_obj(NULL...) -> %d\n", rc);
    return rc == FD_ERR_BAD_ARGUMENT ? 0 : 3;
}
