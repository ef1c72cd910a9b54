// tcgen05 / TMEM / TMA implicit-GEMM convolution (placeholder until the kernel lands: reports "unsupported",
// callers then use fd_conv2d_simt).
#include "fd_common.cuh"

struct fd_gemm_plan { int unused; };

extern "C" int fd_conv2d_tc_supported(const fd_conv_params*) { return 0; }
extern "C" int fd_conv2d_tc_plan_create(const fd_conv_params*, fd_gemm_plan** plan) {
    if (plan) *plan = nullptr;
    return FD_ERR_UNSUPPORTED;
}
extern "C" int fd_conv2d_tc_run(const fd_gemm_plan*, cudaStream_t) { return FD_ERR_UNSUPPORTED; }
extern "C" void fd_conv2d_tc_plan_destroy(fd_gemm_plan*) {}
