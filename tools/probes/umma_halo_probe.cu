// Probe: can a SWIZZLE_128B K-major UMMA A-descriptor address a *shifted window* of a halo tile?
//   smem halo tile: HT_H x HT_W pixels, 128 bytes (64 bf16 channels) per pixel, written with the TMA 128B swizzle
//   MMA rows m = y*8 + x (16 rows x 8 pixels) must read halo pixel (y+dy, x+dx): start = base + (dy*HT_W+dx)*128,
//   SBO = HT_W*128 (not a multiple of 1024).  Tries base_offset = 0 and base_offset = (start >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_halo_probe umma_halo_probe.cu ; prints mismatch counts.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define DEVINL __device__ __forceinline__
constexpr int HT_W = 10, HT_H = 18, NCOL = 64;

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}
__host__ __device__ inline float aval(int R, int k) { return (float)((R * 7 + k * 3) % 17 - 8); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k) % 13 - 6); }

__global__ void __launch_bounds__(128) probe(float* out, int dy, int dx, int use_base_off, int b_is_f16 = 0) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                      // 180 rows x 128 B
    uint8_t* sB = smem + 24 * 1024;          // 64 rows x 128 B
    uint64_t* bar = (uint64_t*)(smem + 33 * 1024);
    uint32_t* slot = (uint32_t*)(bar + 1);
    for (int i = threadIdx.x; i < HT_H * HT_W * 64; i += 128) {
        const int R = i / 64, k = i % 64;
        *(__nv_bfloat16*)(sA + R * 128 + (((k / 8) ^ (R % 8)) * 16) + (k % 8) * 2) = __float2bfloat16(aval(R, k));
    }
    for (int i = threadIdx.x; i < NCOL * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        if (b_is_f16) *(__half*)(sB + n * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2) = __float2half(bval(n, k) + 0.0009765625f * (float)(k & 3));
        else *(__nv_bfloat16*)(sB + n * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2) = __float2bfloat16(bval(n, k));
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(NCOL));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | ((b_is_f16 ? 0u : 1u) << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(sA) + (uint32_t)((dy * HT_W + dx) * 128);
        const uint32_t boff = use_base_off ? ((a0 >> 7) & 7) : 0;
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc(a0 + k * 32, HT_W * 128, boff);
            const uint64_t db = make_desc(smem_u32(sB) + k * 32, 1024, 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(k ? 1u : 0u)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < NCOL; c += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * NCOL + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NCOL));
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 128 * NCOL * sizeof(float));
    float* h = (float*)malloc(128 * NCOL * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    for (int ub = 0; ub < 2; ++ub) {
        int total_bad = 0;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                probe<<<1, 128, 36 * 1024>>>(d_out, dy, dx, ub);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("base_off=%d tap(%d,%d): CUDA error %s\n", ub, dy, dx, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d_out, 128 * NCOL * sizeof(float), cudaMemcpyDeviceToHost);
                int bad = 0, bad_g[16] = {0};
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < NCOL; ++n) {
                        const int R = (m / 8 + dy) * HT_W + (m % 8) + dx;
                        float ref = 0.f;
                        for (int k = 0; k < 64; ++k) ref += aval(R, k) * bval(n, k);
                        if (h[m * NCOL + n] != ref) { ++bad; ++bad_g[m / 8]; }
                    }
                printf("base_off=%d tap(%d,%d): %d mismatches; per 8-row group:", ub, dy, dx, bad);
                for (int g = 0; g < 16; ++g) printf(" %d", bad_g[g]);
                printf("\n");
                total_bad += bad;
            }
        printf("== base_off mode %d: total mismatches %d\n", ub, total_bad);
    }
    // mixed operand formats: A bf16, B fp16 (values with 2^-10 fractions that bf16 could not hold)
    {
        probe<<<1, 128, 36 * 1024>>>(d_out, 1, 1, 0, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mixed A=bf16/B=f16: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d_out, 128 * NCOL * sizeof(float), cudaMemcpyDeviceToHost);
        int bad = 0;
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < NCOL; ++n) {
                const int R = (m / 8 + 1) * HT_W + (m % 8) + 1;
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)aval(R, k) * ((double)bval(n, k) + 0.0009765625 * (k & 3));
                const double err = fabs((double)h[m * NCOL + n] - ref);
                if (err > maxerr) maxerr = err;
                if (err > 1e-3) ++bad;
            }
        printf("== mixed A=bf16 / B=f16: %d mismatches, max abs err %.3e\n", bad, maxerr);
    }
    return 0;
}
