// ubench_tmem.cu -- tcgen05.ld / tcgen05.st throughput for thread-private TMEM scratch (32x32b shape).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld.x4, 1: ld.x16, 2: st.x4, 3: ld.x4 + st.x4, 4: ld.x32
__global__ void __launch_bounds__(256, 1) k(float *out, long long *cyc, int iters) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t ta = base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    // initialise 256 columns
    for (int c = 0; c < 256; c += 4)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 256; c += 64) {
            if (MODE == 0 || MODE == 3) {
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[(4 * u) & 31]), "=r"(r[(4 * u + 1) & 31]), "=r"(r[(4 * u + 2) & 31]), "=r"(r[(4 * u + 3) & 31]) : "r"(ta + c + 4 * u));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (MODE == 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta + c + 16 * u));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (MODE == 2 || MODE == 3) {
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta + c + 4 * u), "r"(r[(4 * u) & 31] + it), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            acc += r[0] + r[5] + r[17];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "r"(512) : "memory");
}

int main() {
    float *out; long long *cyc; long long h[148];
    cudaMalloc(&out, 148 * 256 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    const char *names[] = {"ld.x4", "ld.x16", "st.x4", "ld.x4+st.x4"};
    for (int mode = 0; mode < 4; ++mode) {
        switch (mode) {
            case 0: k<0><<<148, 256>>>(out, cyc, iters); break;
            case 1: k<1><<<148, 256>>>(out, cyc, iters); break;
            case 2: k<2><<<148, 256>>>(out, cyc, iters); break;
            case 3: k<3><<<148, 256>>>(out, cyc, iters); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        double bytes = 256.0 * 256 * 4 * iters * (mode == 3 ? 2 : 1);   // 256 thr x 256 cols x 4 B per iteration
        printf("%-14s %8.1f B/clk/SM  (%.0f cycles) %s\n", names[mode], bytes / c, c, cudaGetErrorString(e));
    }
    return 0;
}
