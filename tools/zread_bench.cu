// Read-pattern probe for bigfft::tail_kernel's Z loads (run under gpurun: nvcc -O3 -arch=sm_100a).
// 148 persistent CTAs x 256 threads, one CTA per SM (170 KB of dynamic shared memory requested, like the tail
// kernel); per frame a CTA reads the 64 KB row Z[frame][k1][0..4096) as 16 streaming 16-byte loads per thread
// (t + 256 r), rows 1 MB apart, optionally with the tail kernel's bulk L2 prefetch of the next row.
// mode 0: loads consumed at once; mode 1: + L2 prefetch one frame ahead; mode 2: + ~2300 cycles of dependent FMAs
// per frame between the loads (the tail kernel's FMA-pipe time per frame) to see what overlaps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) zr(const float4 *z, int frames_per_cta, int mode, float *sink) {
    extern __shared__ unsigned char sm[];
    const int cta = blockIdx.x, t = threadIdx.x;
    const int k1 = cta & 15, grp = cta >> 4;
    float acc = 0.f;
    for (int f = 0; f < frames_per_cta; ++f) {
        const float4 *zf = z + (((long long)grp * frames_per_cta + f) * 16 + k1) * 4096;
        if (mode >= 1 && t == 0 && f + 1 < frames_per_cta)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(zf + 65536), "r"(65536) : "memory");
        float4 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = __ldcs(zf + t + 256 * r);
#pragma unroll
        for (int r = 0; r < 16; ++r) acc += v[r].x + v[r].y + v[r].z + v[r].w;
        if (mode == 2) {
            float a = acc, b = acc * 0.5f;
#pragma unroll 1
            for (int i = 0; i < 290; ++i) {          // 8 dependent-pair FMAs x 290 ~ 2300 issue cycles per warp pair
                a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, a); a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, a);
                a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, a); a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, a);
            }
            acc = a + b;
        }
    }
    if (acc == 1.2345e30f) sink[0] = acc + sm[0];
}
int main() {
    const int frames_per_cta = 56, groups = 10;
    const size_t zbytes = (size_t)groups * frames_per_cta * 65536 * 16;
    float4 *z; float *sink;
    cudaMalloc(&z, zbytes); cudaMalloc(&sink, 4);
    cudaMemset(z, 0, zbytes);
    cudaFuncSetAttribute(zr, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(a);
            zr<<<148, 256, 170 * 1024>>>(z, frames_per_cta, mode, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
        }
        const double rb = 148.0 * frames_per_cta * 65536;
        printf("mode %d: %7.1f us  read %6.0f GB/s (%.0f MB)\n", mode, best * 1e3, rb / best / 1e6, rb / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
