// ubench.cu -- pipe-throughput probes for the kernel design (run under gpurun):
//   FFMA vs FFMA2/FADD2 lanes per clock per SM, PRMT rate, LDS.128 rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, long long *cyc, float s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 m = make_float2(s, s * 0.999f);
    unsigned p[8];
    for (int i = 0; i < 8; ++i) p[i] = threadIdx.x * 2654435761u + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, m.y); a[i].y = fmaf(a[i].y, m.x, m.y); }
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, m); }
            if (MODE == 2) { a[i] = __fadd2_rn(a[i], m); }
            if (MODE == 3) { a[i] = __ffma2_rn(a[i], make_float2(s, s), a[(i + 1) & 7]); }   // broadcast operand
            if (MODE == 4) { p[i] = __byte_perm(p[i], 0x47000000u, 0x7414); a[i].x = __uint_as_float(p[i]) + a[i].x; }
            if (MODE == 5) { a[i] = __ffma2_rn(a[i], m, m); p[i] = __byte_perm(p[i], p[(i + 1) & 7], 0x3120); }
            if (MODE == 6) { a[i].x = (float)(signed char)(p[i] >> 8); p[i] += (unsigned)it; }                 // I2F.S8 + IADD
            if (MODE == 7) { a[i] = __ffma2_rn(a[i], m, m); a[(i + 4) & 7].y += (float)(signed char)(p[i] >> 16); }  // FFMA2 + I2F.S8 + FADD
            if (MODE == 8) { a[i] = __ffma2_rn(a[i], m, make_float2((float)(signed char)(p[i] >> 16), (float)(signed char)(p[i] >> 8))); }  // FFMA2 + 2 I2F
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + (float)p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) lds128(float *out, long long *cyc) {
    __shared__ float4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float4 v = buf[(threadIdx.x + j * 256 + it) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 1024 * sizeof(long long));
    long long h[1024];
    const char *names[] = {"FFMA scalar (2/iter)", "FFMA2", "FADD2", "FFMA2 bcast operand", "PRMT+FADD", "FFMA2+PRMT", "I2F.S8+IADD", "FFMA2+I2F.S8+FADD", "FFMA2+2xI2F.S8"};
    const int ninstr[] = {2, 1, 1, 1, 2, 2, 2, 3, 3};
    for (int threads : {256, 512}) {
        for (int mode = 0; mode < 9; ++mode) {
            int grid = 148;
            switch (mode) {
                case 0: k<0><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 1: k<1><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 2: k<2><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 3: k<3><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 4: k<4><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 5: k<5><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 6: k<6><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 7: k<7><<<grid, threads>>>(out, cyc, 1.0001f); break;
                case 8: k<8><<<grid, threads>>>(out, cyc, 1.0001f); break;
            }
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double c = 0; for (int i = 0; i < grid; ++i) c += h[i]; c /= grid;
            double instr = (double)threads / 32 * ITERS * 8 * ninstr[mode];
            printf("threads=%d %-24s cycles=%.0f  warp-instr/clk/SM=%.3f\n", threads, names[mode], c, instr / c);
        }
    }
    lds128<<<148, 256>>>(out, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("LDS.128: %.1f B/clk/SM (256 thr)\n", 256.0 * 16 * 16 * 1024 / c);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
