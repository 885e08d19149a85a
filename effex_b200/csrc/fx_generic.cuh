// fx_generic.cuh -- shape-generic kernels: block sums, PFB FIR, batched FFT,
// X-engine, finalize/integrate, and the delay-calibration lag search.
// They serve every (ntaps <= 32, nbins = 2^k) the fused kernels do not cover,
// expose the pieces the reference's tests call on their own, and cross-check
// the fused path on the device.
#pragma once
#include "fx_common.cuh"
#include "fx_comm.cuh"

#ifndef FX_SUMS_DP4A
#define FX_SUMS_DP4A 0
#endif

#ifndef FX_FIN_CTAS
#define FX_FIN_CTAS 5      // resident CTAs per SM of the finalize+integrate kernels (register cap; the tail may spill)
#endif

namespace fx {
namespace generic {

// ---------------------------------------------------------------------------
// Exact per-block byte sums (for the DC removal of effex.py:394-395).
// sums[b][0] += sum of I bytes, sums[b][1] += sum of Q bytes (stride: sums + b*stride).
// grid = (chunks, n_blocks, channels).  HBM-bound streaming read.
// ---------------------------------------------------------------------------
__global__ void __maxnreg__(32) block_sums_kernel(const uint8_t *__restrict__ iq0,
                                                         const uint8_t *__restrict__ iq1, long long S,
                                                         unsigned long long *__restrict__ sums0, int stride) {
    const int b = blockIdx.y;
    const uint8_t *base = (blockIdx.z ? iq1 : iq0) + 2ll * S * b;
    unsigned long long *sums = sums0 + 2 * blockIdx.z;          // [block][channel][component]
    unsigned int si = 0, sq = 0;
    const long long nbytes = 2ll * S;
    const bool aligned = ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    const long long nvec = aligned ? nbytes / 16 : 0;
    const uint4 *v = reinterpret_cast<const uint4 *>(base);
    const long long step = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
#if FX_SUMS_DP4A
    auto add4 = [&](uint4 q) {
        si = __dp4a(q.x, 0x00010001u, si); sq = __dp4a(q.x, 0x01000100u, sq);
        si = __dp4a(q.y, 0x00010001u, si); sq = __dp4a(q.y, 0x01000100u, sq);
        si = __dp4a(q.z, 0x00010001u, si); sq = __dp4a(q.z, 0x01000100u, sq);
        si = __dp4a(q.w, 0x00010001u, si); sq = __dp4a(q.w, 0x01000100u, sq);
    };
#else
    // SIMD-within-a-register on the ALU pipe (LOP3/PRMT/IADD): I bytes and Q bytes of a word are added
    // as two 16-bit lanes; four words (<= 8 x 255) fit a lane, then the lanes are folded into si / sq
    auto add4 = [&](uint4 q) {
        const unsigned e = (q.x & 0x00ff00ffu) + (q.y & 0x00ff00ffu) + (q.z & 0x00ff00ffu) + (q.w & 0x00ff00ffu);
        const unsigned o = ((q.x >> 8) & 0x00ff00ffu) + ((q.y >> 8) & 0x00ff00ffu) + ((q.z >> 8) & 0x00ff00ffu) +
                           ((q.w >> 8) & 0x00ff00ffu);
        si += (e & 0xffffu) + (e >> 16);
        sq += (o & 0xffffu) + (o >> 16);
    };
#endif
    for (; i + 3 * step < nvec; i += 4 * step) {       // 4 independent 16-byte loads in flight
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = __ldg(v + i + u * step);
#pragma unroll
        for (int u = 0; u < 4; ++u) add4(q[u]);
    }
    for (; i < nvec; i += step) add4(__ldg(v + i));
    // tail (or everything, when the block is not 16-byte aligned): one sample per thread
    for (long long s = nvec * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S;
         s += (long long)gridDim.x * blockDim.x) {
        si += base[2 * s];
        sq += base[2 * s + 1];
    }
    unsigned long long wi = si, wq = sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wi += __shfl_xor_sync(0xffffffffu, wi, o);
        wq += __shfl_xor_sync(0xffffffffu, wq, o);
    }
    __shared__ unsigned long long red[2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = wi; red[1][warp] = wq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; c += red[1][w]; }
        atomicAdd(sums + (long long)b * stride, a);
        atomicAdd(sums + (long long)b * stride + 1, c);
    }
}

// ---------------------------------------------------------------------------
// PFB FIR, any T <= 32, any N.   w[b][i][p] = sum_{k<=min(i,T-1)} taps[k][p] * x[(i-k)N + p]
// taps[k][p] = h[kN + N-1-p] (already reversed; the u8 variant is pre-scaled by 1/127.5)
// grid = (ceil(N/256), ceil(P/kFirFrames), n_blocks): a thread produces kFirFrames consecutive frames of
// one branch p (independent loads of several frames in flight; means and index set-up paid once)
// ---------------------------------------------------------------------------
constexpr int kFirFrames = 8;
template <bool U8>
__global__ void __launch_bounds__(256) pfb_fir_kernel(const void *__restrict__ in, long long S, int N, int T, int P,
                                                      const float *__restrict__ taps,
                                                      const unsigned long long *__restrict__ sums, int sum_stride,
                                                      int dc_remove, float2 *__restrict__ w,
                                                      const void *__restrict__ halo = nullptr, long long mean_count = 0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * kFirFrames, b = blockIdx.z;
    // block means: float64 division once per CTA, not once per thread
    __shared__ float s_mean[2];
    if (U8 && threadIdx.x == 0) {
        if (dc_remove) {
            const double den = (double)(mean_count > 0 ? mean_count : S);
            s_mean[0] = (float)((double)sums[(long long)b * sum_stride] / den);
            s_mean[1] = (float)((double)sums[(long long)b * sum_stride + 1] / den);
        } else {
            s_mean[0] = s_mean[1] = 127.5f;
        }
    }
    if (U8) __syncthreads();
    if (p >= N) return;
    const float mi = U8 ? s_mean[0] : 0.f, mq = U8 ? s_mean[1] : 0.f;
    const bool have_halo = halo && b == 0;
#pragma unroll 4
    for (int i = i0; i < i0 + kFirFrames; ++i) {
        if (i >= P) break;
        float ar = 0.f, ai = 0.f;
        // zero history before frame 0 -- unless the caller supplied the T-1 preceding frames (streaming mode)
        const int kmax = (i < T - 1 && !have_halo) ? i : T - 1;
        for (int k = 0; k <= kmax; ++k) {
            long long s = (long long)b * S + (long long)(i - k) * N + p;
            const void *src = in;
            if (i - k < 0) { src = halo; s = (long long)(i - k + T - 1) * N + p; }
            float xr, xi;
            if (U8) {
                const uchar2 q = reinterpret_cast<const uchar2 *>(src)[s];
                xr = (float)q.x - mi;
                xi = (float)q.y - mq;
            } else {
                const float2 q = reinterpret_cast<const float2 *>(src)[s];
                xr = q.x;
                xi = q.y;
            }
            const float h = taps[(long long)k * N + p];
            ar = fmaf(h, xr, ar);
            ai = fmaf(h, xi, ai);
        }
        w[((long long)b * P + i) * N + p] = make_float2(ar, ai);
    }
}

// T = 4 (the reference's ntaps, effex.py:115): a thread walks kFir4Frames consecutive frames of one
// branch with the three previous samples in registers -- one load and 8 FMAs per output.
// grid = (ceil(N/256), ceil(P/kFir4Frames), n_blocks)
constexpr int kFir4Frames = 32;
template <bool U8>
__global__ void __launch_bounds__(256) pfb_fir4_kernel(const void *__restrict__ in, long long S, int N, int P,
                                                       const float *__restrict__ taps,
                                                       const unsigned long long *__restrict__ sums, int sum_stride,
                                                       int dc_remove, float2 *__restrict__ w,
                                                       const void *__restrict__ halo = nullptr, long long mean_count = 0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * kFir4Frames, b = blockIdx.z;
    __shared__ float s_mean[2];
    if (U8 && threadIdx.x == 0) {
        if (dc_remove) {
            const double den = (double)(mean_count > 0 ? mean_count : S);
            s_mean[0] = (float)((double)sums[(long long)b * sum_stride] / den);
            s_mean[1] = (float)((double)sums[(long long)b * sum_stride + 1] / den);
        } else {
            s_mean[0] = s_mean[1] = 127.5f;
        }
    }
    if (U8) __syncthreads();
    if (p >= N) return;
    const float mi = U8 ? s_mean[0] : 0.f, mq = U8 ? s_mean[1] : 0.f;
    const bool have_halo = halo && b == 0;
    // sample of frame i (block-relative; i < 0 reads the halo or is zero history)
    auto sample = [&](int i) -> float2 {
        const void *src = in;
        long long s = (long long)b * S + (long long)i * N + p;
        if (i < 0) {
            if (!have_halo) return make_float2(0.f, 0.f);
            src = halo;
            s = (long long)(i + 3) * N + p;
        }
        if (U8) {
            const uchar2 q = reinterpret_cast<const uchar2 *>(src)[s];
            return make_float2((float)q.x - mi, (float)q.y - mq);
        }
        return reinterpret_cast<const float2 *>(src)[s];
    };
    const float t0 = taps[p], t1 = taps[(long long)N + p], t2 = taps[2ll * N + p], t3 = taps[3ll * N + p];
    float2 x1 = sample(i0 - 1), x2 = sample(i0 - 2), x3 = sample(i0 - 3);
    const int i1 = min(i0 + kFir4Frames, P);
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
        const float2 x0 = sample(i);
        w[((long long)b * P + i) * N + p] =
            make_float2(fmaf(t0, x0.x, fmaf(t1, x1.x, fmaf(t2, x2.x, t3 * x3.x))),
                        fmaf(t0, x0.y, fmaf(t1, x1.y, fmaf(t2, x2.y, t3 * x3.y))));
        x3 = x2; x2 = x1; x1 = x0;
    }
}

// ---------------------------------------------------------------------------
// Batched forward/inverse FFT of length N = 2^logN <= 4096, one row per CTA,
// radix-4 Stockham autosort stages in shared memory (twiddle table built once per row).  phase_post: multiply bin c by
// exp(-2*pi*i*c/N) (SURVEY App. A.4 factor, so rows equal channelize_poly's).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(512) fft_rows_kernel(const float2 *__restrict__ in, float2 *__restrict__ out,
                                                       int N, int logN, int inverse, int phase_post) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *bufA = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufB = bufA + N;
    float2 *tw = bufB + N;                       // W_N^i, i < N
    const long long row = blockIdx.x;
    const float2 *src = in + row * N;
    const float sgn = inverse ? 1.f : -1.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        bufA[j] = src[j];
        float sn, cs;
        sincospif(sgn * 2.f * (float)j / (float)N, &sn, &cs);
        tw[j] = make_float2(cs, sn);
    }
    __syncthreads();
    int ns = 1;
    if (logN & 1) {                              // one radix-2 stage (ns = 1: no twiddles)
        const int half = N >> 1;
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            const float2 v0 = bufA[j], v1 = bufA[j + half];
            bufB[2 * j] = make_float2(v0.x + v1.x, v0.y + v1.y);
            bufB[2 * j + 1] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        float2 *t = bufA; bufA = bufB; bufB = t;
        ns = 2;
    }
    const int quarter = N >> 2;
    for (; ns < N; ns <<= 2) {                   // radix-4 autosort stages
        const int step = N / (4 * ns);
        for (int b = threadIdx.x; b < quarter; b += blockDim.x) {
            const int k = b & (ns - 1);
            const float2 a0 = bufA[b];
            float2 a1 = bufA[b + quarter], a2 = bufA[b + 2 * quarter], a3 = bufA[b + 3 * quarter];
            if (k) {
                const float2 w1 = tw[k * step], w2 = tw[2 * k * step], w3 = tw[3 * k * step];
                a1 = make_float2(a1.x * w1.x - a1.y * w1.y, a1.x * w1.y + a1.y * w1.x);
                a2 = make_float2(a2.x * w2.x - a2.y * w2.y, a2.x * w2.y + a2.y * w2.x);
                a3 = make_float2(a3.x * w3.x - a3.y * w3.y, a3.x * w3.y + a3.y * w3.x);
            }
            const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
            const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
            const float2 rot = make_float2(-sgn * d13.y, sgn * d13.x);      // sgn*i*d13  (forward: W4 = -i)
            const int o = ((b - k) << 2) + k;
            bufB[o] = make_float2(s02.x + s13.x, s02.y + s13.y);
            bufB[o + ns] = make_float2(d02.x + rot.x, d02.y + rot.y);
            bufB[o + 2 * ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
            bufB[o + 3 * ns] = make_float2(d02.x - rot.x, d02.y - rot.y);
        }
        __syncthreads();
        float2 *t = bufA; bufA = bufB; bufB = t;
    }
    float2 *dst = out + row * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float2 v = bufA[j];
        if (phase_post) {                        // exp(-2*pi*i*j/N) = forward W_N^j
            const float2 w = tw[j];
            const float wy = inverse ? -w.y : w.y;
            v = make_float2(v.x * w.x - v.y * wy, v.x * wy + v.y * w.x);
        }
        dst[j] = v;
    }
}

// One radix-R Stockham autosort pass over global memory, R = 2^logR <= 256 (any length M = 2^m, batch
// rows of length M).  For j in [0, M/R), k = j mod Ns:
//     out[(j-k)*R + k + c*Ns] = sum_r W_R^(r*c) * W_(R*Ns)^(r*k) * in[j + r*M/R]
// A CTA takes J consecutive j (R*J points): loads them with the inter-pass twiddle applied, runs the
// R-point transforms of its J columns as radix-2 autosort stages in shared memory, and stores.  A
// transform of 2^16 points is 2 passes over HBM, of 2^19 points 3 (a radix-2 pass per stage was 16 / 19).
// Used for transforms that do not fit one CTA: the 2n-point lag-search FFTs and N > 4096 channelizers.
// grid = (M/R/J, batch), dynamic smem = (2*R*(J+1) + R)*sizeof(float2)
constexpr int kPassJ = 16;
__global__ void __launch_bounds__(512) stockham_radix_pass_kernel(const float2 *__restrict__ in,
                                                                  float2 *__restrict__ out, long long M,
                                                                  long long Ns, int logR, int inverse) {
    constexpr int J = kPassJ, JP = J + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = 1 << logR;
    float2 *A = reinterpret_cast<float2 *>(smem_raw);
    float2 *B = A + R * JP;
    const long long stride = M >> logR;
    const long long j0 = (long long)blockIdx.x * J;
    const float2 *src = in + (long long)blockIdx.y * M;
    float2 *dst = out + (long long)blockIdx.y * M;
    const float sgn = inverse ? 1.f : -1.f;
    const float inv_span = 1.f / (float)((long long)R * Ns);       // power of two: r*k*inv_span is exact
    for (int e = threadIdx.x; e < R * J; e += blockDim.x) {
        const int r = e / J, jj = e % J;
        const long long j = j0 + jj;
        float2 v = src[j + r * stride];
        const long long k = j & (Ns - 1);
        if (k != 0 && r != 0) {
            float sn, cs;
            sincospif(sgn * 2.f * (float)(r * k) * inv_span, &sn, &cs);
            v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
        }
        A[r * JP + jj] = v;
    }
    // W_R^i, i < R, once per CTA: a radix-4 stage with sub-length ns uses W_(4 ns)^(q k) = W_R^(q k R/(4 ns))
    float2 *tw = B + R * JP;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        float sn, cs;
        sincospif(sgn * 2.f * (float)i / (float)R, &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    __syncthreads();
    int ns = 1;
    if (logR & 1) {
        // one radix-2 stage (ns = 1: no twiddles)
        const int half = R >> 1;
        for (int e = threadIdx.x; e < half * J; e += blockDim.x) {
            const int b = e / J, jj = e % J;
            const float2 v0 = A[b * JP + jj], v1 = A[(b + half) * JP + jj];
            B[(2 * b) * JP + jj] = make_float2(v0.x + v1.x, v0.y + v1.y);
            B[(2 * b + 1) * JP + jj] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        float2 *t = A; A = B; B = t;
        ns = 2;
    }
    const int quarter = R >> 2;
    for (; ns < R; ns <<= 2) {
        const int step = R / (4 * ns);
        for (int e = threadIdx.x; e < quarter * J; e += blockDim.x) {
            const int b = e / J, jj = e % J;
            const int k = b & (ns - 1);
            const float2 a0 = A[b * JP + jj];
            float2 a1 = A[(b + quarter) * JP + jj], a2 = A[(b + 2 * quarter) * JP + jj], a3 = A[(b + 3 * quarter) * JP + jj];
            if (k) {
                const float2 w1 = tw[k * step], w2 = tw[2 * k * step], w3 = tw[3 * k * step];
                a1 = make_float2(a1.x * w1.x - a1.y * w1.y, a1.x * w1.y + a1.y * w1.x);
                a2 = make_float2(a2.x * w2.x - a2.y * w2.y, a2.x * w2.y + a2.y * w2.x);
                a3 = make_float2(a3.x * w3.x - a3.y * w3.y, a3.x * w3.y + a3.y * w3.x);
            }
            const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
            const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
            // forward: W4 = -i; inverse: +i   (sgn = -1 forward)
            const float2 rot = make_float2(-sgn * d13.y, sgn * d13.x);      // sgn*i*d13
            const int o = ((b - k) << 2) + k;
            B[o * JP + jj] = make_float2(s02.x + s13.x, s02.y + s13.y);
            B[(o + ns) * JP + jj] = make_float2(d02.x + rot.x, d02.y + rot.y);
            B[(o + 2 * ns) * JP + jj] = make_float2(s02.x - s13.x, s02.y - s13.y);
            B[(o + 3 * ns) * JP + jj] = make_float2(d02.x - rot.x, d02.y - rot.y);
        }
        __syncthreads();
        float2 *t = A; A = B; B = t;
    }
    if (Ns < J) {
        // first pass: the R outputs of one j are contiguous
        for (int e = threadIdx.x; e < R * J; e += blockDim.x) {
            const int jj = e / R, c = e % R;
            const long long j = j0 + jj, k = j & (Ns - 1);
            dst[(j - k) * R + k + c * Ns] = A[c * JP + jj];
        }
    } else {
        // later passes: for one output digit c the J columns are contiguous
        for (int e = threadIdx.x; e < R * J; e += blockDim.x) {
            const int c = e / J, jj = e % J;
            const long long j = j0 + jj, k = j & (Ns - 1);
            dst[(j - k) * R + k + c * Ns] = A[c * JP + jj];
        }
    }
}

// ---------------------------------------------------------------------------
// X-engine over stored spectra: part_x[b][c] = sum_i F0*conj(F1), part_a = (sum|F0|^2, sum|F1|^2)
// grid = (ceil(N/256), n_blocks)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xengine_kernel(const float2 *__restrict__ F0, const float2 *__restrict__ F1,
                                                      int N, int P, float2 *__restrict__ part_x,
                                                      float2 *__restrict__ part_a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (c >= N) return;
    float xr = 0.f, xi = 0.f, a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < P; ++i) {
        const float2 f0 = F0[((long long)b * P + i) * N + c];
        const float2 f1 = F1[((long long)b * P + i) * N + c];
        xr = fmaf(f0.x, f1.x, fmaf(f0.y, f1.y, xr));
        xi = fmaf(f0.y, f1.x, fmaf(-f0.x, f1.y, xi));
        a0 = fmaf(f0.x, f0.x, fmaf(f0.y, f0.y, a0));
        a1 = fmaf(f1.x, f1.x, fmaf(f1.y, f1.y, a1));
    }
    part_x[(long long)b * N + c] = make_float2(xr, xi);
    part_a[(long long)b * N + c] = make_float2(a0, a1);
}

// np.fft.fftshift for any N: out[j] = X[(j + ceil(N/2)) mod N]
__device__ __forceinline__ int shifted_bin(int j, int N) {
    const int c = j + ((N + 1) >> 1);
    return c >= N ? c - N : c;
}

// ---------------------------------------------------------------------------
// Bluestein: the N-point DFT of any N as a circular convolution of length M = 2^m >= 2N - 1
// (the reference takes any --resolution through cuFFT, effex.py:553, :734):
//   X[k] = c[k] * sum_n (x[n] c[n]) * conj(c)[k - n],   c[n] = exp(-i pi n^2 / N)
// pre: a[r][n] = x[r][n] * c[n] (n < N), 0 (N <= n < M); then FFT_M, times B = FFT_M(conj chirp, wrapped),
// inverse FFT_M, post: X[r][k] = c[k] * a[r][k] / M (times exp(-2 pi i k / N) when phase_post).
// The chirp and B are built on the host in float64 (n^2 mod 2N in integers).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bluestein_pre_kernel(const float2 *__restrict__ x, const float2 *__restrict__ chirp,
                                                            int N, int M, float2 *__restrict__ a) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = blockIdx.y;
    if (m >= M) return;
    float2 v = make_float2(0.f, 0.f);
    if (m < N) {
        const float2 q = x[r * N + m], c = chirp[m];
        v = make_float2(q.x * c.x - q.y * c.y, q.x * c.y + q.y * c.x);
    }
    a[r * M + m] = v;
}
__global__ void __launch_bounds__(256) bluestein_mul_kernel(float2 *__restrict__ a, const float2 *__restrict__ B, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = blockIdx.y;
    if (m >= M) return;
    const float2 q = a[r * M + m], b = B[m];
    a[r * M + m] = make_float2(q.x * b.x - q.y * b.y, q.x * b.y + q.y * b.x);
}
__global__ void __launch_bounds__(256) bluestein_post_kernel(const float2 *__restrict__ a, const float2 *__restrict__ chirp,
                                                             int N, int M, int phase_post, float2 *__restrict__ x) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = blockIdx.y;
    if (k >= N) return;
    const float inv = 1.0f / (float)M;
    const float2 q = a[r * M + k], c = chirp[k];
    float2 v = make_float2((q.x * c.x - q.y * c.y) * inv, (q.x * c.y + q.y * c.x) * inv);
    if (phase_post) {
        float sn, cs;
        sincospif(-2.f * (float)k / (float)N, &sn, &cs);
        v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    }
    x[r * N + k] = v;
}

// ---------------------------------------------------------------------------
// finalize: rows in the reference's output order (effex.py:519-521):
//   out[b][j] = conj(rot[c]) * (1/P) * sum_{s in segments of block b} part_x[s][c],  c = (j + N/2) mod N
// blk_first[b] .. blk_first[b+1] are block b's segments (NULL: one segment per block, s = b).
// grid = (ceil(N/256), n_blocks)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_rows_kernel(const float2 *__restrict__ part_x,
                                                            const float2 *__restrict__ part_a, int N,
                                                            const int *__restrict__ blk_first, int block0,
                                                            float inv_frames, const float2 *__restrict__ rot,
                                                            float2 *__restrict__ xspec, float *__restrict__ auto0,
                                                            float *__restrict__ auto1) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = block0 + blockIdx.y;
    if (j >= N) return;
    const int c = shifted_bin(j, N);
    const int s0 = blk_first ? blk_first[b] : b;
    const int s1 = blk_first ? blk_first[b + 1] : b + 1;
    float xr = 0.f, xi = 0.f, a0 = 0.f, a1 = 0.f;
    const bool autos = auto0 || auto1;            // part_a is not written when nobody asked for auto-powers
    for (int s = s0; s < s1; ++s) {
        const long long o = (long long)s * N + c;
        const float2 x = part_x[o];
        xr += x.x; xi += x.y;
        if (autos) {
            const float2 a = part_a[o];
            a0 += a.x; a1 += a.y;
        }
    }
    xr *= inv_frames; xi *= inv_frames;
    const float2 r = rot ? rot[c] : make_float2(1.f, 0.f);
    // X * conj(rot)
    xspec[(long long)b * N + j] = make_float2(xr * r.x + xi * r.y, xi * r.x - xr * r.y);
    if (auto0) auto0[(long long)b * N + j] = a0 * inv_frames;
    if (auto1) auto1[(long long)b * N + j] = a1 * inv_frames;
}

// The same for N % 4 == 0 with two bins per thread: 16-byte loads and stores, half the threads (the kernel is a
// short latency-bound tail of every step: 15.7 -> ~10 us for 550 rows of 4096 bins).   grid = (ceil(N/512), n_blocks)
__global__ void __launch_bounds__(256) finalize_rows2_kernel(const float2 *__restrict__ part_x,
                                                             const float2 *__restrict__ part_a, int N,
                                                             const int *__restrict__ blk_first, int block0,
                                                             float inv_frames, const float2 *__restrict__ rot,
                                                             float2 *__restrict__ xspec, float *__restrict__ auto0,
                                                             float *__restrict__ auto1) {
    const int j = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int b = block0 + blockIdx.y;
    if (j >= N) return;
    const int c = shifted_bin(j, N);               // even, and c + 1 is the shifted bin of j + 1
    const int s0 = blk_first ? blk_first[b] : b;
    const int s1 = blk_first ? blk_first[b + 1] : b + 1;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), a = x;
    const bool autos = auto0 || auto1;
    for (int s = s0; s < s1; ++s) {
        const long long o = (long long)s * N + c;
        const float4 px = *reinterpret_cast<const float4 *>(part_x + o);
        x.x += px.x; x.y += px.y; x.z += px.z; x.w += px.w;
        if (autos) {
            const float4 pa = *reinterpret_cast<const float4 *>(part_a + o);
            a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
        }
    }
    x.x *= inv_frames; x.y *= inv_frames; x.z *= inv_frames; x.w *= inv_frames;
    const float4 r = rot ? *reinterpret_cast<const float4 *>(rot + c) : make_float4(1.f, 0.f, 1.f, 0.f);
    // X * conj(rot)
    *reinterpret_cast<float4 *>(xspec + (long long)b * N + j) =
        make_float4(x.x * r.x + x.y * r.y, x.y * r.x - x.x * r.y, x.z * r.z + x.w * r.w, x.w * r.z - x.z * r.w);
    if (auto0) *reinterpret_cast<float2 *>(auto0 + (long long)b * N + j) = make_float2(a.x * inv_frames, a.z * inv_frames);
    if (auto1) *reinterpret_cast<float2 *>(auto1 + (long long)b * N + j) = make_float2(a.y * inv_frames, a.w * inv_frames);
}

// finalize + integrate in one pass over the partial sums: thread (j, g) walks the blocks of group g,
// writes each block's row like finalize_rows_kernel and adds the block's un-normalised sums (float64)
// into scratch[g] in natural bin order.  The G groups are folded either by integrate_stage2_kernel
// afterwards, or -- use_tail -- by the last CTA of each bin tile to finish (comm::integrate_tail), which
// also delivers them: into the caller's accumulators, or into the reduce root's mailbox.
// grid = (ceil(N/256), G)
template <bool AUTOS>
__global__ void __launch_bounds__(256, FX_FIN_CTAS) finalize_integrate_kernel(const float2 *__restrict__ part_x,
                                                                 const float2 *__restrict__ part_a, int N,
                                                                 const int *__restrict__ blk_first, int n_blocks,
                                                                 float inv_frames, const float2 *__restrict__ rot,
                                                                 float2 *__restrict__ xspec, float *__restrict__ auto0,
                                                                 float *__restrict__ auto1, double *__restrict__ scratch,
                                                                 int use_tail, double frames,
                                                                 const comm::IntegrateTail tail) {
    constexpr bool autos = AUTOS;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = j < N;
    const int c = shifted_bin(j, N);
    const int G = gridDim.y, g = blockIdx.y;
    if (valid) {
        const int b0 = (int)((long long)n_blocks * g / G), b1 = (int)((long long)n_blocks * (g + 1) / G);
        const float2 r = rot ? rot[c] : make_float2(1.f, 0.f);
        double dxr = 0, dxi = 0, da0 = 0, da1 = 0;
        // one flat walk over the group's segments, four loads in flight; a row is emitted at every block boundary
        const int sA = blk_first ? blk_first[b0] : b0, sB = blk_first ? blk_first[b1] : b1;
        int b = b0;
        int next = b0 < b1 ? (blk_first ? blk_first[b0 + 1] : b0 + 1) : sB;
        float xr = 0.f, xi = 0.f, a0 = 0.f, a1 = 0.f;
        for (int s4 = sA; s4 < sB; s4 += 4) {
            float2 x[4], a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool in = s4 + u < sB;
                const long long o = (long long)(in ? s4 + u : sA) * N + c;
                x[u] = part_x[o];
                a[u] = autos ? part_a[o] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int sg = s4 + u;
                if (sg >= sB) break;
                xr += x[u].x; xi += x[u].y;
                if (autos) { a0 += a[u].x; a1 += a[u].y; }
                if (sg + 1 == next) {
                    dxr += xr; dxi += xi;
                    if (autos) { da0 += a0; da1 += a1; }
                    xr *= inv_frames; xi *= inv_frames;
                    xspec[(long long)b * N + j] = make_float2(xr * r.x + xi * r.y, xi * r.x - xr * r.y);
                    if (autos && auto0) auto0[(long long)b * N + j] = a0 * inv_frames;
                    if (autos && auto1) auto1[(long long)b * N + j] = a1 * inv_frames;
                    xr = xi = a0 = a1 = 0.f;
                    ++b;
                    next = b < b1 ? (blk_first ? blk_first[b + 1] : b + 1) : sB;
                }
            }
        }
        double *o = scratch + (long long)g * 4 * N;
        *reinterpret_cast<double2 *>(o + 2 * c) = make_double2(dxr, dxi);
        o[2 * N + c] = da0;
        o[3 * N + c] = da1;
    }
    if (use_tail) comm::integrate_tail(tail, scratch, N, G, frames, c >> 8, c, valid);
}

// integrate: float64 accumulators += sum over all segments of the call (natural order, no rot).
// Two deterministic stages: grid (ceil(N/256), G) partial sums over segment slices into scratch[G][4N],
// then the fold of the G slices (integrate_stage2_kernel, or the tail as above).
__global__ void __launch_bounds__(256, FX_FIN_CTAS) integrate_stage1_kernel(const float2 *__restrict__ part_x,
                                                               const float2 *__restrict__ part_a, int N, int n_segs,
                                                               double *__restrict__ scratch, int autos, int use_tail,
                                                               double frames, const comm::IntegrateTail tail) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = c < N;
    const int G = gridDim.y, g = blockIdx.y;
    if (valid) {
        const int s0 = (int)((long long)n_segs * g / G), s1 = (int)((long long)n_segs * (g + 1) / G);
        double xr = 0, xi = 0, a0 = 0, a1 = 0;
        for (int s = s0; s < s1; ++s) {
            const float2 x = part_x[(long long)s * N + c];
            xr += x.x; xi += x.y;
            if (autos) {
                const float2 a = part_a[(long long)s * N + c];
                a0 += a.x; a1 += a.y;
            }
        }
        double *o = scratch + (long long)g * 4 * N;
        o[2 * c] = xr;
        o[2 * c + 1] = xi;
        o[2 * N + c] = a0;
        o[3 * N + c] = a1;
    }
    if (use_tail) comm::integrate_tail(tail, scratch, N, G, frames, c >> 8, c, valid);
}
__global__ void __launch_bounds__(256) integrate_stage2_kernel(const double *__restrict__ scratch, int N, int G,
                                                               double frames, double *__restrict__ acc_x,
                                                               double *__restrict__ acc_a0, double *__restrict__ acc_a1,
                                                               double *__restrict__ acc_frames) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // index into the 4N-element slice
    if (i >= 4 * N) return;
    double v = 0;
    for (int g = 0; g < G; ++g) v += scratch[(long long)g * 4 * N + i];
    if (i < 2 * N) acc_x[i] += v;
    else if (i < 3 * N) acc_a0[i - 2 * N] += v;
    else acc_a1[i - 3 * N] += v;
    if (i == 0 && acc_frames) *acc_frames += frames;
}

// ---------------------------------------------------------------------------
// Lag search (effex.py:583-622)
// ---------------------------------------------------------------------------
// zero-padded load of one block pair into two rows of length M (row 0 = ch0, row 1 = ch1)
template <bool U8>
__global__ void __launch_bounds__(256) lag_load_kernel(const void *__restrict__ in0, const void *__restrict__ in1,
                                                       long long n, long long M, long long block,
                                                       const unsigned long long *__restrict__ sums, int dc_remove,
                                                       float2 *__restrict__ rows) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= M) return;
    const int ch = blockIdx.y;
    float2 v = make_float2(0.f, 0.f);
    if (s < n) {
        if (U8) {
            const uint8_t *base = reinterpret_cast<const uint8_t *>(ch ? in1 : in0) + 2ll * n * block;
            float mi = 127.5f, mq = 127.5f;
            if (dc_remove) {
                mi = (float)((double)sums[4 * block + 2 * ch] / (double)n);
                mq = (float)((double)sums[4 * block + 2 * ch + 1] / (double)n);
            }
            const uchar2 q = reinterpret_cast<const uchar2 *>(base)[s];
            v = make_float2(((float)q.x - mi) * (1.0f / 127.5f), ((float)q.y - mq) * (1.0f / 127.5f));
        } else {
            const float2 *base = reinterpret_cast<const float2 *>(ch ? in1 : in0) + n * block;
            v = base[s];
        }
    }
    rows[(long long)ch * M + s] = v;
}

// acc[k] (+)= A[k]*conj(B[k])
__global__ void __launch_bounds__(256) lag_accum_kernel(const float2 *__restrict__ rows, long long M, int first,
                                                        float2 *__restrict__ acc) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= M) return;
    const float2 a = rows[k], b = rows[M + k];
    float2 r = make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
    if (!first) { const float2 o = acc[k]; r.x += o.x; r.y += o.y; }
    acc[k] = r;
}

struct ArgMax {
    float val;
    long long idx;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
    // larger value wins; on ties the smaller index (numpy argmax = first maximum)
    if (b.val > a.val || (b.val == a.val && b.idx < a.idx)) return b;
    return a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.val = __shfl_xor_sync(0xffffffffu, a.val, o);
        b.idx = __shfl_xor_sync(0xffffffffu, a.idx, o);
        a = better(a, b);
    }
    return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a) {
    __shared__ float sval[32];
    __shared__ long long sidx[32];
    a = warp_argmax(a);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sval[warp] = a.val; sidx[warp] = a.idx; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        ArgMax b;
        b.val = lane < nw ? sval[lane] : -1.f;
        b.idx = lane < nw ? sidx[lane] : (1ll << 62);
        a = warp_argmax(b);
    }
    return a;   // valid in warp 0
}
// xc_shift[j] = xc[(j - n) mod M], j in [0, 2n): linear lag j - n.
__global__ void __launch_bounds__(256) lag_argmax_stage1(const float2 *__restrict__ xc, long long n, long long M,
                                                         float *__restrict__ pval, long long *__restrict__ pidx) {
    ArgMax best{-1.f, 1ll << 62};
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < 2 * n;
         j += (long long)gridDim.x * blockDim.x) {
        const long long l = j - n;
        const float2 v = xc[l >= 0 ? l : l + M];
        best = better(best, ArgMax{v.x * v.x + v.y * v.y, j});
    }
    best = block_argmax(best);
    if (threadIdx.x == 0) { pval[blockIdx.x] = best.val; pidx[blockIdx.x] = best.idx; }
}
// out: idx[0] = imax ; nb[0..2] = |xc_shift[imax-1]|, |..[imax]|, |..[imax+1]| * scale (-1 = out of range)
__global__ void __launch_bounds__(256) lag_argmax_stage2(const float2 *__restrict__ xc, long long n, long long M,
                                                         const float *__restrict__ pval,
                                                         const long long *__restrict__ pidx, int nparts, float scale,
                                                         long long *__restrict__ out_idx, float *__restrict__ out_nb) {
    ArgMax best{-1.f, 1ll << 62};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) best = better(best, ArgMax{pval[i], pidx[i]});
    best = block_argmax(best);
    if (threadIdx.x == 0) {
        const long long imax = best.idx;
        out_idx[0] = imax;
        for (int d = -1; d <= 1; ++d) {
            long long j = imax + d;
            float r = -1.f;
            if (j < 0) j += 2 * n;            // python negative index wraps (xcorr[-1])
            if (j < 2 * n) {
                const long long l = j - n;
                const float2 v = xc[l >= 0 ? l : l + M];
                r = sqrtf(v.x * v.x + v.y * v.y) * scale;
            }
            out_nb[d + 1] = r;
        }
    }
}

}  // namespace generic
}  // namespace fx
