// fx_lag.cuh -- the delay-calibration lag search (effex.py:583-622) on the fused kernel's FFT machinery.
//
// The reference zero-pads both channels of a block to 2n, takes three cuFFT Z2Z transforms of 2n points and
// an argmax.  Here the M-point transforms (M = 2^ceil(log2 2n) = G * 4096, G = 2 .. 256) of BOTH channels run
// together in the two lanes of the packed-FP32 registers, split once like the 8192..65536-bin channelizer
// (decimation in frequency, n = n1*4096 + n2):
//
//   lag_head_kernel   raw bytes (or complex64) of a block pair -> zero padding, DC removal, G-point DFT over n1
//                     in shared memory, twiddle W_M^(n2*k1) -> Z[block][k1][n2]  (16 B: both channels, re/im)
//   tail_kernel       (fx_bigfft.cuh) 4096-point transforms over n2 and the X-engine: with the BLOCKS of the
//                     call in the role of frames, its register accumulators ARE sum_b A_b * conj(B_b)
//   lag_fold_kernel   segments -> d_xacc[k1 + G*k2]
//
// so a 92-block accumulation (BASELINE config 2) is three launches, not 92 x 8.  The inverse transform reuses
// the pair: IFFT(x) = conj(FFT(conj x))/M, only |xc| is needed (effex.py:618-622), and with ONE frame the tail
// kernel's auto-power accumulator is |FFT(conj x)|^2 -- the argmax kernels read it in place.
#pragma once
#include "fx_bigfft.cuh"

namespace fx {
namespace lag {

using fused4096::N;        // 4096

// packed complex (re ch0, re ch1, im ch0, im ch1) times a scalar twiddle (wr + i wi)
__device__ __forceinline__ float4 tw_mul(float4 v, float wr, float wi) {
    return make_float4(v.x * wr - v.z * wi, v.y * wr - v.w * wi, v.x * wi + v.z * wr, v.y * wi + v.w * wr);
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
// v * (-i) (forward) : (re, im) -> (im, -re)
__device__ __forceinline__ float4 mul_mi(float4 v) { return make_float4(v.z, v.w, -v.x, -v.y); }

// One CTA: a tile of tn2 consecutive n2 (tn2 a multiple of 4) for all G values of n1 of one block pair.
// grid = (4096 / tn2, blocks of the chunk), 256 threads, dynamic smem = (2 * G * tn2) float4 + G float2.
// U8: in0/in1 = raw bytes [blocks][2n], DC removal from sums[4*block + ...]; else complex64 [blocks][n]
// (in1 may be NULL: channel 1 = 0).  conj_in conjugates the input (inverse transform through a forward one).
// twh = W_M^(n2*k1) as [G][4096] (host table, float64 -> float32).
template <bool U8>
__global__ void __launch_bounds__(256) lag_head_kernel(const void *__restrict__ in0, const void *__restrict__ in1,
                                                       long long n, int logG, int tn2, long long block0,
                                                       const unsigned long long *__restrict__ sums, int dc_remove,
                                                       int conj_in, const float2 *__restrict__ twh,
                                                       float4 *__restrict__ z) {
    extern __shared__ __align__(16) unsigned char lag_smem[];
    const int G = 1 << logG;
    float4 *A = reinterpret_cast<float4 *>(lag_smem);
    float4 *B = A + G * tn2;
    float2 *tw = reinterpret_cast<float2 *>(B + G * tn2);          // W_G^i
    const int t = threadIdx.x;
    const long long blk = block0 + blockIdx.y;
    const int n2_0 = blockIdx.x * tn2;
    for (int i = t; i < G; i += 256) {
        float sn, cs;
        sincospif(-2.f * (float)i / (float)G, &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    float m0i = 127.5f, m0q = 127.5f, m1i = 127.5f, m1q = 127.5f;
    if (U8 && dc_remove) {
        const double inv = 1.0 / (double)n;
        m0i = (float)((double)sums[4 * blk + 0] * inv); m0q = (float)((double)sums[4 * blk + 1] * inv);
        m1i = (float)((double)sums[4 * blk + 2] * inv); m1q = (float)((double)sums[4 * blk + 3] * inv);
    }
    const float sc = 1.0f / 127.5f;
    const float cj = conj_in ? -1.f : 1.f;
    // work item = (n1, group of 4 consecutive n2): one 8-byte load per channel (raw bytes) or two 16-byte
    // loads (complex64); rows beyond the data (zero padding) are written without touching memory
    const int q4 = tn2 >> 2;
    const size_t elem = U8 ? 2 : 8;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(in0) | (in1 ? reinterpret_cast<uintptr_t>(in1) : 0)) % (U8 ? 8 : 16)) == 0 &&
                        ((size_t)n * elem) % (U8 ? 8 : 16) == 0;
    for (int it = t; it < G * q4; it += 256) {
        const int n1 = it / q4, j2 = (it % q4) << 2;
        const long long s = (long long)n1 * N + n2_0 + j2;
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s + 3 < n && vec_ok) {
            if (U8) {
                const uint2 a = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(in0) + 2 * (blk * n + s));
                const uint2 b = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(in1) + 2 * (blk * n + s));
                const unsigned aw[2] = {a.x, a.y}, bw[2] = {b.x, b.y};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned pa = aw[u >> 1] >> (16 * (u & 1)), pb = bw[u >> 1] >> (16 * (u & 1));
                    v[u] = make_float4(((float)(pa & 255u) - m0i) * sc, ((float)(pb & 255u) - m1i) * sc,
                                       ((float)((pa >> 8) & 255u) - m0q) * sc * cj, ((float)((pb >> 8) & 255u) - m1q) * sc * cj);
                }
            } else {
                const float4 *pa = reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(in0) + blk * n + s);
                const float4 a0 = pa[0], a1 = pa[1];
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (in1) {
                    const float4 *pb = reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(in1) + blk * n + s);
                    b0 = pb[0]; b1 = pb[1];
                }
                v[0] = make_float4(a0.x, b0.x, a0.y * cj, b0.y * cj);
                v[1] = make_float4(a0.z, b0.z, a0.w * cj, b0.w * cj);
                v[2] = make_float4(a1.x, b1.x, a1.y * cj, b1.y * cj);
                v[3] = make_float4(a1.z, b1.z, a1.w * cj, b1.w * cj);
            }
        } else if (s < n) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s + u >= n) break;
                if (U8) {
                    const uchar2 a = reinterpret_cast<const uchar2 *>(in0)[blk * n + s + u];
                    const uchar2 b = reinterpret_cast<const uchar2 *>(in1)[blk * n + s + u];
                    v[u] = make_float4(((float)a.x - m0i) * sc, ((float)b.x - m1i) * sc, ((float)a.y - m0q) * sc * cj,
                                       ((float)b.y - m1q) * sc * cj);
                } else {
                    const float2 a = reinterpret_cast<const float2 *>(in0)[blk * n + s + u];
                    const float2 b = in1 ? reinterpret_cast<const float2 *>(in1)[blk * n + s + u] : make_float2(0.f, 0.f);
                    v[u] = make_float4(a.x, b.x, a.y * cj, b.y * cj);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) A[n1 * tn2 + j2 + u] = v[u];
    }
    __syncthreads();
    // G-point forward Stockham autosort over n1 for the tn2 columns (element (r, j2) at r*tn2 + j2)
    int ns = 1;
    if (logG & 1) {
        const int half = G >> 1;
        for (int e = t; e < half * tn2; e += 256) {
            const int b = e / tn2, j2 = e % tn2;
            const float4 v0 = A[b * tn2 + j2], v1 = A[(b + half) * tn2 + j2];
            B[(2 * b) * tn2 + j2] = f4add(v0, v1);
            B[(2 * b + 1) * tn2 + j2] = f4sub(v0, v1);
        }
        __syncthreads();
        float4 *tmp = A; A = B; B = tmp;
        ns = 2;
    }
    const int quarter = G >> 2;
    for (; ns < G; ns <<= 2) {
        const int step = G / (4 * ns);
        for (int e = t; e < quarter * tn2; e += 256) {
            const int b = e / tn2, j2 = e % tn2;
            const int k = b & (ns - 1);
            const float4 a0 = A[b * tn2 + j2];
            float4 a1 = A[(b + quarter) * tn2 + j2], a2 = A[(b + 2 * quarter) * tn2 + j2], a3 = A[(b + 3 * quarter) * tn2 + j2];
            if (k) {
                const float2 w1 = tw[k * step], w2 = tw[2 * k * step], w3 = tw[3 * k * step];
                a1 = tw_mul(a1, w1.x, w1.y);
                a2 = tw_mul(a2, w2.x, w2.y);
                a3 = tw_mul(a3, w3.x, w3.y);
            }
            const float4 s02 = f4add(a0, a2), d02 = f4sub(a0, a2);
            const float4 s13 = f4add(a1, a3), d13 = mul_mi(f4sub(a1, a3));        // -i * (a1 - a3)
            const int o = ((b - k) << 2) + k;
            B[o * tn2 + j2] = f4add(s02, s13);
            B[(o + ns) * tn2 + j2] = f4add(d02, d13);
            B[(o + 2 * ns) * tn2 + j2] = f4sub(s02, s13);
            B[(o + 3 * ns) * tn2 + j2] = f4sub(d02, d13);
        }
        __syncthreads();
        float4 *tmp = A; A = B; B = tmp;
    }
    // twiddle W_M^(n2*k1), M = G*4096, and store Z[block][k1][n2]
    float4 *zb = z + (long long)blockIdx.y * G * N;
    for (int e = t; e < G * tn2; e += 256) {
        const int k1 = e / tn2, j2 = e % tn2;
        const int n2 = n2_0 + j2;
        float4 v = A[e];
        if (k1) {
            const float2 w = twh[(long long)k1 * N + n2];
            v = tw_mul(v, w.x, w.y);
        }
        __stcs(zb + (long long)k1 * N + n2, v);
    }
}

// ---- the same head with the G-point transform in REGISTERS ----------------------------------------------
// lag_head_kernel walks log4(G) Stockham passes through shared memory (10 x 16 bytes of LSU traffic per
// element: 207 us of shared-memory time for C2's 92 blocks, the kernel took 460 us).  Here G = Ga*Gb
// (Ga = min(G, 16)), n1 = Gb*a + b, k1 = ka + Ga*kb:
//     W_G^(n1*k1) = W_Ga^(a*ka) * W_G^(b*ka) * W_Gb^(b*kb)
//   step 1  thread (b, n2): its Ga inputs x[(Gb*a + b)*4096 + n2] (rows past the data are the zero padding and
//           are never read), DFT over a in registers with the packed-FP32 butterflies of fx_common.cuh,
//           twiddle W_G^(b*ka)                                   -> shared memory [ka][b][n2]   (one exchange)
//   step 2  thread (ka, n2): DFT over b in registers, twiddle W_M^(n2*k1), one 16-byte store per k1 into Z
// (G <= 16: no exchange at all).  grid = (4096/TN2, blocks), 256 threads, TN2 = 256/Gb columns per CTA,
// shared memory G*TN2*16 B (64 KB for G >= 32) + the W_G table.
#ifndef FX_LAG_THREADS
#define FX_LAG_THREADS 256
#endif
// three CTAs per SM (80 registers, no spills; the 64 KB exchange tiles allow no more): the kernel is
// latency-bound, 458 -> 400 us for C2's 92 blocks against two per SM (profiles/r02_lag_head2.txt)
#ifndef FX_LAG_MINB
#define FX_LAG_MINB 3
#endif
constexpr int kLagThreads = FX_LAG_THREADS;
template <int LOGG>
struct LagSplit {
    static constexpr int G = 1 << LOGG;
    static constexpr int LA = LOGG < 4 ? LOGG : 4;
    static constexpr int Ga = 1 << LA, Gb = G / Ga;
    static constexpr int TN2 = kLagThreads / Gb;
    static constexpr size_t smem = (Gb > 1 ? (size_t)G * TN2 * sizeof(float4) : 0) + (size_t)G * sizeof(float2);
};
template <int R>
__device__ __forceinline__ void dft_regs(C2 *v) {
    if constexpr (R == 16) {
        C2(&v16)[16] = reinterpret_cast<C2(&)[16]>(*v);
        dft16(v16);
    } else if constexpr (R > 1) {
        fused4096::dft_small<R>(v);
    }
}
template <bool U8, int LOGG>
__global__ void __launch_bounds__(kLagThreads, FX_LAG_MINB) lag_head2_kernel(const void *__restrict__ in0, const void *__restrict__ in1,
                                                           long long n, long long block0,
                                                           const unsigned long long *__restrict__ sums, int dc_remove,
                                                           int conj_in, const float2 *__restrict__ twh,
                                                           float4 *__restrict__ z) {
    using L = LagSplit<LOGG>;
    constexpr int G = L::G, Ga = L::Ga, Gb = L::Gb, TN2 = L::TN2;
    extern __shared__ __align__(16) unsigned char lag_smem[];
    float4 *X = reinterpret_cast<float4 *>(lag_smem);                                   // [Ga][Gb][TN2]
    float2 *tw = reinterpret_cast<float2 *>(lag_smem + (Gb > 1 ? (size_t)G * TN2 * sizeof(float4) : 0));   // W_G^i
    const int t = threadIdx.x;
    const long long blk = block0 + blockIdx.y;
    const int n2_0 = blockIdx.x * TN2;
    for (int i = t; i < G; i += kLagThreads) {
        float sn, cs;
        sincospif(-2.f * (float)i / (float)G, &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    float m0i = 127.5f, m0q = 127.5f, m1i = 127.5f, m1q = 127.5f;
    if (U8 && dc_remove) {
        const double inv = 1.0 / (double)n;
        m0i = (float)((double)sums[4 * blk + 0] * inv); m0q = (float)((double)sums[4 * blk + 1] * inv);
        m1i = (float)((double)sums[4 * blk + 2] * inv); m1q = (float)((double)sums[4 * blk + 3] * inv);
    }
    const float sc = 1.0f / 127.5f;
    const float cj = conj_in ? -1.f : 1.f;
    if (Gb > 1) __syncthreads();                       // the W_G table is read in step 1
    float4 *zb = z + (long long)blockIdx.y * G * fused4096::N;
    // ---- step 1 ----
    {
        const int b = t / TN2, j2 = t % TN2, n2 = n2_0 + j2;
        C2 v[Ga];
#pragma unroll
        for (int a = 0; a < Ga; ++a) {
            const long long s = (long long)(Gb * a + b) * fused4096::N + n2;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < n) {
                if (U8) {
                    const uchar2 pa = reinterpret_cast<const uchar2 *>(in0)[blk * n + s];
                    const uchar2 pb = reinterpret_cast<const uchar2 *>(in1)[blk * n + s];
                    q = make_float4(((float)pa.x - m0i) * sc, ((float)pb.x - m1i) * sc, ((float)pa.y - m0q) * sc * cj,
                                    ((float)pb.y - m1q) * sc * cj);
                } else {
                    const float2 pa = reinterpret_cast<const float2 *>(in0)[blk * n + s];
                    const float2 pb = in1 ? reinterpret_cast<const float2 *>(in1)[blk * n + s] : make_float2(0.f, 0.f);
                    q = make_float4(pa.x, pb.x, pa.y * cj, pb.y * cj);
                }
            }
            v[a] = {f2(q.x, q.y), f2(q.z, q.w)};
        }
        dft_regs<Ga>(v);
#pragma unroll
        for (int j = 0; j < Ga; ++j) {
            const int ka = fused4096::perm_rp(Ga, j);
            C2 y = v[j];
            if (Gb > 1) {
                if (ka != 0) {                          // b == 0 multiplies by 1 (kept: no divergence inside a warp for TN2 >= 32)
                    const float2 w = tw[(b * ka) & (G - 1)];
                    y = cmuls(y, w.x, w.y);
                }
                X[(ka * Gb + b) * TN2 + j2] = make_float4(y.r.x, y.r.y, y.i.x, y.i.y);
            } else {
                if (ka != 0) {
                    const float2 w = twh[(long long)ka * fused4096::N + n2];
                    y = cmuls(y, w.x, w.y);
                }
                __stcs(zb + (long long)ka * fused4096::N + n2, make_float4(y.r.x, y.r.y, y.i.x, y.i.y));
            }
        }
    }
    if (Gb == 1) return;
    __syncthreads();
    // ---- step 2 ----
#pragma unroll 1
    for (int task = t; task < Ga * TN2; task += kLagThreads) {
        const int ka = task / TN2, j2 = task % TN2, n2 = n2_0 + j2;
        C2 u[Gb];
#pragma unroll
        for (int b = 0; b < Gb; ++b) {
            const float4 q = X[(ka * Gb + b) * TN2 + j2];
            u[b] = {f2(q.x, q.y), f2(q.z, q.w)};
        }
        dft_regs<Gb>(u);
#pragma unroll
        for (int j = 0; j < Gb; ++j) {
            const int k1 = ka + Ga * fused4096::perm_rp(Gb, j);
            C2 y = u[j];
            if (k1 != 0) {
                const float2 w = twh[(long long)k1 * fused4096::N + n2];
                y = cmuls(y, w.x, w.y);
            }
            __stcs(zb + (long long)k1 * fused4096::N + n2, make_float4(y.r.x, y.r.y, y.i.x, y.i.y));
        }
    }
}

// d_xacc[k1 + G*k2] (=|+=) sum over the segments of virtual block k1 of part_x[s][k2].   grid = (M/256)
__global__ void __launch_bounds__(256) lag_fold_kernel(const float2 *__restrict__ part_x, int logG,
                                                       const int *__restrict__ vblk_first, int first,
                                                       float2 *__restrict__ xacc) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int k1 = (int)(c & ((1 << logG) - 1));
    const long long k2 = c >> logG;
    float xr = 0.f, xi = 0.f;
    for (int s = vblk_first[k1]; s < vblk_first[k1 + 1]; ++s) {
        const float2 x = part_x[(long long)s * N + k2];
        xr += x.x; xi += x.y;
    }
    if (!first) { const float2 o = xacc[c]; xr += o.x; xi += o.y; }
    xacc[c] = make_float2(xr, xi);
}

// |xc|^2 (times M^2) of lag index idx in [0, M), read from the tail kernel's auto-power partials of the
// inverse pass (one frame, one segment per virtual block)
__device__ __forceinline__ float lag_mag2(const float2 *part_a, const int *vblk_first, int logG, long long idx) {
    const int k1 = (int)(idx & ((1 << logG) - 1));
    return part_a[(long long)vblk_first[k1] * N + (idx >> logG)].x;
}

__global__ void __launch_bounds__(256) lag_argmax_big_stage1(const float2 *__restrict__ part_a,
                                                             const int *__restrict__ vblk_first, int logG, long long n,
                                                             long long M, float *__restrict__ pval,
                                                             long long *__restrict__ pidx) {
    generic::ArgMax best{-1.f, 1ll << 62};
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < 2 * n;
         j += (long long)gridDim.x * blockDim.x) {
        const long long l = j - n;
        best = generic::better(best, generic::ArgMax{lag_mag2(part_a, vblk_first, logG, l >= 0 ? l : l + M), j});
    }
    best = generic::block_argmax(best);
    if (threadIdx.x == 0) { pval[blockIdx.x] = best.val; pidx[blockIdx.x] = best.idx; }
}

__global__ void __launch_bounds__(256) lag_argmax_big_stage2(const float2 *__restrict__ part_a,
                                                             const int *__restrict__ vblk_first, int logG, long long n,
                                                             long long M, const float *__restrict__ pval,
                                                             const long long *__restrict__ pidx, int nparts, float scale,
                                                             long long *__restrict__ out_idx, float *__restrict__ out_nb) {
    generic::ArgMax best{-1.f, 1ll << 62};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) best = generic::better(best, generic::ArgMax{pval[i], pidx[i]});
    best = generic::block_argmax(best);
    if (threadIdx.x == 0) {
        const long long imax = best.idx;
        out_idx[0] = imax;
        for (int d = -1; d <= 1; ++d) {
            long long j = imax + d;
            float r = -1.f;
            if (j < 0) j += 2 * n;            // python negative index wraps (xcorr[-1])
            if (j < 2 * n) {
                const long long l = j - n;
                r = sqrtf(lag_mag2(part_a, vblk_first, logG, l >= 0 ? l : l + M)) * scale;
            }
            out_nb[d + 1] = r;
        }
    }
}

}  // namespace lag
}  // namespace fx
