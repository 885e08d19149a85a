// fx_bigfft.cuh -- nbins = G * 4096 (8192 .. 65536), ntaps = 4: two kernels around ONE intermediate.
//
// FIR state for more than 4096 positions does not fit tensor memory, and a frame does not fit shared
// memory, so the N-point transform is split once, decimation in frequency (n = n1*4096 + n2, n1 < G):
//
//   X[k1 + G*k2] = sum_n2 W4096^(n2*k2) * [ W_N^(n2*k1) * sum_n1 w[n1*4096 + n2] * W_G^(n1*k1) ]
//
//   head_kernel   unpack + DC removal + 4-tap FIR (transposed form, state in registers) per polyphase position,
//                 G-point DFT over n1 in registers, twiddle W_N^(n2*k1)  ->  Z[block][frame][k1][n2]  (16 B:
//                 both channels, re/im), written once, read once
//   tail_kernel   the fused kernel's stages A, B, C and X-engine (fx_fused4096s.cuh: staggered warp
//                 groups, padded exchange tiles, packed FP32) with stage A fed from Z instead of the FIR;
//                 a "virtual block" is (block, k1): its 4096 accumulated bins are k1 + G*k2
//
// HBM traffic per pair-sample: 4 B raw (re-read through L2 by the four taps) + 16 B Z out + 16 B Z in,
// against ~100 B for the unfused kernels (FIR out, two FFT passes in and out, X-engine in).
// Replaces the same reference lines as the fused kernel (effex.py:394-395, :508-509, :520-521, :553).
#pragma once
#include "fx_fused4096s.cuh"

namespace fx {
namespace bigfft {

using fused4096::N;        // 4096: length of the transforms the tail kernel runs
using fused4096::NT;
using fused4096::NP;
using fused4096::ROWP;
using fused4096::Segment;
using fused4096::TILE;

// ---- head ------------------------------------------------------------------------------------------
// A CTA owns 256 polyphase positions -- a tile of TN2 = 256/G consecutive n2 for all G values of n1 --
// and walks kHeadFrames consecutive frames of one block (plus 3 frames of warm-up when it does not start
// at the block's first frame).  Two phases per batch of G frames:
//   1. thread = position: raw bytes of the batch are loaded up front (one uchar2 per channel per frame),
//      unpacked ONCE (PRMT into the mantissa of 2^15, as in the fused kernel) and pushed through the
//      transposed-form FIR whose state (z1..z3) stays in registers; FIR outputs go to shared memory;
//   2. thread = (frame of the batch, n2): G-point DFT over n1 in registers, twiddle W_N^(n2*k1) (host table,
//      the CTA's slice staged in shared memory), the next batch's raw loads already in flight, and one
//      16-byte store per k1 into Z (256-byte runs per warp half).
// grid = (4096/TN2, ceil(P/kHeadFrames), n_blocks), 256 threads, dynamic smem G*4 KB
#ifndef FX_HEAD_FRAMES
#define FX_HEAD_FRAMES 64
#endif
constexpr int kHeadFrames = FX_HEAD_FRAMES;
#ifndef FX_HEAD_CTAS
#define FX_HEAD_CTAS 2
#endif
// SPAN = false: reference mode (whole blocks, i_begin = 0, per-block mean over S samples, zero history);
// SPAN = true: a chunk of frames of one streaming span (mean over mean_count samples, optional halo)
template <int LOGG, bool SPAN>
__global__ void __launch_bounds__(256, FX_HEAD_CTAS) head_kernel(const uint8_t *__restrict__ iq0, const uint8_t *__restrict__ iq1,
                                                   long long S, int i_begin, int i_end, const float *__restrict__ taps,
                                                   const unsigned long long *__restrict__ sums, int dc_remove,
                                                   long long mean_count, const uint8_t *__restrict__ halo0,
                                                   const uint8_t *__restrict__ halo1, const float2 *__restrict__ twh,
                                                   float4 *__restrict__ z) {
    using fused4096::byte_to_magic;
    using fused4096::kMagic;
    constexpr int G = 1 << LOGG;
    constexpr int NB = N << LOGG;                  // nbins
    constexpr int TN2 = 256 >> LOGG;               // n2 values per CTA
    extern __shared__ __align__(16) unsigned char head_smem[];
    float4(*W)[256] = reinterpret_cast<float4(*)[256]>(head_smem);        // [G][256] FIR outputs of a batch
    __shared__ float s_nm[4];                      // 128 - byte mean: ch0 I, ch0 Q, ch1 I, ch1 Q
    __shared__ float2 s_tw[G][TN2];                // W_N^(n2*k1) of this CTA's n2 tile (fixed for all its frames)
    const int t = threadIdx.x;
    const int b = blockIdx.z;
    // frames [i_begin, i_end) of the block go to Z rows 0 .. i_end - i_begin (reference mode: the whole block;
    // streaming mode: one chunk of the span, mean over mean_count samples, halo = the 3 frames before frame 0)
    if (!SPAN) i_begin = 0;
    const int i0 = i_begin + blockIdx.y * kHeadFrames;
    const int i1 = min(i0 + kHeadFrames, i_end);
    if (t < 4) s_nm[t] = dc_remove ? (float)(128.0 - (double)sums[4ll * b + t] / (double)(SPAN ? mean_count : S)) : 0.5f;
    s_tw[t / TN2][t % TN2] = twh[(t / TN2) * N + blockIdx.x * TN2 + t % TN2];
    __syncthreads();
    const float2 nmI = f2(s_nm[0], s_nm[2]), nmQ = f2(s_nm[1], s_nm[3]);
    const float2 mg = f2(-kMagic, -kMagic);
    // phase-1 role: position (n1, n2)
    const int n1 = t / TN2, n2 = blockIdx.x * TN2 + t % TN2;
    const int n = n1 * N + n2;
    const unsigned short *x0 = reinterpret_cast<const unsigned short *>(iq0) + (long long)b * S + n;
    const unsigned short *x1 = reinterpret_cast<const unsigned short *>(iq1) + (long long)b * S + n;
    const float t0 = taps[n], t1 = taps[(long long)NB + n], t2 = taps[2ll * NB + n], t3 = taps[3ll * NB + n];
    float2 z1r = f2(0.f, 0.f), z1i = z1r, z2r = z1r, z2i = z1r, z3r = z1r, z3i = z1r;
    // one sample through the FIR; returns the output of its frame
    auto push = [&](uint32_t w) -> C2 {
        const float2 yr = f2add(f2add(f2(byte_to_magic<0>(w), byte_to_magic<2>(w)), mg), nmI);
        const float2 yi = f2add(f2add(f2(byte_to_magic<1>(w), byte_to_magic<3>(w)), mg), nmQ);
        const C2 out = {f2fmas(yr, t0, z1r), f2fmas(yi, t0, z1i)};
        z1r = f2fmas(yr, t1, z2r); z1i = f2fmas(yi, t1, z2i);
        z2r = f2fmas(yr, t2, z3r); z2i = f2fmas(yi, t2, z3i);
        z3r = f2muls(yr, t3);      z3i = f2muls(yi, t3);
        return out;
    };
    auto raw = [&](int i) -> uint32_t {            // (I0, Q0, I1, Q1) of frame i >= 0 at this position
        const long long s = (long long)i * NB;
        return (uint32_t)x0[s] | ((uint32_t)x1[s] << 16);
    };
    auto raw_hist = [&](int i) -> uint32_t {       // warm-up only: frames -3..-1 come from the halo (streaming mode)
        if (!SPAN || i >= 0) return raw(i);
        const long long s = (long long)(i + 3) * NB + n;
        return (uint32_t)reinterpret_cast<const unsigned short *>(halo0)[s] |
               ((uint32_t)reinterpret_cast<const unsigned short *>(halo1)[s] << 16);
    };
    const int first_hist = (SPAN && halo0) ? -3 : 0;         // earliest frame that exists
    // phase-2 role: (frame slot, n2)
    const int fs = t / TN2, j2 = t % TN2;
    const int m2 = blockIdx.x * TN2 + j2;
    uint32_t w[G];                                 // raw words of the current batch, loaded one batch ahead
    {
        // warm-up (history of the first frame) and first batch: all loads issued before the first use
        uint32_t wu[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) wu[k] = i0 - 3 + k >= first_hist ? raw_hist(i0 - 3 + k) : 0u;
#pragma unroll
        for (int f = 0; f < G; ++f) w[f] = i0 + f < i1 ? raw(i0 + f) : 0u;
        if (i0 - 3 >= first_hist) {
#pragma unroll
            for (int k = 0; k < 3; ++k) push(wu[k]);
        } else {
            for (int k = 3 - (i0 - first_hist); k < 3; ++k) push(wu[k]);   // zero history before the first frame
        }
    }
    for (int ib = i0; ib < i1; ib += G) {
#pragma unroll
        for (int f = 0; f < G; ++f) {
            const C2 o = push(w[f]);
            W[f][t] = make_float4(o.r.x, o.r.y, o.i.x, o.i.y);
        }
        __syncthreads();
        // next batch's bytes: in flight during phase 2
#pragma unroll
        for (int f = 0; f < G; ++f) w[f] = ib + G + f < i1 ? raw(ib + G + f) : 0u;
        if (ib + fs < i1) {
            C2 v[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 q = W[fs][g * TN2 + j2];
                v[g] = {f2(q.x, q.y), f2(q.z, q.w)};
            }
            if constexpr (G == 16) {
                C2(&v16)[16] = reinterpret_cast<C2(&)[16]>(v);
                dft16(v16);
            } else {
                fused4096::dft_small<G>(v);
            }
            float4 *zf = z + ((long long)b * (i_end - i_begin) + (ib + fs - i_begin)) * (long long)NB + m2;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const int k1 = fused4096::perm_rp(G, j);
                C2 y = v[j];
                if (k1 != 0) {
                    const float2 wk = s_tw[k1][j2];
                    y = cmuls(y, wk.x, wk.y);
                }
                __stcs(zf + (long long)k1 * N, make_float4(y.r.x, y.r.y, y.i.x, y.i.y));
            }
        }
        __syncthreads();
    }
}

// ---- head, persistent variant ---------------------------------------------------------------------
// The same arithmetic as head_kernel with the fused kernel's machinery instead of two phases around a
// shared-memory transpose: a thread owns 16 points of the frame, n = n1*4096 + n2 with n1 = r % G and
// n2 = tile*(4096/G) + 256*(r / G) + t, so the G-point DFT over n1 runs on its own registers; the FIR state
// (16 points x 12 floats) and the taps live in TENSOR MEMORY, the raw bytes arrive through a ring of TMA bulk
// copies (32 rows of 512 bytes per frame) with full/empty mbarriers and no CTA-wide barrier, and every
// thread stores its 16 values of Z.  One persistent CTA (8 compute warps + 1 producer warp) per SM walks segments of virtual blocks
// vb = block*G + tile -- the SAME segment plan the tail kernel walks over vb = block*G + k1.
// Reference mode (whole blocks, per-block mean, zero history) and streaming spans (chunks of one span's frames,
// recording-wide mean, halo frames); head_kernel remains for inputs that are not 16-byte aligned.
#ifndef FX_HRING
#define FX_HRING 4
#endif
constexpr int HRING = FX_HRING;
struct __align__(16) SmemH {
    unsigned short raw[HRING][2][16][256];       // 4 x 16 KB: [slot][channel][point slot r][t] (I, Q) byte pairs
    unsigned long long full[HRING], empty[HRING];
    uint32_t tmem_base;
};
struct Head2Params {
    const uint8_t *iq0, *iq1;        // [n_blocks][2*S]
    long long S;
    int P;                           // frames per block
    const float *taps;               // [4][NB] reversed within a branch, scaled by 1/127.5
    const unsigned long long *sums;  // [n_blocks][2 ch][2 comp]
    int dc_remove;
    const float2 *twh;               // W_NB^(n2*k1), [G][4096]
    float4 *z;                       // [n_blocks][P][G][4096]
    const Segment *segs;             // segments over virtual blocks vb = block*G + tile
    const int *cta_first;
    // streaming spans (one unit walked in chunks of frames): Z row 0 is frame i_begin of the span, the mean is taken
    // over mean_count samples, and the frames before frame 0 come from the halo (3 frames of NB samples per channel,
    // 16-byte aligned) when there is one.  Reference mode: i_begin = 0, first_hist = 0, mean_count = S, no halo.
    int i_begin, first_hist;
    long long mean_count;
    const uint8_t *halo0, *halo1;
};
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fused4096::smem_u32(bar)) : "memory");
}

constexpr int kHead2Threads = 288;      // 8 compute warps + the producer warp
template <int LOGG>
__global__ void __launch_bounds__(kHead2Threads, 1) head2_kernel(const Head2Params prm) {
    using namespace fused4096;
    constexpr int G = 1 << LOGG;
    constexpr int NB = N << LOGG;                  // nbins
    constexpr int TW = N >> LOGG;                  // n2 values per tile = 256 * (16 / G)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemH &sm = *reinterpret_cast<SmemH *>(smem_raw);
    const int t = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    const int lane = t & 31;
    if (t == 0) {
        for (int s = 0; s < HRING; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 256);          // every compute thread arrives once it has read its bytes
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_pts = sm.tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));

    const int seg_begin = prm.cta_first[blockIdx.x], seg_end = prm.cta_first[blockIdx.x + 1];
    // ---- producer warp (warp 8): walks the CTA's ingest items in order, item q -> ring slot q % HRING; lane =
    //      16*channel + r issues the 512-byte row of point slot r (ptxas serialises the 32 bulk copies of a
    //      frame on the uniform datapath: ~25 cycles each, which is why they have a warp of their own) ----
    if (warp == 8) {
        uint32_t pcnt = 0;
        const int ch = lane >> 4, r = lane & 15;
        for (int ps = seg_begin; ps < seg_end; ++ps) {
            const Segment g = prm.segs[ps];
            const int lo_f = prm.first_hist - prm.i_begin;             // earliest frame that exists, chunk-relative
            const int g0 = g.f0 - 3 > lo_f ? g.f0 - 3 : lo_f;
            const int n_ing = g.f0 + g.nf - g0;
            const int blk = g.block >> LOGG, tile = g.block & (G - 1);
            const long long row = (long long)(r & (G - 1)) * N + tile * TW + (r >> LOGG) * 256;     // samples into a frame
            const uint8_t *body = (ch ? prm.iq1 : prm.iq0) + 2ll * prm.S * blk + 2ll * row;
            const uint8_t *halo = (ch ? prm.halo1 : prm.halo0) + 2ll * row;
            for (int pj = 0; pj < n_ing; ++pj, ++pcnt) {
                const long long a = (long long)prm.i_begin + g0 + pj;  // frame of the span (negative: halo)
                const uint8_t *src = a >= 0 ? body + 2ll * a * NB : halo + 2ll * (a + 3) * NB;
                const uint32_t slot = pcnt % HRING;
                if (pcnt >= HRING) mbar_wait(&sm.empty[slot], ((pcnt / HRING) - 1) & 1u);
                if (lane == 0) mbar_expect_tx(&sm.full[slot], 2u * 16u * 512u);
                __syncwarp();
                tma_load_1d(&sm.raw[slot][ch][r][0], src, 512u, &sm.full[slot]);
            }
        }
    }

    uint32_t ccnt = 0;
    for (int seg = seg_begin; seg < seg_end && warp < 8; ++seg) {
        const Segment sg = prm.segs[seg];
        const int blk = sg.block >> LOGG, tile = sg.block & (G - 1);
        float2 nmI, nmQ;
        if (prm.dc_remove) {
            const unsigned long long *su = prm.sums + 4ll * blk;
            const double inv = 1.0 / (double)prm.mean_count;
            nmI = f2((float)(128.0 - (double)su[0] * inv), (float)(128.0 - (double)su[2] * inv));
            nmQ = f2((float)(128.0 - (double)su[1] * inv), (float)(128.0 - (double)su[3] * inv));
        } else {
            nmI = f2(0.5f, 0.5f);
            nmQ = f2(0.5f, 0.5f);
        }
        // taps of this tile's points and zero FIR state -> tensor memory; twiddles of its outputs -> registers
        tmem_wait_st();
        float2 tw[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n2 = tile * TW + (r >> LOGG) * 256 + t;
            const int n = (r & (G - 1)) * N + n2;
            tmem_st4(tm_pts + 16 * r, make_float4(prm.taps[n], prm.taps[(long long)NB + n], prm.taps[2ll * NB + n],
                                                  prm.taps[3ll * NB + n]));
            tmem_st4(tm_pts + 16 * r + 4, make_float4(0.f, 0.f, 0.f, 0.f));
            tmem_st4(tm_pts + 16 * r + 8, make_float4(0.f, 0.f, 0.f, 0.f));
            tmem_st4(tm_pts + 16 * r + 12, make_float4(0.f, 0.f, 0.f, 0.f));
            // register r of its column's DFT output holds k1 = perm_rp(G, r % G)
            tw[r] = prm.twh[perm_rp(G, r & (G - 1)) * N + n2];
        }
        const int lo_f = prm.first_hist - prm.i_begin;
        const int g0 = sg.f0 - 3 > lo_f ? sg.f0 - 3 : lo_f;
        const int n_ing = sg.f0 + sg.nf - g0;
        float4 *zblk = prm.z + ((long long)blk * prm.P * G) * N + tile * TW + t;
        const float2 mg = f2(-kMagic, -kMagic);
#pragma unroll 1
        for (int j = 0; j < n_ing; ++j, ++ccnt) {
            const uint32_t slot = ccnt % HRING;
            mbar_wait(&sm.full[slot], (ccnt / HRING) & 1u);
            uint32_t cur[16];
#pragma unroll
            for (int r = 0; r < 16; ++r)
                cur[r] = __byte_perm((uint32_t)sm.raw[slot][0][r][t], (uint32_t)sm.raw[slot][1][r][t], 0x5410);   // (I0,Q0,I1,Q1)
            mbar_arrive(&sm.empty[slot]);          // (per thread, not per warp: compute-sanitizer's racecheck follows it)
            tmem_wait_st();                        // the previous frame's state stores have landed
            C2 v[16];
            {
                float4 tp[2], z1[2], z2[2], z3[2];
                tmem_ld8(tm_pts, tp[0], z1[0]);
                tmem_ld8(tm_pts + 8, z2[0], z3[0]);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int c = r & 1, nx = c ^ 1;
                    tmem_wait_ld(tp[c], z1[c], z2[c], z3[c]);
                    if (r + 1 < 16) {
                        tmem_ld8(tm_pts + 16 * (r + 1), tp[nx], z1[nx]);
                        tmem_ld8(tm_pts + 16 * (r + 1) + 8, z2[nx], z3[nx]);
                    }
                    const uint32_t w = cur[r];
                    const float2 yr = f2add(f2add(f2(byte_to_magic<0>(w), byte_to_magic<2>(w)), mg), nmI);
                    const float2 yi = f2add(f2add(f2(byte_to_magic<1>(w), byte_to_magic<3>(w)), mg), nmQ);
                    v[r] = {f2fmas(yr, tp[c].x, f2(z1[c].x, z1[c].y)), f2fmas(yi, tp[c].x, f2(z1[c].z, z1[c].w))};
                    const float2 s1r = f2fmas(yr, tp[c].y, f2(z2[c].x, z2[c].y)), s1i = f2fmas(yi, tp[c].y, f2(z2[c].z, z2[c].w));
                    const float2 s2r = f2fmas(yr, tp[c].z, f2(z3[c].x, z3[c].y)), s2i = f2fmas(yi, tp[c].z, f2(z3[c].z, z3[c].w));
                    const float2 s3r = f2muls(yr, tp[c].w), s3i = f2muls(yi, tp[c].w);
                    tmem_st4(tm_pts + 16 * r + 4, make_float4(s1r.x, s1r.y, s1i.x, s1i.y));
                    tmem_st4(tm_pts + 16 * r + 8, make_float4(s2r.x, s2r.y, s2i.x, s2i.y));
                    tmem_st4(tm_pts + 16 * r + 12, make_float4(s3r.x, s3r.y, s3i.x, s3i.y));
                }
            }
            if (g0 + j < sg.f0) continue;          // history only
            if constexpr (G == 16) {
                dft16(v);
            } else {
#pragma unroll
                for (int c = 0; c < 16 / G; ++c) dft_small<G>(&v[c * G]);
            }
            float4 *zf = zblk + (long long)(g0 + j) * G * N;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int k1 = perm_rp(G, r & (G - 1));
                C2 y = v[r];
                if (k1 != 0) y = cmuls(y, tw[r].x, tw[r].y);
                __stcs(zf + (long long)k1 * N + (r >> LOGG) * 256, make_float4(y.r.x, y.r.y, y.i.x, y.i.y));
            }
        }
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(sm.tmem_base, 512);
}

// ---- tail ------------------------------------------------------------------------------------------
struct __align__(16) SmemT {
    float4 X[2][NP];             // exchange planes, as in the fused kernel
    float4 twA[8][NT];
    float4 twB[8][16];
    unsigned long long mbar;
};

struct TailParams {
    const float4 *z;             // [n_blocks][P][G][4096]
    const float4 *twAp, *twBp;   // the fused kernel's tables for 4096 bins
    const Segment *segs;         // segments over virtual blocks vb = block * G + k1
    const int *cta_first;
    float2 *part_x, *part_a;     // [n_segs][4096]
    int G, P;
};

template <bool AUTOS>
__global__ void __maxnreg__(FX_MAXNREG) tail_kernel(const TailParams prm) {
    using namespace fused4096;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemT &sm = *reinterpret_cast<SmemT *>(smem_raw);
    const int t = threadIdx.x;
    const int grp = __shfl_sync(0xffffffffu, t >> 7, 0);     // 0: stage A first, 1: stages B, C first
    if (t == 0) {
        mbar_init(&sm.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.mbar, (uint32_t)(sizeof(sm.twA) + sizeof(sm.twB)));
        tma_load_1d(&sm.twA[0][0], prm.twAp, (uint32_t)sizeof(sm.twA), &sm.mbar);
        tma_load_1d(&sm.twB[0][0], prm.twBp, (uint32_t)sizeof(sm.twB), &sm.mbar);
    }
    __syncthreads();
    mbar_wait(&sm.mbar, 0);
    const int k1B = t >> 4, lo = t & 15;
    const long long fstride = (long long)prm.G * N;            // float4 elements between frames of one k1

    const int seg_end = prm.cta_first[blockIdx.x + 1];
    for (int seg = prm.cta_first[blockIdx.x]; seg < seg_end; ++seg) {
        const Segment sg = prm.segs[seg];
        const int blk = sg.block / prm.G, k1 = sg.block % prm.G;
        const float4 *zseg = prm.z + (((long long)blk * prm.P + sg.f0) * prm.G + k1) * N;
        float2 accx[16], acca[AUTOS ? 16 : 1];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            accx[j] = f2(0.f, 0.f);
            if (AUTOS) acca[j] = f2(0.f, 0.f);
        }
        C2 v[16];

        // stage A of frame j of the segment: 16 values of Z -> DFT16 -> twiddle -> exchange plane `buf`
        auto stage_a = [&](int j, int buf) {
            const float4 *zf = zseg + (long long)j * fstride;
            if (t == 0 && j + 1 < sg.nf)       // pull the next frame (64 KB) into L2 while this one is transformed
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(zf + fstride), "r"(N * 16) : "memory");
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float4 q = __ldcs(zf + t + NT * r);      // read once: streaming
                v[r] = {f2(q.x, q.y), f2(q.z, q.w)};
            }
            dft16(v);
            float4 *xx = sm.X[buf];
            float4 tq[2], tqn[2];
            tq[0] = sm.twA[0][t];
            tq[1] = sm.twA[1][t];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g < 3) {
                    tqn[0] = sm.twA[2 * g + 2][t];
                    tqn[1] = sm.twA[2 * g + 3][t];
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int row = g + 4 * b;                 // = perm16(4g + b)
                    C2 zv = v[4 * g + b];
                    const float4 q = tq[b >> 1];
                    if (row != 0) zv = (b & 1) ? cmuls(zv, q.z, q.w) : cmuls(zv, q.x, q.y);
                    sts_c2(&xx[row * TILE + k1B * ROWP + lo], zv);
                }
                tq[0] = tqn[0];
                tq[1] = tqn[1];
            }
        };
        // stages B, C and the X-engine of the frame held in exchange plane `buf`
        auto fft_rest = [&](int buf) {
            float4 *xx = sm.X[buf];
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) v[n2] = lds_c2(&xx[k1B * TILE + n2 * ROWP + lo]);
            dft16(v);
            __syncwarp();
            {
                float4 tq[2], tqn[2];
                tq[0] = sm.twB[0][lo];
                tq[1] = sm.twB[1][lo];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < 3) {
                        tqn[0] = sm.twB[2 * g + 2][lo];
                        tqn[1] = sm.twB[2 * g + 3][lo];
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int k2 = g + 4 * b;
                        C2 zv = v[4 * g + b];
                        const float4 q = tq[b >> 1];
                        if (k2 != 0) zv = (b & 1) ? cmuls(zv, q.z, q.w) : cmuls(zv, q.x, q.y);
                        sts_c2(&xx[k1B * TILE + k2 * ROWP + lo], zv);
                    }
                    tq[0] = tqn[0];
                    tq[1] = tqn[1];
                }
            }
            __syncwarp();
#pragma unroll
            for (int n3 = 0; n3 < 16; ++n3) v[n3] = lds_c2(&xx[k1B * TILE + lo * ROWP + n3]);
            dft16(v);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float re0 = v[jj].r.x, re1 = v[jj].r.y, im0 = v[jj].i.x, im1 = v[jj].i.y;
                accx[jj].x = fmaf(re0, re1, fmaf(im0, im1, accx[jj].x));
                accx[jj].y = fmaf(im0, re1, fmaf(-re0, im1, accx[jj].y));
                if (AUTOS) acca[jj] = f2fma(v[jj].r, v[jj].r, f2fma(v[jj].i, v[jj].i, acca[jj]));
            }
        };

#pragma unroll 1
        for (int s = 0; s <= sg.nf; ++s) {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (half == grp) {
                    if (s < sg.nf) stage_a(s, s & 1);
                } else {
                    if (s >= 1) fft_rest((s - 1) & 1);
                }
            }
            __syncthreads();
        }

        // epilogue: bin k2 + 16*lo... in the fused kernel's order, k2 index = k1B + 16*lo + 256*perm16(jj)
        {
            float2 *xs = reinterpret_cast<float2 *>(&sm.X[0][0]);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int idx = k1B + 16 * lo + 256 * perm16(jj);
                const int sw = idx ^ ((idx >> 4) & 15);
                sts_pair(&xs[sw], accx[jj]);
                if (AUTOS) sts_pair(&xs[N + sw], acca[jj]);
            }
            __syncthreads();
            float2 *px = prm.part_x + (long long)seg * N;
            float2 *pa = prm.part_a + (long long)seg * N;
#pragma unroll
            for (int q = 0; q < N / NT; ++q) {
                const int o = t + NT * q;
                const int sw = o ^ ((o >> 4) & 15);
                px[o] = xs[sw];
                if (AUTOS) pa[o] = xs[N + sw];
            }
            __syncthreads();
        }
    }
}

// ---- finalize ---------------------------------------------------------------------------------------
// out[b][j] = conj(rot[c]) * (1/P) * sum over the segments of virtual block (b, k1) of part_x[s][k2],
// c = (j + NB/2) mod NB = k1 + G*k2.   grid = (NB/256, n_blocks)
__global__ void __launch_bounds__(256) finalize_kernel(const float2 *__restrict__ part_x,
                                                       const float2 *__restrict__ part_a, int NB, int logG,
                                                       const int *__restrict__ vblk_first, float inv_frames,
                                                       const float2 *__restrict__ rot, float2 *__restrict__ xspec,
                                                       float *__restrict__ auto0, float *__restrict__ auto1) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= NB) return;
    const int c = (j + (NB >> 1)) & (NB - 1);
    const int k1 = c & ((1 << logG) - 1), k2 = c >> logG;
    const int vb = (b << logG) + k1;
    float xr = 0.f, xi = 0.f, a0 = 0.f, a1 = 0.f;
    const bool autos = auto0 || auto1;
    for (int s = vblk_first[vb]; s < vblk_first[vb + 1]; ++s) {
        const float2 x = part_x[(long long)s * N + k2];
        xr += x.x; xi += x.y;
        if (autos) {
            const float2 a = part_a[(long long)s * N + k2];
            a0 += a.x; a1 += a.y;
        }
    }
    xr *= inv_frames; xi *= inv_frames;
    const float2 r = rot ? rot[c] : make_float2(1.f, 0.f);
    xspec[(long long)b * NB + j] = make_float2(xr * r.x + xi * r.y, xi * r.x - xr * r.y);
    if (auto0) auto0[(long long)b * NB + j] = a0 * inv_frames;
    if (auto1) auto1[(long long)b * NB + j] = a1 * inv_frames;
}

// ---- integrate --------------------------------------------------------------------------------------
// float64 partial sums of the un-normalised spectra over the blocks of group g, natural bin order, in the
// layout integrate_stage2_kernel folds: scratch[g][ 2c, 2c+1 | 2NB + c | 3NB + c ].   grid = (NB/256, groups)
__global__ void __launch_bounds__(256) integrate_kernel(const float2 *__restrict__ part_x,
                                                        const float2 *__restrict__ part_a, int NB, int logG,
                                                        const int *__restrict__ vblk_first, int n_blocks,
                                                        double *__restrict__ scratch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= NB) return;
    const int k1 = c & ((1 << logG) - 1), k2 = c >> logG;
    const int groups = gridDim.y, g = blockIdx.y;
    const int b0 = (int)((long long)n_blocks * g / groups), b1 = (int)((long long)n_blocks * (g + 1) / groups);
    double xr = 0, xi = 0, a0 = 0, a1 = 0;
    for (int b = b0; b < b1; ++b) {
        const int vb = (b << logG) + k1;
        for (int s = vblk_first[vb]; s < vblk_first[vb + 1]; ++s) {
            const float2 x = part_x[(long long)s * N + k2];
            const float2 a = part_a[(long long)s * N + k2];
            xr += x.x; xi += x.y; a0 += a.x; a1 += a.y;
        }
    }
    double *o = scratch + (long long)g * 4 * NB;
    o[2 * c] = xr;
    o[2 * c + 1] = xi;
    o[2 * NB + c] = a0;
    o[3 * NB + c] = a1;
}

}  // namespace bigfft
}  // namespace fx
