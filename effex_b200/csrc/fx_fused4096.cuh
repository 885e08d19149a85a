// fx_fused4096.cuh -- fused unpack -> 4-tap PFB FIR -> 4096-point FFT -> X-engine (lock-step version).
// The production kernel is fused_kernel_stag in fx_fused4096s.cuh; this simpler one (direct-form FIR from
// three frames of history in registers, taps in shared memory, all warps in the same phase) is kept as an
// on-device cross-check (FX_FLAG_LOCKSTEP_KERNEL) and shares its helpers, tables and thread mapping.
//
// Replaces, for N = 4096, T = 4, the per-block chain of the reference:
//   pyrtlsdr packed_bytes_to_iq  (uint8 -> complex, effex.py:652)
//   DC removal                   (effex.py:394-395; the block means come from fx_block_sums)
//   cusignal channelize_poly     (FIR + cuFFT, effex.py:553), twice (effex.py:508-509)
//   f0*conj(f1) and the frame mean (effex.py:520-521; rot/fftshift/1/P in fx_finalize)
//
// One CTA (256 threads) walks the frames of one segment (a run of frames of
// one block).  Both channels are carried in the two lanes of f32x2 registers,
// so every FIR / butterfly / twiddle instruction is an FFMA2/FADD2/FMUL2 that
// serves both channels.  Per thread: 16 points (radix-16 x 3 stages), 16 bins
// of accumulators (cross re/im + two autos) in registers for the whole segment.
//
// Dataflow per frame (t = thread, u = thread after exchange 1):
//   raw bytes   TMA bulk copy (cp.async.bulk + mbarrier) -> smem ring, 8 KB/channel/frame
//   FIR         w[p] = sum_k h[kN+N-1-p]/127.5 * (b[(i-k)N+p] - mean),  p = t + 256 r
//   stage A     DFT16 over r, twiddle W4096^(t*k1)           -> X[k1*256 + t]
//   stage B     DFT16 over n2 (u = 16*k1 + n3), W256^(n3*k2) -> X[k1*256 + k2*16 + (n3^k2)]  (warp-local)
//   stage C     DFT16 over n3 -> bin k1 + 16*k2 + 256*k3
//   X-engine    acc += F0*conj(F1), |F0|^2, |F1|^2
// tools/proto_fft4096.py emulates exactly this mapping in numpy.
#pragma once
#include "fx_common.cuh"

namespace fx {
namespace fused4096 {

constexpr int N = 4096;
constexpr int T = 4;
constexpr int NT = 256;          // threads per CTA
constexpr int RING = 3;          // raw-frame ring slots
constexpr int FRAME_BYTES = 2 * N;   // per channel

struct __align__(16) Smem {
    float2 Xr[N];                // 32 KB exchange plane: (re ch0, re ch1)
    float2 Xi[N];                // 32 KB exchange plane: (im ch0, im ch1)   (must follow Xr: the epilogue uses both as one 64 KB buffer)
    float4 taps[N];              // 64 KB: taps[p] = h[kN+N-1-p]/127.5, k=0..3
    float2 twA[16][NT];          // W4096^(t*k1), row 0 unused
    float2 twB[16][16];          // W256^(n3*k2)
    unsigned short raw[RING][2][N];   // 48 KB: (I,Q) byte pairs per channel
    unsigned long long mbar[RING + 1];   // [RING] = tables
};

struct Segment {
    int block;       // block index within the call
    int f0;          // first output frame (block-relative); super-frames when nbins < 4096
    int nf;          // number of output frames (super-frames)
    int pad;
};

struct Params {
    const uint8_t *iq0, *iq1;           // [n_blocks][2*S]
    const unsigned long long *sums;     // [n_blocks][2 ch][2 comp] byte sums over the whole block
    const float4 *taps;                 // [N]
    const float2 *twA;                  // [16][256]
    const float2 *twB;                  // [16][16]
    const float4 *twAp, *twBp;          // the same tables, rows paired (k, k+4) per float4 (staggered kernel)
    const Segment *segs;                // all segments, in global frame order
    const int *cta_first;               // [gridDim.x + 1]: CTA c walks segs[cta_first[c] .. cta_first[c+1])
    float2 *part_x;                     // [n_segs][N] sum F0*conj(F1) (natural bin order)
    float2 *part_a;                     // [n_segs][N] (sum|F0|^2, sum|F1|^2)
    long long S;                        // samples per block
    long long mean_count;               // samples the byte sums were taken over (= S unless recording-wide sums are supplied)
    const uint8_t *halo0, *halo1;       // streaming mode: the (T-1) frames that precede frame 0 of block 0, right-aligned in
                                        // ceil((T-1)/F) super-frames of 4096 samples (16-byte aligned), or NULL
    int n_segs;
    int dc_remove;
    int P;                              // frames per block
    int Psf;                            // super-frames per block = ceil(P / F)  (= P for 4096 bins; see fx_fused4096s.cuh)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Unpack of byte J of a packed (I0,Q0,I1,Q1) word to the exact float (b - 128): PRMT drops the byte
// into the mantissa of 2^15 (ALU pipe), then one FADD2 per pair of components subtracts 2^15 + 128
// (FMA pipe).  (I2F.S8 Rd, Rs.BJ does it in one instruction but issues on the XU pipe at 16
// lanes/clk/SM -- measured 1.18 ms vs 0.95 ms per launch -- so it is not used.)
template <int J>
__device__ __forceinline__ float byte_to_magic(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7404u | (J << 4)));
}
constexpr float kMagic = 32768.0f + 128.0f;   // 2^15 + 128: f - kMagic = byte - 128, exact

// (I, Q) channel pairs of one packed word as exact floats (b - 128)
__device__ __forceinline__ void unpack_pairs(uint32_t w, float2 &pi, float2 &pq) {
    const float2 mg = f2(-kMagic, -kMagic);
    pi = f2add(f2(byte_to_magic<0>(w), byte_to_magic<2>(w)), mg);
    pq = f2add(f2(byte_to_magic<1>(w), byte_to_magic<3>(w)), mg);
}

__global__ void __launch_bounds__(NT, 1) fused_kernel(const Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int t = threadIdx.x;

    // ---- one-time per CTA: mbarriers, then the tables arrive by TMA bulk copy ---
    if (t == 0) {
        for (int s = 0; s <= RING; ++s) mbar_init(&sm.mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.mbar[RING], (uint32_t)(sizeof(sm.taps) + sizeof(sm.twA) + sizeof(sm.twB)));
        tma_load_1d(&sm.taps[0], prm.taps, (uint32_t)sizeof(sm.taps), &sm.mbar[RING]);
        tma_load_1d(&sm.twA[0][0], prm.twA, (uint32_t)sizeof(sm.twA), &sm.mbar[RING]);
        tma_load_1d(&sm.twB[0][0], prm.twB, (uint32_t)sizeof(sm.twB), &sm.mbar[RING]);
    }
    __syncthreads();
    bool tables_ready = false;

    const int k1B = t >> 4;    // after exchange 1: k1
    const int lo = t & 15;     // n3 (stage B) / k2 (stage C)
    uint32_t ring_cnt = 0;     // ingest counter: slot = cnt % RING, parity = (cnt / RING) & 1

    int already = 0;           // ingest items of the coming segment already issued by the previous one (thread 0)
    const int seg_end = prm.cta_first[blockIdx.x + 1];
    for (int seg = prm.cta_first[blockIdx.x]; seg < seg_end; ++seg) {
        const Segment sg = prm.segs[seg];
        const uint8_t *b0 = prm.iq0 + 2ll * prm.S * sg.block;
        const uint8_t *b1 = prm.iq1 + 2ll * prm.S * sg.block;

        // block means -> (mean - 128), negated, channel-packed
        float2 nmI, nmQ;
        if (prm.dc_remove) {
            const unsigned long long *su = prm.sums + 4ll * sg.block;
            const double inv = 1.0 / (double)prm.mean_count;
            nmI = f2((float)(128.0 - (double)su[0] * inv), (float)(128.0 - (double)su[2] * inv));
            nmQ = f2((float)(128.0 - (double)su[1] * inv), (float)(128.0 - (double)su[3] * inv));
        } else {
            nmI = f2(0.5f, 0.5f);     // x = (b - 127.5)/127.5
            nmQ = f2(0.5f, 0.5f);
        }

        // ingest frames g0 .. f0+nf-1 ; frames before f0 only fill the FIR history
        const int g0 = sg.f0 - (T - 1) > 0 ? sg.f0 - (T - 1) : 0;
        const int n_ing = sg.f0 + sg.nf - g0;
        // the segment after this one (same CTA): its first frames are prefetched during our last ones
        const int nseg = seg + 1;
        const bool have_next = nseg < seg_end && n_ing >= RING;
        Segment ng = sg;
        if (have_next) ng = prm.segs[nseg];
        const int ng0 = ng.f0 - (T - 1) > 0 ? ng.f0 - (T - 1) : 0;
        const int n_ing_next = have_next ? ng.f0 + ng.nf - ng0 : 0;
        const uint8_t *nb0 = prm.iq0 + 2ll * prm.S * ng.block;
        const uint8_t *nb1 = prm.iq1 + 2ll * prm.S * ng.block;
        if (t == 0) {
            const int pre = n_ing < RING ? n_ing : RING;
            for (int j = already; j < pre; ++j) {
                const uint32_t s = (ring_cnt + j) % RING;
                mbar_expect_tx(&sm.mbar[s], 2 * FRAME_BYTES);
                tma_load_1d(&sm.raw[s][0][0], b0 + (long long)(g0 + j) * FRAME_BYTES, FRAME_BYTES, &sm.mbar[s]);
                tma_load_1d(&sm.raw[s][1][0], b1 + (long long)(g0 + j) * FRAME_BYTES, FRAME_BYTES, &sm.mbar[s]);
            }
            already = 0;
        }

        uint32_t hist[T - 1][16];     // packed (I0,Q0,I1,Q1) of frames i-1, i-2, i-3
        float2 accx[16], acca[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            accx[j] = f2(0.f, 0.f);
            acca[j] = f2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < T - 1; ++k) hist[k][j] = 0u;
        }

        for (int j = 0; j < n_ing; ++j, ++ring_cnt) {
            const int fi = g0 + j;                  // block-relative frame index
            const uint32_t slot = ring_cnt % RING;
            mbar_wait(&sm.mbar[slot], (ring_cnt / RING) & 1u);

            uint32_t cur[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t a = sm.raw[slot][0][t + NT * r];
                const uint32_t b = sm.raw[slot][1][t + NT * r];
                cur[r] = __byte_perm(a, b, 0x5410);                 // (I0,Q0,I1,Q1)
            }

            const bool compute = fi >= sg.f0;
            C2 v[16];
            if (compute) {
                if (!tables_ready) {
                    mbar_wait(&sm.mbar[RING], 0);
                    tables_ready = true;
                }
                // zero-history semantics of channelize_poly: frame fi sees taps k <= fi only
                const int kmax = fi < T - 1 ? fi : T - 1;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    float4 tp = sm.taps[t + NT * r];
                    if (kmax < T - 1) {
                        tp.w = 0.f;
                        if (kmax < 2) tp.z = 0.f;
                        if (kmax < 1) tp.y = 0.f;
                    }
                    const float hs = (tp.x + tp.y) + (tp.z + tp.w);
                    float2 ar = f2muls(nmI, hs);
                    float2 ai = f2muls(nmQ, hs);
                    float2 pi, pq;
                    unpack_pairs(cur[r], pi, pq);
                    ar = f2fmas(pi, tp.x, ar);
                    ai = f2fmas(pq, tp.x, ai);
                    unpack_pairs(hist[0][r], pi, pq);
                    ar = f2fmas(pi, tp.y, ar);
                    ai = f2fmas(pq, tp.y, ai);
                    unpack_pairs(hist[1][r], pi, pq);
                    ar = f2fmas(pi, tp.z, ar);
                    ai = f2fmas(pq, tp.z, ai);
                    unpack_pairs(hist[2][r], pi, pq);
                    ar = f2fmas(pi, tp.w, ar);
                    ai = f2fmas(pq, tp.w, ai);
                    v[r] = {ar, ai};
                }
                // ---- stage A ------------------------------------------------
                dft16(v);
            }
            // all threads have consumed ring slot `slot` (and last frame's X2 reads are done)
            __syncthreads();
            if (t == 0) {
                // refill the slot just consumed with the ingest item RING ahead (possibly of the next segment)
                const int jn = j + RING;
                const uint8_t *s0 = nullptr, *s1 = nullptr;
                if (jn < n_ing) {
                    s0 = b0 + (long long)(g0 + jn) * FRAME_BYTES;
                    s1 = b1 + (long long)(g0 + jn) * FRAME_BYTES;
                } else if (jn - n_ing < n_ing_next) {
                    s0 = nb0 + (long long)(ng0 + jn - n_ing) * FRAME_BYTES;
                    s1 = nb1 + (long long)(ng0 + jn - n_ing) * FRAME_BYTES;
                    already = jn - n_ing + 1;
                }
                if (s0) {
                    mbar_expect_tx(&sm.mbar[slot], 2 * FRAME_BYTES);
                    tma_load_1d(&sm.raw[slot][0][0], s0, FRAME_BYTES, &sm.mbar[slot]);
                    tma_load_1d(&sm.raw[slot][1][0], s1, FRAME_BYTES, &sm.mbar[slot]);
                }
            }
            // rotate FIR history
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                hist[2][r] = hist[1][r];
                hist[1][r] = hist[0][r];
                hist[0][r] = cur[r];
            }
            if (!compute) continue;

            // ---- exchange 1: X[k1*256 + t] --------------------------------
            // register position 4g+b holds k1 = g + 4b; twiddles are fetched one group ahead
            {
                float2 tw[4], twn[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) tw[b] = sm.twA[4 * b][t];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < 3) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) twn[b] = sm.twA[g + 1 + 4 * b][t];
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int k1 = g + 4 * b;
                        C2 z = v[4 * g + b];
                        if (k1 != 0) z = cmuls(z, tw[b].x, tw[b].y);
                        sm.Xr[k1 * NT + t] = z.r;
                        sm.Xi[k1 * NT + t] = z.i;
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) tw[b] = twn[b];
                }
            }
            __syncthreads();
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                v[n2] = {sm.Xr[k1B * 256 + n2 * 16 + lo], sm.Xi[k1B * 256 + n2 * 16 + lo]};
            }
            // ---- stage B ---------------------------------------------------
            dft16(v);
            // exchange 2 (inside this half-warp's own 256-element region, XOR swizzle)
            // every lane of the warp must have finished reading X before anyone overwrites it
            __syncwarp();
            {
                float2 tw[4], twn[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) tw[b] = sm.twB[4 * b][lo];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < 3) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) twn[b] = sm.twB[g + 1 + 4 * b][lo];
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int k2 = g + 4 * b;
                        C2 z = v[4 * g + b];
                        if (k2 != 0) z = cmuls(z, tw[b].x, tw[b].y);
                        sm.Xr[k1B * 256 + k2 * 16 + (lo ^ k2)] = z.r;
                        sm.Xi[k1B * 256 + k2 * 16 + (lo ^ k2)] = z.i;
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) tw[b] = twn[b];
                }
            }
            __syncwarp();
#pragma unroll
            for (int n3 = 0; n3 < 16; ++n3) {
                v[n3] = {sm.Xr[k1B * 256 + lo * 16 + (n3 ^ lo)], sm.Xi[k1B * 256 + lo * 16 + (n3 ^ lo)]};
            }
            // ---- stage C + X-engine -----------------------------------------
            dft16(v);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float re0 = v[jj].r.x, re1 = v[jj].r.y, im0 = v[jj].i.x, im1 = v[jj].i.y;
                accx[jj].x = fmaf(re0, re1, fmaf(im0, im1, accx[jj].x));
                accx[jj].y = fmaf(im0, re1, fmaf(-re0, im1, accx[jj].y));
                acca[jj] = f2fma(v[jj].r, v[jj].r, f2fma(v[jj].i, v[jj].i, acca[jj]));
            }
        }

        // ---- segment epilogue: bins k1 + 16*k2 + 256*k3, staged through smem so the
        //      global writes are coalesced 16-byte stores ---------------------------
        __syncthreads();                       // every warp is done with X (exchange-2 reads)
        {
            float2 *xs = sm.Xr;                                // [0,4096): cross, [4096,8192): autos (= Xi)
            // element `bin` is staged at bin ^ ((bin >> 4) & 15): conflict-free for these scattered
            // stores (lanes differ in bits 4..7 of bin) and for the linear read-out below
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int bin = k1B + 16 * lo + 256 * perm16(jj);
                const int sw = bin ^ lo;                       // (bin >> 4) & 15 == lo
                xs[sw] = accx[jj];
                xs[N + sw] = acca[jj];
            }
            __syncthreads();
            float2 *px = prm.part_x + (long long)seg * N;
            float2 *pa = prm.part_a + (long long)seg * N;
#pragma unroll
            for (int q = 0; q < N / NT; ++q) {
                const int o = t + NT * q;
                const int sw = o ^ ((o >> 4) & 15);
                px[o] = xs[sw];
                pa[o] = xs[N + sw];
            }
            // the next frame's exchange-1 stores wait behind the barrier at the top of its iteration
        }
    }
}

}  // namespace fused4096
}  // namespace fx
