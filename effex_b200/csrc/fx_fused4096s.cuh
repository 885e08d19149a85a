// fx_fused4096s.cuh -- staggered variant of the fused kernel (same math, same thread/point/bin
// mapping as fx_fused4096.cuh; see that file for the dataflow).
//
// Why: in fused_kernel all 8 warps run the same phase at the same time, so the shared-memory
// bursts of the two exchanges and the FMA-heavy butterflies/FIR never overlap (measured: time ~
// compute-only time + exchange time).  Here the CTA is split into two warp groups that do the
// two halves of a frame's work in OPPOSITE order:
//     group 0 (warps 0-3):  FIR + stage A of frame q+1   then   stages B, C + X-engine of frame q
//     group 1 (warps 4-7):  stages B, C + X-engine of q  then   FIR + stage A of frame q+1
// Every scheduler holds one warp of each group (warp w and w+4), so at any time it has one
// ALU/FMA-heavy FIR warp and one exchange-heavy FFT warp to pick from.  Exchange 1 is double
// buffered (frame q+1 is written while frame q is read), which also removes one of the two CTA
// barriers per frame.  The FIR taps live in TENSOR MEMORY (each thread's 64 taps are private to
// it: tcgen05.st once, tcgen05.ld per frame) because the second exchange buffer takes the shared
// memory they used to occupy.
//
// The FIR runs in TRANSPOSED form: each incoming sample is unpacked ONCE and scattered into the
// partial sums of the four frames it contributes to,
//     out_i = t0*y_i + z1 ;  z1' = t1*y_i + z2 ;  z2' = t2*y_i + z3 ;  z3' = t3*y_i
// (identical arithmetic to the direct form up to float32 summation order; zero state at a block
// start IS channelize_poly's zero history).  The state z1..z3 (12 floats per point, 48 KB/warp pair)
// is thread-private, so it lives in tensor memory next to the taps: two tcgen05.ld.x8 and three
// tcgen05.st.x4 per point per frame, the loads issued one point ahead (measured tcgen05.ld 2.6 KB/clk/SM, tcgen05.st
// 285 B/clk/SM, tools/ubench_tmem.cu).  Compared with re-unpacking the three history frames this
// removes 12 of 16 PRMT and 4 of 8 FADD2 per point and the 48 history registers.
#pragma once
#include "fx_fused4096.cuh"

namespace fx {
namespace fused4096 {

constexpr int RING_S = 2;
// exchange planes: 16 tiles of 16 rows x 16 columns of 16-byte elements (one complex pair-of-channels
// value per LDS.128/STS.128), rows padded to 17 elements.  Column-wise and row-wise accesses of a
// quarter-warp are both bank-conflict free, and every offset other than the thread's own (row or column)
// is a compile-time immediate -- no swizzle arithmetic per access.
constexpr int ROWP = 17;
constexpr int TILE = 16 * ROWP;
constexpr int NP = 16 * TILE;

struct __align__(16) SmemS {
    float4 X[2][NP];             // 2 x 68 KB exchange planes of (re ch0, re ch1, im ch0, im ch1), double buffered
    float4 twA[8][NT];           // row 2g+h: stage-A twiddles of registers 4g+2h and 4g+2h+1 (one LDS.128 per pair)
    float4 twB[8][16];           // row 2g+h: W256^(n3*k2) of registers 4g+2h and 4g+2h+1, k2 = perm16(register)
    unsigned short raw[RING_S][2][N];
    unsigned long long mbar[RING_S + 1];
    uint32_t tmem_base;
    // TMA producer state of the current segment (written and read by thread 0 only; kept out of registers)
    const uint8_t *src0, *src1, *nsrc0, *nsrc1, *hsrc0, *hsrc1;
    int n_ing, n_ing_next, n_halo, g0, ng0;
};

// ---- tensor memory helpers (tcgen05; 32 lanes x 32-bit columns per warp quarter) -------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float4 &p, float4 &q) {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h)
                 : "r"(taddr));
    p = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
    q = make_float4(__uint_as_float(e), __uint_as_float(f), __uint_as_float(g), __uint_as_float(h));
}
// shared-memory store of one packed pair, spelled as two 32-bit registers: with a plain `*p = v` ptxas
// (12.9) copies every 64-bit FFMA2 result into one staging pair before its STS.64 (2 MOV per store,
// serialised on that pair's scoreboard) -- 32 stores per exchange, 6% of the kernel's time
__device__ __forceinline__ void sts_c2(float4 *p, const C2 &z) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(z.r.x),
                 "f"(z.r.y), "f"(z.i.x), "f"(z.i.y)
                 : "memory");
}
__device__ __forceinline__ C2 lds_c2(const float4 *p) {
    const float4 q = *p;
    return {f2(q.x, q.y), f2(q.z, q.w)};
}
__device__ __forceinline__ void sts_pair(float2 *p, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y) : "memory");
}
// the loaded registers are routed through the wait so that no use can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld(float4 &a, float4 &b, float4 &c, float4 &d) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w), "+f"(c.x),
                   "+f"(c.y), "+f"(c.z), "+f"(c.w), "+f"(d.x), "+f"(d.y), "+f"(d.z), "+f"(d.w)::"memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- frame lengths below 4096 ------------------------------------------------------------------
// LOGF > 0: a ring slot holds one SUPER-FRAME of 4096 consecutive samples = F = 2^LOGF consecutive
// frames of NL = 4096/F bins.  A thread still owns the 16 samples t + 256 r; they are now RP = 16/F
// positions p' = t + 256 r' of each of the F frames (r = f*RP + r').  What changes:
//   FIR      per position r' the state is loaded once and carried in registers through the F frames;
//   stage A  F DFTs of RP points (instead of one of 16), twiddle W_NL^(t*k1'), tile = f*RP + k1';
//   B, C     unchanged: every tile is still one 256-point transform over t;
//   bins     tile f*RP + k1' holds bins k1' + RP*(k2 + 16*k3) of frame f; the F frame slots are added
//            in the segment epilogue.
// Segments, the ring and the TMA copies count super-frames; the last super-frame of a block may hold
// fewer than F frames (P mod F): its copy is shortened and its missing frames are not accumulated.
__host__ __device__ constexpr int perm_rp(int RP, int jj) {
    return RP == 16 ? perm16(jj) : RP == 8 ? (jj < 4 ? 2 * jj : 2 * (jj - 4) + 1) : jj;
}
// exchange-1 tile written by register j of stage A
__host__ __device__ constexpr int row_of(int LOGF, int j) {
    return (j / (16 >> LOGF)) * (16 >> LOGF) + perm_rp(16 >> LOGF, j % (16 >> LOGF));
}
// forward 8-point DFT in place; v[m] = Y[2m], v[4+m] = Y[2m+1]
__device__ __forceinline__ void dft8(C2 *v) {
    constexpr float R2 = 0.70710678118654752f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const C2 s = cadd(v[n], v[n + 4]), d = csub(v[n], v[n + 4]);
        v[n] = s;
        v[n + 4] = d;
    }
    v[5] = {f2muls(f2add(v[5].r, v[5].i), R2), f2muls(f2sub(v[5].i, v[5].r), R2)};      // * W8^1
    v[6] = {v[6].i, f2neg(v[6].r)};                                                      // * W8^2 = -i
    v[7] = {f2muls(f2sub(v[7].i, v[7].r), R2), f2muls(f2add(v[7].r, v[7].i), -R2)};     // * W8^3
    radix4(v[0], v[1], v[2], v[3]);
    radix4(v[4], v[5], v[6], v[7]);
}
template <int RP>
__device__ __forceinline__ void dft_small(C2 *v) {
    if constexpr (RP == 8) dft8(v);
    if constexpr (RP == 4) radix4(v[0], v[1], v[2], v[3]);
    if constexpr (RP == 2) {
        const C2 s = cadd(v[0], v[1]), d = csub(v[0], v[1]);
        v[0] = s;
        v[1] = d;
    }
}

// 232 registers x 256 threads leave room on the SM for two CTAs of the byte-sum pre-pass (128 threads x
// 32 registers): the pre-pass of call k+1 (HBM-bound) then runs underneath this kernel (FP32-bound) of
// call k instead of after it.  ptxas needs 219 registers at this cap and does not spill.
#ifndef FX_MAXNREG
#define FX_MAXNREG 232
#endif
// AUTOS = false drops the auto-power accumulators (32 registers, 32 FFMA2 per frame and thread) and the
// part_a stores: fx_process without d_auto0/d_auto1 and without accumulators never reads them.
template <int LOGF, bool AUTOS>
__global__ void __maxnreg__(FX_MAXNREG) fused_kernel_stag(const Params prm) {
    constexpr int F = 1 << LOGF;        // frames per super-frame
    constexpr int RP = 16 >> LOGF;      // positions of one frame held by a thread = stage-A radix
    constexpr int NL = N >> LOGF;       // frame length = number of bins
    constexpr int HSF = (T - 1 + F - 1) / F;   // super-frames of history re-ingested at a segment start
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemS &sm = *reinterpret_cast<SmemS *>(smem_raw);
    const int t = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);   // broadcast: lets the compiler keep it (and TMEM addresses) uniform
    const int grp = warp >> 2;          // 0: FIR first, 1: FFT first

    if (t == 0) {
        for (int s = 0; s <= RING_S; ++s) mbar_init(&sm.mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.mbar[RING_S], (uint32_t)(sizeof(sm.twA) + sizeof(sm.twB)));
        tma_load_1d(&sm.twA[0][0], prm.twAp, (uint32_t)sizeof(sm.twA), &sm.mbar[RING_S]);
        tma_load_1d(&sm.twB[0][0], prm.twBp, (uint32_t)sizeof(sm.twB), &sm.mbar[RING_S]);
    }
    // all 512 columns: warps w and w+4 share a lane quarter, 256 columns each =
    // 16 points x [4 taps | z1 | z2 | z3 (4 floats each: re ch0, re ch1, im ch0, im ch1)]
    if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_pts = sm.tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * grp);
#pragma unroll
    for (int r = 0; r < RP; ++r) tmem_st4(tm_pts + 16 * r, prm.taps[t + NT * r]);
    tmem_wait_st();
    // frames in the last super-frame of a block, and its copy size per channel
    const int tail_valid = prm.P - F * (prm.Psf - 1);
    const uint32_t tail_bytes = (uint32_t)tail_valid * (uint32_t)(2 * NL);
    const int fslot = (t >> 4) >> (4 - LOGF);          // frame slot of this thread's tile in stages B, C
    bool tables_ready = false;

    const int k1B = t >> 4;
    const int lo = t & 15;
    uint32_t ring_cnt = 0;
    int already = 0;

    const int seg_end = prm.cta_first[blockIdx.x + 1];
    for (int seg = prm.cta_first[blockIdx.x]; seg < seg_end; ++seg) {
        const Segment sg = prm.segs[seg];
        const uint8_t *b0 = prm.iq0 + 2ll * prm.S * sg.block;
        const uint8_t *b1 = prm.iq1 + 2ll * prm.S * sg.block;
        float2 nmI, nmQ;
        if (prm.dc_remove) {
            const unsigned long long *su = prm.sums + 4ll * sg.block;
            const double inv = 1.0 / (double)prm.mean_count;
            nmI = f2((float)(128.0 - (double)su[0] * inv), (float)(128.0 - (double)su[2] * inv));
            nmQ = f2((float)(128.0 - (double)su[1] * inv), (float)(128.0 - (double)su[3] * inv));
        } else {
            nmI = f2(0.5f, 0.5f);
            nmQ = f2(0.5f, 0.5f);
        }
        // first ingested frame: T-1 frames before the first output frame; before frame 0 of block 0 only
        // when the caller supplied the preceding samples (streaming mode), otherwise zero history
        // (segments count super-frames; the host hands over the T-1 halo frames right-aligned in HSF super-frames)
        const bool use_halo = prm.halo0 != nullptr && sg.block == 0 && sg.f0 < HSF;
        const int g0 = (use_halo || sg.f0 - HSF > 0) ? sg.f0 - HSF : 0;
        const int n_ing = sg.f0 + sg.nf - g0;
        if (t == 0) {
            // the segment after this one (same CTA): its first frames are prefetched during our last ones
            const int nseg = seg + 1;
            const bool have_next = nseg < seg_end && n_ing >= RING_S;
            Segment ng = sg;
            if (have_next) ng = prm.segs[nseg];
            const int ng0 = ng.f0 - HSF > 0 ? ng.f0 - HSF : 0;
            sm.src0 = b0 + (long long)g0 * FRAME_BYTES;        // ingest item j reads src + j*FRAME_BYTES ...
            sm.src1 = b1 + (long long)g0 * FRAME_BYTES;
            sm.n_halo = g0 < 0 ? -g0 : 0;                      // ... except the first n_halo items: halo frames
            sm.hsrc0 = prm.halo0 + (long long)(HSF + g0) * FRAME_BYTES;      // halo buffer: HSF super-frames
            sm.hsrc1 = prm.halo1 + (long long)(HSF + g0) * FRAME_BYTES;
            sm.nsrc0 = prm.iq0 + 2ll * prm.S * ng.block + (long long)ng0 * FRAME_BYTES;
            sm.nsrc1 = prm.iq1 + 2ll * prm.S * ng.block + (long long)ng0 * FRAME_BYTES;
            sm.g0 = g0;
            sm.ng0 = ng0;
            sm.n_ing = n_ing;
            sm.n_ing_next = have_next ? ng.f0 + ng.nf - ng0 : 0;
            const int pre = n_ing < RING_S ? n_ing : RING_S;
            for (int j = already; j < pre; ++j) {
                const uint32_t s = (ring_cnt + j) % RING_S;
                const uint32_t nb = (g0 + j == prm.Psf - 1) ? tail_bytes : (uint32_t)FRAME_BYTES;
                mbar_expect_tx(&sm.mbar[s], 2 * nb);
                const bool hal = j < sm.n_halo;
                tma_load_1d(&sm.raw[s][0][0], (hal ? sm.hsrc0 : sm.src0) + (long long)j * FRAME_BYTES, nb, &sm.mbar[s]);
                tma_load_1d(&sm.raw[s][1][0], (hal ? sm.hsrc1 : sm.src1) + (long long)j * FRAME_BYTES, nb, &sm.mbar[s]);
            }
            already = 0;
        }
        // refill of the ring slot used by ingest item j (called by thread 0 after the barrier that
        // follows every group's reads of that slot)
        auto refill = [&](int j, uint32_t slot) {
            const int jn = j + RING_S;
            const int ni = sm.n_ing;
            const uint8_t *s0 = nullptr, *s1 = nullptr;
            uint32_t nb = (uint32_t)FRAME_BYTES;
            if (jn < ni) {
                if (sm.g0 + jn == prm.Psf - 1) nb = tail_bytes;
                const bool hal = jn < sm.n_halo;
                s0 = (hal ? sm.hsrc0 : sm.src0) + (long long)jn * FRAME_BYTES;
                s1 = (hal ? sm.hsrc1 : sm.src1) + (long long)jn * FRAME_BYTES;
            } else if (jn - ni < sm.n_ing_next) {
                s0 = sm.nsrc0 + (long long)(jn - ni) * FRAME_BYTES;
                s1 = sm.nsrc1 + (long long)(jn - ni) * FRAME_BYTES;
                if (sm.ng0 + jn - ni == prm.Psf - 1) nb = tail_bytes;
                already = jn - ni + 1;
            }
            if (s0) {
                mbar_expect_tx(&sm.mbar[slot], 2 * nb);
                tma_load_1d(&sm.raw[slot][0][0], s0, nb, &sm.mbar[slot]);
                tma_load_1d(&sm.raw[slot][1][0], s1, nb, &sm.mbar[slot]);
            }
        };

        float2 accx[16], acca[AUTOS ? 16 : 1];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            accx[j] = f2(0.f, 0.f);
            if (AUTOS) acca[j] = f2(0.f, 0.f);
        }
        C2 v[16];
        // zero FIR state: a segment either starts a block (zero history is the reference's semantics)
        // or first re-ingests the T-1 frames before its first output frame
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            tmem_st4(tm_pts + 16 * r + 4, make_float4(0.f, 0.f, 0.f, 0.f));
            tmem_st4(tm_pts + 16 * r + 8, make_float4(0.f, 0.f, 0.f, 0.f));
            tmem_st4(tm_pts + 16 * r + 12, make_float4(0.f, 0.f, 0.f, 0.f));
        }

        // ---- ingest item j: raw -> history; if `compute`, FIR + stage A -> exchange buffer `buf` ----
        auto fir_stage_a = [&](int j, bool compute, int buf) {
            const uint32_t cnt = ring_cnt + (uint32_t)j;          // ring_cnt = counter at segment start
            const uint32_t slot = cnt % RING_S;
            mbar_wait(&sm.mbar[slot], (cnt / RING_S) & 1u);
            tmem_wait_st();                                       // last frame's state stores have landed
            const float2 mg = f2(-kMagic, -kMagic);
            // one point: tp = taps, z1..z3 = state (re ch0, re ch1, im ch0, im ch1); out = t0*y + z1
            // raw bytes of this frame's 16 points first: their shared-memory latency is paid once, up front
            // (the two channels' words stay separate: merging them into one (I0,Q0,I1,Q1) word costs a PRMT per point
            //  and saves 16 registers the kernel does not need -- 578 -> 574.5 us per launch)
            uint32_t cura[16], curb[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                cura[r] = sm.raw[slot][0][t + NT * r];
                curb[r] = sm.raw[slot][1][t + NT * r];
            }
            // position rp of every frame f of the super-frame: sample r = f*RP + rp, taps tp, state z1..z3
            // (re ch0, re ch1, im ch0, im ch1) carried through the F frames; out = t0*y + z1
            auto position = [&](int rp, const float4 &tp, const float4 &z1, const float4 &z2, const float4 &z3) {
                float2 s1r = f2(z1.x, z1.y), s1i = f2(z1.z, z1.w);
                float2 s2r = f2(z2.x, z2.y), s2i = f2(z2.z, z2.w);
                float2 s3r = f2(z3.x, z3.y), s3i = f2(z3.z, z3.w);
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int r = f * RP + rp;
                    const uint32_t wa = cura[r], wb = curb[r];              // (I, Q, -, -) of channel 0 and of channel 1
                    const float2 yr = f2add(f2add(f2(byte_to_magic<0>(wa), byte_to_magic<0>(wb)), mg), nmI);
                    const float2 yi = f2add(f2add(f2(byte_to_magic<1>(wa), byte_to_magic<1>(wb)), mg), nmQ);
                    v[r] = {f2fmas(yr, tp.x, s1r), f2fmas(yi, tp.x, s1i)};
                    s1r = f2fmas(yr, tp.y, s2r);
                    s1i = f2fmas(yi, tp.y, s2i);
                    s2r = f2fmas(yr, tp.z, s3r);
                    s2i = f2fmas(yi, tp.z, s3i);
                    s3r = f2muls(yr, tp.w);
                    s3i = f2muls(yi, tp.w);
                }
                tmem_st4(tm_pts + 16 * rp + 4, make_float4(s1r.x, s1r.y, s1i.x, s1i.y));
                tmem_st4(tm_pts + 16 * rp + 8, make_float4(s2r.x, s2r.y, s2i.x, s2i.y));
                tmem_st4(tm_pts + 16 * rp + 12, make_float4(s3r.x, s3r.y, s3i.x, s3i.y));
            };
            // software pipeline: the tensor-memory loads of position rp+1 are issued before position rp
            // is computed, so its tcgen05.wait::ld finds them complete
            {
                float4 tp[2], z1[2], z2[2], z3[2];
                tmem_ld8(tm_pts, tp[0], z1[0]);
                tmem_ld8(tm_pts + 8, z2[0], z3[0]);
#pragma unroll
                for (int rp = 0; rp < RP; ++rp) {
                    const int c = rp & 1, n = c ^ 1;
                    tmem_wait_ld(tp[c], z1[c], z2[c], z3[c]);
                    if (rp + 1 < RP) {
                        tmem_ld8(tm_pts + 16 * (rp + 1), tp[n], z1[n]);
                        tmem_ld8(tm_pts + 16 * (rp + 1) + 8, z2[n], z3[n]);
                    }
                    position(rp, tp[c], z1[c], z2[c], z3[c]);
                }
            }
            if (compute) {
                if (!tables_ready) {
                    mbar_wait(&sm.mbar[RING_S], 0);
                    tables_ready = true;
                }
                if constexpr (RP == 16) {
                    dft16(v);
                } else {
#pragma unroll
                    for (int f = 0; f < F; ++f) dft_small<RP>(&v[f * RP]);
                }
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
                        const int row = row_of(LOGF, 4 * g + b);      // tile = frame slot * RP + k1'
                        C2 z = v[4 * g + b];
                        const float4 q = tq[b >> 1];
                        if (row % RP != 0) z = (b & 1) ? cmuls(z, q.z, q.w) : cmuls(z, q.x, q.y);
                        sts_c2(&xx[row * TILE + k1B * ROWP + lo], z);
                    }
                    tq[0] = tqn[0];
                    tq[1] = tqn[1];
                }
            }
        };

        // ---- stages B, C and the X-engine of the frame held in exchange buffer `buf` ---------------
        auto fft_rest = [&](int buf, int nv) {
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
                        C2 z = v[4 * g + b];
                        const float4 q = tq[b >> 1];
                        if (k2 != 0) z = (b & 1) ? cmuls(z, q.z, q.w) : cmuls(z, q.x, q.y);
                        sts_c2(&xx[k1B * TILE + k2 * ROWP + lo], z);
                    }
                    tq[0] = tqn[0];
                    tq[1] = tqn[1];
                }
            }
            __syncwarp();
#pragma unroll
            for (int n3 = 0; n3 < 16; ++n3) v[n3] = lds_c2(&xx[k1B * TILE + lo * ROWP + n3]);
            dft16(v);
            if (LOGF == 0 || fslot < nv) {        // frame slots beyond a block's last frame hold no data
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const float re0 = v[jj].r.x, re1 = v[jj].r.y, im0 = v[jj].i.x, im1 = v[jj].i.y;
                    accx[jj].x = fmaf(re0, re1, fmaf(im0, im1, accx[jj].x));
                    accx[jj].y = fmaf(im0, re1, fmaf(-re0, im1, accx[jj].y));
                    if (AUTOS) acca[jj] = f2fma(v[jj].r, v[jj].r, f2fma(v[jj].i, v[jj].i, acca[jj]));
                }
            }
        };

        // step s: FIR/stage A of ingest item s (history-only while s < j0) and stages B, C + X-engine of
        // the frame whose stage A finished in step s-1, in group-dependent order; one CTA barrier per step
        const int j0 = sg.f0 - g0;
#pragma unroll 1
        for (int s = 0; s <= n_ing; ++s) {
            const int q = s - j0 - 1;                 // frame (segment-relative) whose exchange buffer is ready
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (half == grp) {
                    if (s < n_ing) fir_stage_a(s, s >= j0, (s - j0) & 1);
                } else {
                    if (q >= 0) fft_rest(q & 1, (sg.f0 + q == prm.Psf - 1) ? tail_valid : F);
                }
            }
            __syncthreads();
            if (t == 0 && s < n_ing) refill(s, (ring_cnt + (uint32_t)s) % RING_S);
        }
        ring_cnt += (uint32_t)n_ing;

        // ---- segment epilogue (both exchange buffers are free after the last barrier) ----------------
        {
            float2 *xs = reinterpret_cast<float2 *>(&sm.X[0][0]);   // cross in xs[0..N), autos in xs[N..2N)
            // staging index = frame slot * NL + bin; bin = k1' + RP*(k2 + 16*k3) with k2 = lo, k3 = perm16(jj)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int idx = fslot * NL + (k1B & (RP - 1)) + RP * lo + 16 * RP * perm16(jj);
                const int sw = idx ^ ((idx >> 4) & 15);
                sts_pair(&xs[sw], accx[jj]);
                if (AUTOS) sts_pair(&xs[N + sw], acca[jj]);
            }
            __syncthreads();
            float2 *px = prm.part_x + (long long)seg * NL;
            float2 *pa = prm.part_a + (long long)seg * NL;
#pragma unroll
            for (int q = 0; q < NL / NT; ++q) {
                const int o = t + NT * q;
                float2 sx = f2(0.f, 0.f), sa = f2(0.f, 0.f);
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int idx = f * NL + o;
                    const int sw = idx ^ ((idx >> 4) & 15);
                    sx = f2add(sx, xs[sw]);
                    if (AUTOS) sa = f2add(sa, xs[N + sw]);
                }
                px[o] = sx;
                if (AUTOS) pa[o] = sa;
            }
            __syncthreads();       // the next segment's first exchange stores must not overtake these reads
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace fused4096
}  // namespace fx
