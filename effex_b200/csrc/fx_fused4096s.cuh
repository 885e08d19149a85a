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
// exchange planes: 16 tiles of 16 rows x 16 columns, rows padded to 17 elements.  Column-wise and
// row-wise accesses of a half-warp are both bank-conflict free, and every offset other than the
// thread's own (row or column) is a compile-time immediate -- no swizzle arithmetic per access.
constexpr int ROWP = 17;
constexpr int TILE = 16 * ROWP;
constexpr int NP = 16 * TILE;

struct __align__(16) SmemS {
    float2 Xr[2][NP];            // 2 x 34 KB exchange planes (re ch0, re ch1), double buffered
    float2 Xi[2][NP];            // 2 x 34 KB exchange planes (im ch0, im ch1)
    float4 twA[8][NT];           // row 2g+h: (W4096^(t*k1), W4096^(t*(k1+4))) for k1 = g + 8h  (one LDS.128 per pair)
    float4 twB[8][16];           // row 2g+h: (W256^(n3*k2), W256^(n3*(k2+4))) for k2 = g + 8h
    unsigned short raw[RING_S][2][N];
    unsigned long long mbar[RING_S + 1];
    uint32_t tmem_base;
    // TMA producer state of the current segment (written and read by thread 0 only; kept out of registers)
    const uint8_t *src0, *src1, *nsrc0, *nsrc1, *hsrc0, *hsrc1;
    int n_ing, n_ing_next, n_halo;
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

__global__ void __launch_bounds__(NT, 1) fused_kernel_stag(const Params prm) {
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
    for (int r = 0; r < 16; ++r) tmem_st4(tm_pts + 16 * r, prm.taps[t + NT * r]);
    tmem_wait_st();
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
        const bool use_halo = prm.halo0 != nullptr && sg.block == 0 && sg.f0 < T - 1;
        const int g0 = (use_halo || sg.f0 - (T - 1) > 0) ? sg.f0 - (T - 1) : 0;
        const int n_ing = sg.f0 + sg.nf - g0;
        if (t == 0) {
            // the segment after this one (same CTA): its first frames are prefetched during our last ones
            const int nseg = seg + 1;
            const bool have_next = nseg < seg_end && n_ing >= RING_S;
            Segment ng = sg;
            if (have_next) ng = prm.segs[nseg];
            const int ng0 = ng.f0 - (T - 1) > 0 ? ng.f0 - (T - 1) : 0;
            sm.src0 = b0 + (long long)g0 * FRAME_BYTES;        // ingest item j reads src + j*FRAME_BYTES ...
            sm.src1 = b1 + (long long)g0 * FRAME_BYTES;
            sm.n_halo = g0 < 0 ? -g0 : 0;                      // ... except the first n_halo items: halo frames
            sm.hsrc0 = prm.halo0 + (long long)(T - 1 + g0) * FRAME_BYTES;
            sm.hsrc1 = prm.halo1 + (long long)(T - 1 + g0) * FRAME_BYTES;
            sm.nsrc0 = prm.iq0 + 2ll * prm.S * ng.block + (long long)ng0 * FRAME_BYTES;
            sm.nsrc1 = prm.iq1 + 2ll * prm.S * ng.block + (long long)ng0 * FRAME_BYTES;
            sm.n_ing = n_ing;
            sm.n_ing_next = have_next ? ng.f0 + ng.nf - ng0 : 0;
            const int pre = n_ing < RING_S ? n_ing : RING_S;
            for (int j = already; j < pre; ++j) {
                const uint32_t s = (ring_cnt + j) % RING_S;
                mbar_expect_tx(&sm.mbar[s], 2 * FRAME_BYTES);
                const bool hal = j < sm.n_halo;
                tma_load_1d(&sm.raw[s][0][0], (hal ? sm.hsrc0 : sm.src0) + (long long)j * FRAME_BYTES, FRAME_BYTES, &sm.mbar[s]);
                tma_load_1d(&sm.raw[s][1][0], (hal ? sm.hsrc1 : sm.src1) + (long long)j * FRAME_BYTES, FRAME_BYTES, &sm.mbar[s]);
            }
            already = 0;
        }
        // refill of the ring slot used by ingest item j (called by thread 0 after the barrier that
        // follows every group's reads of that slot)
        auto refill = [&](int j, uint32_t slot) {
            const int jn = j + RING_S;
            const int ni = sm.n_ing;
            const uint8_t *s0 = nullptr, *s1 = nullptr;
            if (jn < ni) {
                const bool hal = jn < sm.n_halo;
                s0 = (hal ? sm.hsrc0 : sm.src0) + (long long)jn * FRAME_BYTES;
                s1 = (hal ? sm.hsrc1 : sm.src1) + (long long)jn * FRAME_BYTES;
            } else if (jn - ni < sm.n_ing_next) {
                s0 = sm.nsrc0 + (long long)(jn - ni) * FRAME_BYTES;
                s1 = sm.nsrc1 + (long long)(jn - ni) * FRAME_BYTES;
                already = jn - ni + 1;
            }
            if (s0) {
                mbar_expect_tx(&sm.mbar[slot], 2 * FRAME_BYTES);
                tma_load_1d(&sm.raw[slot][0][0], s0, FRAME_BYTES, &sm.mbar[slot]);
                tma_load_1d(&sm.raw[slot][1][0], s1, FRAME_BYTES, &sm.mbar[slot]);
            }
        };

        float2 accx[16], acca[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            accx[j] = f2(0.f, 0.f);
            acca[j] = f2(0.f, 0.f);
        }
        C2 v[16];
        // zero FIR state: a segment either starts a block (zero history is the reference's semantics)
        // or first re-ingests the T-1 frames before its first output frame
#pragma unroll
        for (int r = 0; r < 16; ++r) {
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
            uint32_t cur[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t a = sm.raw[slot][0][t + NT * r];
                const uint32_t b = sm.raw[slot][1][t + NT * r];
                cur[r] = __byte_perm(a, b, 0x5410);               // (I0,Q0,I1,Q1)
            }
            auto point = [&](int r, const float4 &tp, const float4 &z1, const float4 &z2, const float4 &z3) {
                const uint32_t w = cur[r];
                const float2 yr = f2add(f2add(f2(byte_to_magic<0>(w), byte_to_magic<2>(w)), mg), nmI);
                const float2 yi = f2add(f2add(f2(byte_to_magic<1>(w), byte_to_magic<3>(w)), mg), nmQ);
                v[r] = {f2fmas(yr, tp.x, f2(z1.x, z1.y)), f2fmas(yi, tp.x, f2(z1.z, z1.w))};
                const float2 n1r = f2fmas(yr, tp.y, f2(z2.x, z2.y)), n1i = f2fmas(yi, tp.y, f2(z2.z, z2.w));
                const float2 n2r = f2fmas(yr, tp.z, f2(z3.x, z3.y)), n2i = f2fmas(yi, tp.z, f2(z3.z, z3.w));
                const float2 n3r = f2muls(yr, tp.w), n3i = f2muls(yi, tp.w);
                tmem_st4(tm_pts + 16 * r + 4, make_float4(n1r.x, n1r.y, n1i.x, n1i.y));
                tmem_st4(tm_pts + 16 * r + 8, make_float4(n2r.x, n2r.y, n2i.x, n2i.y));
                tmem_st4(tm_pts + 16 * r + 12, make_float4(n3r.x, n3r.y, n3i.x, n3i.y));
            };
            // software pipeline: the tensor-memory loads of point r+1 are issued before point r is
            // computed, so its tcgen05.wait::ld finds them complete
            {
                float4 tp[2], z1[2], z2[2], z3[2];
                tmem_ld8(tm_pts, tp[0], z1[0]);
                tmem_ld8(tm_pts + 8, z2[0], z3[0]);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int c = r & 1, n = c ^ 1;
                    tmem_wait_ld(tp[c], z1[c], z2[c], z3[c]);
                    if (r + 1 < 16) {
                        tmem_ld8(tm_pts + 16 * (r + 1), tp[n], z1[n]);
                        tmem_ld8(tm_pts + 16 * (r + 1) + 8, z2[n], z3[n]);
                    }
                    point(r, tp[c], z1[c], z2[c], z3[c]);
                }
            }
            if (compute) {
                if (!tables_ready) {
                    mbar_wait(&sm.mbar[RING_S], 0);
                    tables_ready = true;
                }
                dft16(v);
                float2 *xr = sm.Xr[buf], *xi = sm.Xi[buf];
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
                        const int k1 = g + 4 * b;
                        C2 z = v[4 * g + b];
                        const float4 q = tq[b >> 1];
                        if (k1 != 0) z = (b & 1) ? cmuls(z, q.z, q.w) : cmuls(z, q.x, q.y);
                        sts_pair(&xr[k1 * TILE + k1B * ROWP + lo], z.r);
                        sts_pair(&xi[k1 * TILE + k1B * ROWP + lo], z.i);
                    }
                    tq[0] = tqn[0];
                    tq[1] = tqn[1];
                }
            }
        };

        // ---- stages B, C and the X-engine of the frame held in exchange buffer `buf` ---------------
        auto fft_rest = [&](int buf) {
            float2 *xr = sm.Xr[buf], *xi = sm.Xi[buf];
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) v[n2] = {xr[k1B * TILE + n2 * ROWP + lo], xi[k1B * TILE + n2 * ROWP + lo]};
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
                        sts_pair(&xr[k1B * TILE + k2 * ROWP + lo], z.r);
                        sts_pair(&xi[k1B * TILE + k2 * ROWP + lo], z.i);
                    }
                    tq[0] = tqn[0];
                    tq[1] = tqn[1];
                }
            }
            __syncwarp();
#pragma unroll
            for (int n3 = 0; n3 < 16; ++n3) v[n3] = {xr[k1B * TILE + lo * ROWP + n3], xi[k1B * TILE + lo * ROWP + n3]};
            dft16(v);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float re0 = v[jj].r.x, re1 = v[jj].r.y, im0 = v[jj].i.x, im1 = v[jj].i.y;
                accx[jj].x = fmaf(re0, re1, fmaf(im0, im1, accx[jj].x));
                accx[jj].y = fmaf(im0, re1, fmaf(-re0, im1, accx[jj].y));
                acca[jj] = f2fma(v[jj].r, v[jj].r, f2fma(v[jj].i, v[jj].i, acca[jj]));
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
                    if (q >= 0) fft_rest(q & 1);
                }
            }
            __syncthreads();
            if (t == 0 && s < n_ing) refill(s, (ring_cnt + (uint32_t)s) % RING_S);
        }
        ring_cnt += (uint32_t)n_ing;

        // ---- segment epilogue (both exchange buffers are free after the last barrier) ----------------
        {
            float2 *xs = &sm.Xr[0][0];                          // cross in Xr[0], autos in Xr[1]
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int bin = k1B + 16 * lo + 256 * perm16(jj);
                const int sw = bin ^ lo;
                sts_pair(&xs[sw], accx[jj]);
                sts_pair(&xs[N + sw], acca[jj]);
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
            __syncthreads();       // the next segment's first exchange stores must not overtake these reads
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(sm.tmem_base, 512);
}

}  // namespace fused4096
}  // namespace fx
