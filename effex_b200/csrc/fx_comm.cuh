// fx_comm.cuh -- the one collective of the FX path, done by our own kernels over NVLink peer memory.
//
// What is reduced is tiny (the float64 accumulators of one integration: 4N+1 doubles = 128 KB at 4096 bins;
// the 2n-point lag cross-spectrum: 4 MB), so the cost of a library collective is its launches, its
// rendezvous on the step's critical path and the SMs its kernel takes from the FP32-bound fused kernel --
// not bandwidth.  Here every rank owns a MAILBOX in its HBM (cudaMalloc, exported with cudaIpcGetMemHandle
// and mapped by every peer): two parities x world slots + per-CTA flags.  A reduce to `root` is
//
//   every rank (root too):  PUSH   its contribution straight into slot[parity][rank] of the ROOT's mailbox
//                                   (plain 16-byte stores to the peer mapping, i.e. NVLink writes), then one
//                                   st.release.sys of the epoch number per CTA into flag[parity][rank][cta];
//   root only:              FOLD   CTA c waits (ld.acquire.sys) until flag[parity][r][c] == epoch for every r,
//                                   adds the world slots IN RANK ORDER (deterministic float64 sums) into the
//                                   destination, and publishes done[c] = epoch.
//
// The push of an integration is fused into the kernel that folds the per-block partial sums
// (integrate_push_kernel below = integrate_stage2_kernel + push), so a non-root rank adds NO launch to its
// step.  The root issues the fold of epoch e one collective call LATER (at the end of step e+1, or in
// fx_comm_fence / fx_sync): the peers' pushes have had a whole step to land, so the fold does not wait and
// no rank's step ever stalls on another rank -- the ranks stay decoupled, as without a collective.
// Back-pressure: a slot of parity p is rewritten at epoch e+2; the pusher first waits for done[c] >= e.
// Epochs advance in lock step on all ranks (collective call order), as with any collective.
// All spins carry a clock64() deadline and raise an error word instead of hanging the GPU.
#pragma once
#include "fx_common.cuh"

namespace fx {
namespace comm {

constexpr int kThreads = 256;
constexpr int kMaxWorld = 16;
constexpr unsigned int kErrTimeoutPush = 1u, kErrTimeoutFold = 2u;

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until *p >= want (epochs are monotonic); false on timeout
__device__ __forceinline__ bool spin_until(const unsigned int *p, unsigned int want, long long timeout_cycles) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - want) < 0) {
        if (clock64() - t0 > timeout_cycles) return false;
        __nanosleep(100);
    }
    return true;
}

// element range of CTA c when n elements are cut into gridDim.x contiguous chunks (multiples of kThreads)
__device__ __forceinline__ void cta_range(size_t n, size_t &lo, size_t &hi) {
    size_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + kThreads - 1) / kThreads * kThreads;
    lo = (size_t)blockIdx.x * per;
    hi = lo + per < n ? lo + per : n;
    if (lo > n) lo = n;
}

struct PushTarget {
    void *slot;                  // root's slot[parity][my rank] (peer mapping, or local on the root)
    unsigned int *flag;          // root's flag[parity][my rank][gridDim.x]
    const unsigned int *done;    // root's done[gridDim.x]
    unsigned int *err;           // my own error word
    unsigned int epoch;
    long long timeout_cycles;
};

__device__ __forceinline__ void push_prologue(const PushTarget &t) {
    // the slot still holds epoch-2's contribution until the root has folded it
    if (threadIdx.x == 0 && t.epoch > 2) {
        if (!spin_until(t.done + blockIdx.x, t.epoch - 2, t.timeout_cycles)) atomicOr(t.err, kErrTimeoutPush);
    }
    __syncthreads();
}
__device__ __forceinline__ void push_epilogue(const PushTarget &t) {
    __syncthreads();                                  // the CTA's stores happen-before thread 0's release
    if (threadIdx.x == 0) st_release_sys(t.flag + blockIdx.x, t.epoch);
}

// generic contribution: n elements of T from src into the root's slot
template <typename T>
__global__ void __launch_bounds__(kThreads) push_kernel(const T *__restrict__ src, size_t n, PushTarget t) {
    push_prologue(t);
    size_t lo, hi;
    cta_range(n, lo, hi);
    T *slot = reinterpret_cast<T *>(t.slot);
    for (size_t i = lo + threadIdx.x; i < hi; i += kThreads) slot[i] = src[i];
    push_epilogue(t);
}

// integrate_stage2_kernel fused with the push: the G per-group float64 partial sums of this call
// (scratch[g][ 2c, 2c+1 | 2N + c | 3N + c ]) are folded and written straight into the root's slot as
// [acc_x (2N) | acc_a0 (N) | acc_a1 (N) | frames (1)] -- the layout of FxEngine.new_accumulators()["flat"].
__global__ void __launch_bounds__(kThreads) integrate_push_kernel(const double *__restrict__ scratch, int N, int G,
                                                                  double frames, PushTarget t) {
    push_prologue(t);
    const size_t n = 4 * (size_t)N + 1;
    size_t lo, hi;
    cta_range(n, lo, hi);
    double *slot = reinterpret_cast<double *>(t.slot);
    for (size_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        double v = frames;
        if (i < 4 * (size_t)N) {
            v = 0;
            for (int g = 0; g < G; ++g) v += scratch[(size_t)g * 4 * N + i];
        }
        slot[i] = v;
    }
    push_epilogue(t);
}

// root: dst[i] (+)= sum over ranks, in rank order, of slot[r][i]
template <typename T>
__global__ void __launch_bounds__(kThreads) fold_kernel(T *__restrict__ dst, size_t n, const void *slots0,
                                                        size_t slot_stride_bytes, const unsigned int *flags0,
                                                        int flag_stride, int world, unsigned int *done,
                                                        unsigned int *err, unsigned int epoch, int accumulate,
                                                        long long timeout_cycles) {
    if ((int)threadIdx.x < world) {
        if (!spin_until(flags0 + (size_t)threadIdx.x * flag_stride + blockIdx.x, epoch, timeout_cycles))
            atomicOr(err, kErrTimeoutFold);
    }
    __syncthreads();
    size_t lo, hi;
    cta_range(n, lo, hi);
    for (size_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        T v = accumulate ? dst[i] : T(0);
        for (int r = 0; r < world; ++r)
            v += __ldcg(reinterpret_cast<const T *>(reinterpret_cast<const char *>(slots0) + (size_t)r * slot_stride_bytes) + i);
        dst[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(done + blockIdx.x, epoch);
}

// ---- the same push / fold as the TAIL of the kernel that produces the per-group partial sums -------------
// finalize_integrate_kernel / integrate_stage1_kernel run a grid (N/256 bin tiles, G groups); the LAST CTA of
// a bin tile to finish (ticket counter) folds the tile's G partials and, in one go,
//   LOCAL  adds them to the caller's accumulators                       (fx_process_acc, fx_integrate)
//   PUSH   writes them into the root's slot and releases the flags       (fx_process_reduce, any rank)
//   +FOLD  on the root: also adds the world slots of the PREVIOUS epoch into the root's accumulators
// so a step with accumulators and the cross-GPU reduce launches exactly the kernels of a step without.
// A bin tile t (natural bins [256t, 256t+256)) covers these 256-element ranges of the flat accumulator
// layout [x 2N | a0 N | a1 N | frames 1], i.e. these CTA indices of push_kernel/fold_kernel over 4N+1
// elements (so the two forms interoperate on the same flags): 2t, 2t+1, N/128 + t, 3N/256 + t, and N/64
// (the frame count) for tile 0.
struct IntegrateTail {
    int mode;                    // 0: local accumulators, 1: push
    int autos;                   // a0/a1 parts carried
    int *counters;               // [N/256] tickets, zero between calls
    double *acc_x, *acc_a0, *acc_a1, *acc_frames;      // mode 0
    PushTarget push;             // mode 1
    // root only: fold of the pending epoch (fold_epoch != 0)
    unsigned int fold_epoch;
    double *fold_dst;            // flat [4N+1]
    const void *fold_slots;      // slot[parity(fold_epoch)][0]
    size_t slot_stride_bytes;
    const unsigned int *fold_flags;   // flag[parity(fold_epoch)][0][0]
    int flag_stride, world;
    unsigned int *done;
};

__device__ __forceinline__ int tail_range_cta(int k, int tile, int N) {
    return k == 0 ? 2 * tile : k == 1 ? 2 * tile + 1 : k == 2 ? N / 128 + tile : k == 3 ? 3 * N / 256 + tile : N / 64;
}

// called by every thread of a 256-thread CTA after it has written scratch[g][...] for bin tile `tile`
// (natural order); c = this thread's natural bin (256*tile + something), G = gridDim.y
__device__ __forceinline__ void integrate_tail(const IntegrateTail &t, const double *scratch, int N, int G, double frames,
                                               int tile, int c, bool valid) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(t.counters + tile, 1) == G - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) t.counters[tile] = 0;
    __threadfence();
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    if (valid) {
        // G <= 64 group sums per value, added in group order; 8 independent loads in flight
        for (int g0 = 0; g0 < G; g0 += 8) {
            double2 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                x[u] = __ldcg(reinterpret_cast<const double2 *>(scratch + (size_t)(g0 + u < G ? g0 + u : g0) * 4 * N + 2 * c));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (g0 + u < G) { v0 += x[u].x; v1 += x[u].y; }
        }
        if (t.autos) {
            for (int g0 = 0; g0 < G; g0 += 8) {
                double p[8], q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double *o = scratch + (size_t)(g0 + u < G ? g0 + u : g0) * 4 * N;
                    p[u] = __ldcg(o + 2 * N + c);
                    q[u] = __ldcg(o + 3 * N + c);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (g0 + u < G) { v2 += p[u]; v3 += q[u]; }
            }
        }
    }
    const int nr = tile == 0 ? 5 : 4;          // flag ranges of this tile
    if (t.mode == 0) {
        if (valid) {
            t.acc_x[2 * c] += v0; t.acc_x[2 * c + 1] += v1;
            if (t.autos) { t.acc_a0[c] += v2; t.acc_a1[c] += v3; }
        }
        if (tile == 0 && threadIdx.x == 0 && t.acc_frames) *t.acc_frames += frames;
    } else {
        if ((int)threadIdx.x < nr && t.push.epoch > 2) {
            if (!spin_until(t.push.done + tail_range_cta(threadIdx.x, tile, N), t.push.epoch - 2, t.push.timeout_cycles))
                atomicOr(t.push.err, kErrTimeoutPush);
        }
        __syncthreads();
        double *slot = reinterpret_cast<double *>(t.push.slot);
        if (valid) {
            *reinterpret_cast<double2 *>(slot + 2 * c) = make_double2(v0, v1);
            // cross-only handles never write the auto-power parts of a slot: they stay zero from fx_comm_export
            if (t.autos) { slot[2 * N + c] = v2; slot[3 * N + c] = v3; }
        }
        if (tile == 0 && threadIdx.x == 0) slot[4 * (size_t)N] = frames;
        __syncthreads();
        if ((int)threadIdx.x < nr) st_release_sys(t.push.flag + tail_range_cta(threadIdx.x, tile, N), t.push.epoch);
    }
    if (t.fold_epoch) {
        // root: the world slots of the previous epoch, in rank order
        if ((int)threadIdx.x < nr * t.world) {
            const int r = threadIdx.x / nr, k = threadIdx.x % nr;
            if (!spin_until(t.fold_flags + (size_t)r * t.flag_stride + tail_range_cta(k, tile, N), t.fold_epoch,
                            t.push.timeout_cycles))
                atomicOr(t.push.err, kErrTimeoutFold);
        }
        __syncthreads();
        double *d = t.fold_dst;
        if (valid) {
            double2 ax = *reinterpret_cast<double2 *>(d + 2 * c);
            double a0 = 0, a1 = 0;
            if (t.autos) { a0 = d[2 * N + c]; a1 = d[3 * N + c]; }
            for (int r = 0; r < t.world; ++r) {
                const double *sl = reinterpret_cast<const double *>(reinterpret_cast<const char *>(t.fold_slots) + (size_t)r * t.slot_stride_bytes);
                const double2 x = __ldcg(reinterpret_cast<const double2 *>(sl + 2 * c));
                ax.x += x.x; ax.y += x.y;
                if (t.autos) { a0 += __ldcg(sl + 2 * N + c); a1 += __ldcg(sl + 3 * N + c); }
            }
            *reinterpret_cast<double2 *>(d + 2 * c) = ax;
            if (t.autos) { d[2 * N + c] = a0; d[3 * N + c] = a1; }
        }
        if (tile == 0 && threadIdx.x == 0) {
            double f = d[4 * (size_t)N];
            for (int r = 0; r < t.world; ++r)
                f += __ldcg(reinterpret_cast<const double *>(reinterpret_cast<const char *>(t.fold_slots) + (size_t)r * t.slot_stride_bytes) + 4 * (size_t)N);
            d[4 * (size_t)N] = f;
        }
        __syncthreads();
        if ((int)threadIdx.x < nr) st_release_sys(t.done + tail_range_cta(threadIdx.x, tile, N), t.fold_epoch);
    }
}

}  // namespace comm
}  // namespace fx
