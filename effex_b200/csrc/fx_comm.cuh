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
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(t.flag + blockIdx.x, t.epoch);
    }
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
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(done + blockIdx.x, epoch);
    }
}

}  // namespace comm
}  // namespace fx
