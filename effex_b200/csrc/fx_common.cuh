// fx_common.cuh -- shared declarations for libeffex_fx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libeffex_fx is written for sm_100a (B200) only"
#endif

namespace fx {

constexpr int kMaxTaps = 32;           // cuSignal channelize_poly cap (n_taps > 32 raises)
constexpr float kByteCentre = 128.0f;  // bytes are centred on 128 exactly; the mean handles the rest

// ---- channel-packed complex: lane x = channel 0, lane y = channel 1 --------
// Every arithmetic op below is one FADD2/FMUL2/FFMA2 on sm_100a; scalar
// twiddles use the .F32 broadcast operand form, so they cost one register.
struct C2 {
    float2 r, i;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, f2neg(b)); }
__device__ __forceinline__ float2 f2muls(float2 a, float s) { return __fmul2_rn(a, f2(s, s)); }
__device__ __forceinline__ float2 f2fmas(float2 a, float s, float2 c) { return __ffma2_rn(a, f2(s, s), c); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ C2 cadd(C2 a, C2 b) { return {f2add(a.r, b.r), f2add(a.i, b.i)}; }
__device__ __forceinline__ C2 csub(C2 a, C2 b) { return {f2sub(a.r, b.r), f2sub(a.i, b.i)}; }
// a + (-i)*b  and  a + (+i)*b
__device__ __forceinline__ C2 cadd_mi(C2 a, C2 b) { return {f2add(a.r, b.i), f2sub(a.i, b.r)}; }
__device__ __forceinline__ C2 cadd_pi(C2 a, C2 b) { return {f2sub(a.r, b.i), f2add(a.i, b.r)}; }
// z * (wr + i*wi), scalar twiddle shared by both channels
__device__ __forceinline__ C2 cmuls(C2 z, float wr, float wi) {
    float2 t = f2muls(z.i, wi);
    float2 u = f2muls(z.i, wr);
    return {f2fmas(z.r, wr, f2neg(t)), f2fmas(z.r, wi, u)};
}

// forward radix-4 butterfly in place: (a0..a3) -> (Y0..Y3), W4 = -i
__device__ __forceinline__ void radix4(C2 &a0, C2 &a1, C2 &a2, C2 &a3) {
    C2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
    C2 s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd_mi(d02, d13);
    a3 = cadd_pi(d02, d13);
}

// register position j of dft16() output holds Y[perm16(j)] (base-4 digit reversal)
__host__ __device__ constexpr int perm16(int j) { return (j >> 2) + 4 * (j & 3); }

// forward 16-point DFT in place on 16 channel-packed complex registers.
__device__ __forceinline__ void dft16(C2 (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;   // sin(pi/8)
    constexpr float R2 = 0.70710678118654752f;   // sqrt(1/2)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) radix4(v[nb], v[nb + 4], v[nb + 8], v[nb + 12]);
    // internal twiddles W16^(nb*ka) on v[nb + 4*ka]
    v[5] = cmuls(v[5], C1, -S1);                                   // W^1
    {   // W^2 = (1-i)/sqrt2 : (x+y, y-x)*R2
        C2 z = v[6];
        v[6] = {f2muls(f2add(z.r, z.i), R2), f2muls(f2sub(z.i, z.r), R2)};
        z = v[9];
        v[9] = {f2muls(f2add(z.r, z.i), R2), f2muls(f2sub(z.i, z.r), R2)};
    }
    v[7] = cmuls(v[7], S1, -C1);                                   // W^3
    v[13] = cmuls(v[13], S1, -C1);                                 // W^3
    {   // W^4 = -i : (y, -x)
        C2 z = v[10];
        v[10] = {z.i, f2neg(z.r)};
    }
    {   // W^6 = (-1-i)/sqrt2 : (y-x, -(x+y))*R2
        C2 z = v[11];
        v[11] = {f2muls(f2sub(z.i, z.r), R2), f2muls(f2add(z.r, z.i), -R2)};
        z = v[14];
        v[14] = {f2muls(f2sub(z.i, z.r), R2), f2muls(f2add(z.r, z.i), -R2)};
    }
    v[15] = cmuls(v[15], -C1, S1);                                 // W^9 = -W^1
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) radix4(v[4 * ka], v[4 * ka + 1], v[4 * ka + 2], v[4 * ka + 3]);
}

// ---- launch bookkeeping ------------------------------------------------------
struct Counters {
    long long launches = 0;
};

}  // namespace fx
