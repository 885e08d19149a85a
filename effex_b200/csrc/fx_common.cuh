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

// ---- helpers with a folded real scale: every one is a pair of FFMA2 ---------------------------
// (x + c*y, y - c*x) = z * (1 - i*c): a twiddle cos*(1 - i*tan) without its cos factor
__device__ __forceinline__ C2 ctw(C2 z, float c) { return {f2fmas(z.i, c, z.r), f2fmas(z.r, -c, z.i)}; }
__device__ __forceinline__ C2 cadd_s(C2 a, float s, C2 b) { return {f2fmas(b.r, s, a.r), f2fmas(b.i, s, a.i)}; }
// a + (-i)*s*b  and  a + (+i)*s*b
__device__ __forceinline__ C2 cadd_mi_s(C2 a, float s, C2 b) { return {f2fmas(b.i, s, a.r), f2fmas(b.r, -s, a.i)}; }
__device__ __forceinline__ C2 cadd_pi_s(C2 a, float s, C2 b) { return {f2fmas(b.i, -s, a.r), f2fmas(b.r, s, a.i)}; }

// second-layer radix-4 whose inputs carry real scales that are folded into the butterfly's FMAs:
//   a1 = al1*a1p, a2 = al2*a2p, a3 = (ratio*al1)*a3p   ->   (a0..a3) := radix4(a0, a1, a2, a3)
__device__ __forceinline__ void radix4_scaled(C2 &a0, C2 &a1p, C2 &a2p, C2 &a3p, float al2, float ratio, float al1) {
    const C2 s02 = cadd_s(a0, al2, a2p), d02 = cadd_s(a0, -al2, a2p);
    const C2 s13 = cadd_s(a1p, ratio, a3p), d13 = cadd_s(a1p, -ratio, a3p);
    a0 = cadd_s(s02, al1, s13);
    a2p = cadd_s(s02, -al1, s13);
    a1p = cadd_mi_s(d02, al1, d13);
    a3p = cadd_pi_s(d02, al1, d13);
}

// forward 16-point DFT in place on 16 channel-packed complex registers (144 packed-FP32 ops).
// Radix 4 x 4; the inner twiddles W16^m are written as alpha*(1 - i*tan) so that only the cheap
// (1 - i*tan) part is applied to the data (2 FFMA2, or 2 FADD2 for m = 2, 6, or nothing for m = 4) and
// the real factor alpha rides for free on the second layer's additions, which become FFMA2.
__device__ __forceinline__ void dft16(C2 (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;   // sin(pi/8)
    constexpr float R2 = 0.70710678118654752f;   // sqrt(1/2)
    constexpr float T1 = 0.41421356237309505f;   // tan(pi/8)
    constexpr float IT1 = 2.41421356237309505f;  // 1/tan(pi/8) = tan(3pi/8)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) radix4(v[nb], v[nb + 4], v[nb + 8], v[nb + 12]);
    // ka = 0: no twiddles
    radix4(v[0], v[1], v[2], v[3]);
    // ka = 1: W^1 = C1(1 - i T1), W^2 = R2(1 - i), W^3 = S1(1 - i/T1)
    {
        C2 a1 = ctw(v[5], T1);
        C2 a2 = {f2add(v[6].r, v[6].i), f2sub(v[6].i, v[6].r)};
        C2 a3 = ctw(v[7], IT1);
        radix4_scaled(v[4], a1, a2, a3, R2, T1, C1);
        v[5] = a1; v[6] = a2; v[7] = a3;
    }
    // ka = 2: W^2 = R2(1 - i), W^4 = -i, W^6 = -R2(1 + i)
    {
        C2 a1 = {f2add(v[9].r, v[9].i), f2sub(v[9].i, v[9].r)};
        C2 a2 = {v[10].i, f2neg(v[10].r)};
        C2 a3 = {f2sub(v[11].r, v[11].i), f2add(v[11].r, v[11].i)};
        const C2 s02 = cadd(v[8], a2), d02 = csub(v[8], a2);
        const C2 s13 = csub(a1, a3), d13 = cadd(a1, a3);       // ratio = -1
        v[8] = cadd_s(s02, R2, s13);
        v[10] = cadd_s(s02, -R2, s13);
        v[9] = cadd_mi_s(d02, R2, d13);
        v[11] = cadd_pi_s(d02, R2, d13);
    }
    // ka = 3: W^3 = S1(1 - i/T1), W^6 = -R2(1 + i), W^9 = -C1(1 - i T1)
    {
        C2 a1 = ctw(v[13], IT1);
        C2 a2 = {f2sub(v[14].r, v[14].i), f2add(v[14].r, v[14].i)};
        C2 a3 = ctw(v[15], T1);
        radix4_scaled(v[12], a1, a2, a3, -R2, -IT1, S1);
        v[13] = a1; v[14] = a2; v[15] = a3;
    }
}

// ---- launch bookkeeping ------------------------------------------------------
struct Counters {
    long long launches = 0;
};

}  // namespace fx
