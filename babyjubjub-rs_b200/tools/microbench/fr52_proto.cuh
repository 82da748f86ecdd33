// PROTOTYPE (measurement only): Montgomery multiplication on the FP64 pipe.
// 5 limbs of 52 bits held as doubles; every 52x52 limb product is split exactly into its high and low
// 52 bits by two round-toward-zero DFMAs (Emmart's trick) and accumulated as 64-bit integers:
//     hi = fma_rz(a, b, 2^104)            = 2^104 + floor(ab / 2^52) * 2^52
//     lo = fma_rz(a, b, 2^104 + 2^52 - hi) = 2^52 + (ab mod 2^52)
// The mantissa fields of hi / lo ARE the integers floor(ab/2^52) and ab mod 2^52.
// R = 2^260.  Purpose: B200's DFMA issues at 64 lanes/clk/SM on its own pipe, while IMAD.WIDE
// (32 lanes/clk/SM) saturates the fma pipe -- the two multipliers could run concurrently in
// different warps.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define F52_HD __host__ __device__ __forceinline__
#else
#include <cfenv>
#include <cmath>
#define F52_HD inline
#endif

namespace bjj52 {

struct Fr52 {
    double v[5];    // integers in [0, 2^52)
};

constexpr uint64_t MASK52 = (1ull << 52) - 1;
// Q in radix 2^52
F52_HD double q_limb(int j) {
    constexpr double t[5] = {(double)0x1f593f0000001ull, (double)0x4879b9709143eull, (double)0x181585d2833e8ull,
                             (double)0xa029b85045b68ull, (double)0x030644e72e131ull};
    return t[j];
}
// -Q^-1 mod 2^52
constexpr uint64_t NINV52 = 0x1f593efffffffull;

F52_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return std::fma(a, b, c);      // caller sets FE_TOWARDZERO
#endif
}
F52_HD uint64_t dbits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    __builtin_memcpy(&u, &d, 8);
    return u;
#endif
}
F52_HD double bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    __builtin_memcpy(&d, &u, 8);
    return d;
#endif
}
// integer in [0, 2^52) -> double, exactly, without an I2F conversion
F52_HD double int52_to_double(uint64_t x) { return bits_to_double(x | 0x4330000000000000ull) - 4503599627370496.0; }

// r = a * b / 2^260 mod Q (lazy, r < 2Q for a, b < 2^256)
F52_HD void mul(Fr52& r, const Fr52& a, const Fr52& b) {
    const double C1 = 20282409603651670423947251286016.0;               // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;   // 2^104 + 2^52
    const uint64_t BIAS_HI = 0x4670000000000000ull;                        // bits(2^104)
    const uint64_t BIAS_LO = 0x4330000000000000ull;                        // bits(2^52)
    uint64_t col[11];
#pragma unroll
    for (int k = 0; k < 11; k++) col[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double hi = fma_rz(a.v[i], b.v[j], C1);
            double lo = fma_rz(a.v[i], b.v[j], C2 - hi);
            col[i + j + 1] += dbits(hi) - BIAS_HI;
            col[i + j] += dbits(lo) - BIAS_LO;
        }
        // reduction step for column i
        if (i > 0) col[i] += col[i - 1] >> 52;
        uint64_t q = ((col[i] & MASK52) * NINV52) & MASK52;
        double qd = int52_to_double(q);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double hi = fma_rz(qd, q_limb(j), C1);
            double lo = fma_rz(qd, q_limb(j), C2 - hi);
            col[i + j + 1] += dbits(hi) - BIAS_HI;
            col[i + j] += dbits(lo) - BIAS_LO;
        }
    }
    uint64_t carry = col[4] >> 52;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint64_t t = col[5 + k] + carry;
        r.v[k] = int52_to_double(k < 4 ? (t & MASK52) : t);
        carry = t >> 52;
    }
}

}  // namespace bjj52
