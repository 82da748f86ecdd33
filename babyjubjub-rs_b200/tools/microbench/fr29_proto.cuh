// PROTOTYPE (measurement only): carry-free Montgomery multiplication in radix 2^29 (9 limbs).
// Every limb product is accumulated with a plain IMAD.WIDE.U32 (64-bit accumulate, no carry flag).
#pragma once
#include <stdint.h>

namespace bjj29 {

#if defined(__CUDACC__)
#define B29_HD __host__ __device__ __forceinline__
#else
#define B29_HD inline
#endif

constexpr uint32_t M29 = (1u << 29) - 1;
// Q in radix 2^29
B29_HD uint32_t QLF(int j) {
    constexpr uint32_t t[9] = {0x10000001u, 0x1f0fac9fu, 0x0e5c2450u, 0x07d090f3u, 0x1585d283u,
                            0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu};
    return t[j];
}

struct Fr29 {
    uint32_t v[9];
};

B29_HD void mac(uint64_t& acc, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
    acc += (uint64_t)a * b;
#endif
}

// r = a*b / 2^261 mod Q (lazy: r < 2Q whenever a*b < Q * 2^261); limbs of r < 2^29 (top limb small)
B29_HD void mul(Fr29& r, const Fr29& a, const Fr29& b) {
    uint64_t acc[18];
#pragma unroll
    for (int i = 0; i < 18; i++) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int j = 0; j < 9; j++) mac(acc[i + j], a.v[j], b.v[i]);
        if (i > 0) acc[i] += acc[i - 1] >> 29;
        uint32_t lo = (uint32_t)acc[i];
        uint32_t m = ((lo << 28) - lo) & M29;          // lo * (2^28 - 1) = lo * (-Q^-1) mod 2^29
#pragma unroll
        for (int j = 0; j < 9; j++) mac(acc[i + j], m, QLF(j));
    }
    uint64_t carry = acc[8] >> 29;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        uint64_t t = acc[9 + k] + carry;
        r.v[k] = (k < 8) ? ((uint32_t)t & M29) : (uint32_t)t;
        carry = t >> 29;
    }
}

B29_HD void sqr(Fr29& r, const Fr29& a) {
    uint64_t acc[18];
    uint32_t ad[9];
#pragma unroll
    for (int i = 0; i < 18; i++) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) ad[i] = a.v[i] << 1;
#pragma unroll
    for (int c = 0; c < 17; c++) {
#pragma unroll
        for (int p = 0; p < 9; p++) {
            int q = c - p;
            if (q > p && q < 9) mac(acc[c], a.v[p], ad[q]);
        }
        if ((c & 1) == 0) mac(acc[c], a.v[c / 2], a.v[c / 2]);
        if (c < 9) {
            if (c > 0) acc[c] += acc[c - 1] >> 29;
            uint32_t lo = (uint32_t)acc[c];
            uint32_t m = ((lo << 28) - lo) & M29;
#pragma unroll
            for (int j = 0; j < 9; j++) mac(acc[c + j], m, QLF(j));
        }
    }
    uint64_t carry = acc[8] >> 29;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        uint64_t t = acc[9 + k] + carry;
        r.v[k] = (k < 8) ? ((uint32_t)t & M29) : (uint32_t)t;
        carry = t >> 29;
    }
}

}  // namespace bjj29
