// Dedicated Montgomery squaring prototype (NOT part of the library; only tools/microbench/imad_bench.cu includes it).
// Kept for the record of the measurement: 70-75 G/s in isolation against 68.3 G/s for fr_mul(a, a), but a net loss
// inside k_verify_hash / k_verify_ec (13.1 vs 13.9 M verifies/s; retested in the final round-1 kernels: k_verify_ec
// 72.4 ms against 67.0 ms) -- the 28 multiplies saved cost more carry-flag traffic and registers than they remove.
#pragma once
#include "../../csrc/fr.cuh"

namespace bjj {

// ------------------------------------------------------------------------------------------------
// dedicated squaring: 28 cross products (doubled by a 1-bit shift) + 8 squares + 64 reduction MACs
// = 100 wide MACs instead of 128
// ------------------------------------------------------------------------------------------------

// c[0..2N-1] += {a0..a(N-1)} * b (lo -> c[2k], hi -> c[2k+1]); carry-out added into c[2N].
template <int N>
BJJ_HD void macn(uint32_t* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if BJJ_DEVICE_CODE
    if (N == 1) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]) : "r"(a0), "r"(b));
    } else if (N == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\taddc.u32 %4, %4, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]) : "r"(a0), "r"(a1), "r"(b));
    } else if (N == 3) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\taddc.u32 %6, %6, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6])
            : "r"(a0), "r"(a1), "r"(a2), "r"(b));
    } else {
        mac4<false, 1>(c, a0, a1, a2, a3, b);
    }
#else
    const uint32_t a[4] = {a0, a1, a2, a3};
    uint64_t carry = 0;
    for (int k = 0; k < N; k++) {
        uint64_t p = (uint64_t)a[k] * b;
        uint64_t t = (uint64_t)c[2 * k] + (uint32_t)p + carry;
        c[2 * k] = (uint32_t)t;
        carry = t >> 32;
        t = (uint64_t)c[2 * k + 1] + (p >> 32) + carry;
        c[2 * k + 1] = (uint32_t)t;
        carry = t >> 32;
    }
    c[2 * N] += (uint32_t)carry;
#endif
}

// Montgomery reduction of a 512-bit value T (16 limbs, T < 2^256 * 1.2 Q):  r = T / 2^256 mod Q, r < 2Q.
// redc(T) = mont_mul(T_lo, 1) + T_hi: the accumulators start as (X, Y) = (T_lo, 0) and run the eight
// reduction-only steps of fr_mul (same column bounds: every chain top lands in an empty column); the
// upper half of T is added once at the end.
BJJ_HD void fr_redc16(Fr& r, const uint32_t* T) {
    uint32_t X[18], Y[18];
#pragma unroll
    for (int i = 0; i < 18; i++) {
        X[i] = i < 8 ? T[i] : 0;
        Y[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* S = (i & 1) ? Y : X;                 // chain that starts at column i
        uint32_t* N = (i & 1) ? X : Y;
        // carry of column i-1: X[i-1] + Y[i-1] is 0 or 2^32
        const uint32_t c = i ? (((X[i ? i - 1 : 0] | Y[i ? i - 1 : 0]) != 0) ? 1u : 0u) : 0u;
        const uint32_t m = (S[i] + N[i] + c) * BJJ_NINV32;
        if (i == 0)
            mac4<false, 1>(&S[i], BJJ_Q0, BJJ_Q2, BJJ_Q4, BJJ_Q6, m);
        else
            mac4<true, 1>(&S[i], BJJ_Q0, BJJ_Q2, BJJ_Q4, BJJ_Q6, m, X[i ? i - 1 : 0], Y[i ? i - 1 : 0]);
        mac4<false, 0>(&N[i + 1], BJJ_Q1, BJJ_Q3, BJJ_Q5, BJJ_Q7, m);
    }
    Fr u;
    fr_merge_xy(u, X, Y);
    add256(r.v, u.v, T + 8);
}

BJJ_HD void fr_sqr_dedicated(Fr& r, const Fr& a) {
    uint32_t X[18], Y[18];
#pragma unroll
    for (int i = 0; i < 18; i++) X[i] = Y[i] = 0;
    // cross products a_i * a_j (i < j): (i + j) odd -> Y, even -> X; see the column audit in DESIGN.md
    macn<4>(&Y[1], a.v[1], a.v[3], a.v[5], a.v[7], a.v[0]);
    macn<3>(&X[2], a.v[2], a.v[4], a.v[6], 0, a.v[0]);
    macn<3>(&Y[3], a.v[2], a.v[4], a.v[6], 0, a.v[1]);
    macn<3>(&X[4], a.v[3], a.v[5], a.v[7], 0, a.v[1]);
    macn<3>(&Y[5], a.v[3], a.v[5], a.v[7], 0, a.v[2]);
    macn<2>(&X[6], a.v[4], a.v[6], 0, 0, a.v[2]);
    macn<2>(&Y[7], a.v[4], a.v[6], 0, 0, a.v[3]);
    macn<2>(&X[8], a.v[5], a.v[7], 0, 0, a.v[3]);
    macn<2>(&Y[9], a.v[5], a.v[7], 0, 0, a.v[4]);
    macn<1>(&X[10], a.v[6], 0, 0, 0, a.v[4]);
    macn<1>(&Y[11], a.v[6], 0, 0, 0, a.v[5]);
    macn<1>(&X[12], a.v[7], 0, 0, 0, a.v[5]);
    macn<1>(&Y[13], a.v[7], 0, 0, 0, a.v[6]);
    // Z = X + Y (columns 0..15), then doubled, then the squares a_i^2 at columns 2i, 2i+1
    uint32_t T[16];
    {
        uint32_t cin[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; i++) cin[i] = 0;
        cin[0] = add256(T, X, Y);          // lower half; its carry enters the upper half
        add256(hi, X + 8, Y + 8);
        add256(T + 8, hi, cin);
    }
#pragma unroll
    for (int i = 15; i > 0; i--) T[i] = (T[i] << 1) | (T[i - 1] >> 31);
    T[0] <<= 1;
#if BJJ_DEVICE_CODE
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\tmadc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\tmadc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\tmadc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\tmadc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\tmadc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\tmadc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\tmadc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\tmadc.hi.u32 %15, %23, %23, %15;"
        : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]),
          "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
#else
    {
        uint64_t carry = 0;
        for (int i = 0; i < 8; i++) {
            uint64_t p = (uint64_t)a.v[i] * a.v[i];
            uint64_t t = (uint64_t)T[2 * i] + (uint32_t)p + carry;
            T[2 * i] = (uint32_t)t;
            carry = t >> 32;
            t = (uint64_t)T[2 * i + 1] + (p >> 32) + carry;
            T[2 * i + 1] = (uint32_t)t;
            carry = t >> 32;
        }
    }
#endif
    fr_redc16(r, T);
}


}  // namespace bjj
