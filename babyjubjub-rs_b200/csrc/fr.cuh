// Fr: the BN254 scalar field (BabyJubJub base field), 8 x 32-bit limbs, Montgomery form, R = 2^256.
//
// Replaces the reference's `Fr` (src/lib.rs:7 = poseidon_rs::Fr, an ff_ce #[derive(PrimeField)]
// 4x64-bit Montgomery field; Cargo.toml:12,20).  Field arithmetic is exact, so any correct mod-Q
// implementation is bit-identical at the canonical-bytes boundary.
//
// Representation invariant ("lazy"): every Fr held in registers is in [0, 2Q).  Because 4Q < 2^256,
//   * mont_mul of two values < 2Q returns a value < 2Q with NO final conditional subtraction,
//   * a + b < 4Q never overflows 256 bits, one conditional subtraction of 2Q restores the range.
// fr_reduce() produces the canonical value in [0, Q) before any comparison / store.
//
// The multiplier is built from 4-product carry chains: for a fixed b_i, the products a_j*b_i with j
// even put lo at column i+j and hi at column i+j+1 -- 8 distinct consecutive columns, one carry
// chain.  Products with j odd form a second chain shifted by one column.  Two accumulators X / Y
// (absolute columns) receive the chains alternately, so no product ever needs a carry fix-up.
// nvcc -arch=sm_100a fuses each mad.lo.cc / madc.hi.cc pair into ONE IMAD.WIDE.U32[.X]:
// one Montgomery multiplication = 128 IMAD.WIDE + 8 IMAD (m_i) = 136 fma-pipe instructions.
//
// Everything is __host__ __device__: the host build (BJJ_HOST_EMU, used ONLY by the CPU test
// harness tests/hostemu) swaps the PTX chains for a 64-bit C emulation so the limb schedules,
// curve formulas, recodings and table logic can be validated against the oracle without a GPU.
// The shipped library (libbjj_cuda.so) contains device code only and has no CPU fallback.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BJJ_HD __host__ __device__ __forceinline__
#define BJJ_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define BJJ_HD inline
#define BJJ_HD_NOINLINE static
#endif

#if defined(__CUDA_ARCH__)
#define BJJ_DEVICE_CODE 1
#else
#define BJJ_DEVICE_CODE 0
#endif

#if defined(__CUDACC__) && !defined(BJJ_HOST_EMU)
#define BJJ_CONST __constant__ const
#define BJJ_TABLE __device__ const
#else
#define BJJ_CONST static const
#define BJJ_TABLE static const
#endif

#include "generated/bjj_consts.inc"

namespace bjj {

struct Fr {
    uint32_t v[8];
};

// ------------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------------

// c[0..7] += {a0,a1,a2,a3} * b  laid out lo->c[2k], hi->c[2k+1].
// TOP = 0: the chain provably produces no carry-out;  1: carry-out added into c[8];
//       2: carry-out rippled into c[8], c[9] (dot products, where c[8] already holds data).
// SEED: the chain starts with carry-in = carry(sx + sy)  (column fold of the previous step).
#define BJJ_MAC4_BODY(FIRST)                                                                        \
    FIRST ".lo.cc.u32 %0, %10, %14, %0;\n\tmadc.hi.cc.u32 %1, %10, %14, %1;\n\t"                    \
          "madc.lo.cc.u32 %2, %11, %14, %2;\n\tmadc.hi.cc.u32 %3, %11, %14, %3;\n\t"                \
          "madc.lo.cc.u32 %4, %12, %14, %4;\n\tmadc.hi.cc.u32 %5, %12, %14, %5;\n\t"                \
          "madc.lo.cc.u32 %6, %13, %14, %6;\n\tmadc.hi.cc.u32 %7, %13, %14, %7;\n\t"                \
          "addc.cc.u32 %8, %8, 0;\n\taddc.u32 %9, %9, 0;\n\t"
#define BJJ_MAC4_OPERANDS                                                                           \
    : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), \
      "+r"(t8), "+r"(t9)                                                                            \
    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b), "r"(sx), "r"(sy)

template <bool SEED, int TOP>
BJJ_HD void mac4(uint32_t* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b,
                 uint32_t sx = 0, uint32_t sy = 0) {
#if BJJ_DEVICE_CODE
    // Unused tops are fed as dead zero temporaries; ptxas drops the dead addc instructions.
    uint32_t t8 = (TOP >= 1) ? c[8] : 0u, t9 = (TOP >= 2) ? c[9] : 0u;
    if (SEED) {
        asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %15, %16;\n\t" BJJ_MAC4_BODY("madc") "}" BJJ_MAC4_OPERANDS);
    } else {
        asm(BJJ_MAC4_BODY("mad") BJJ_MAC4_OPERANDS);
    }
    if (TOP >= 1) c[8] = t8;
    if (TOP >= 2) c[9] = t9;
#else
    uint64_t carry = SEED ? (((uint64_t)sx + sy) >> 32) : 0;
    const uint32_t a[4] = {a0, a1, a2, a3};
    for (int k = 0; k < 4; k++) {
        uint64_t p = (uint64_t)a[k] * b;
        uint64_t t = (uint64_t)c[2 * k] + (uint32_t)p + carry;
        c[2 * k] = (uint32_t)t;
        carry = t >> 32;
        t = (uint64_t)c[2 * k + 1] + (p >> 32) + carry;
        c[2 * k + 1] = (uint32_t)t;
        carry = t >> 32;
    }
    if (TOP >= 1) {
        uint64_t t = (uint64_t)c[8] + carry;
        c[8] = (uint32_t)t;
        carry = t >> 32;
        if (TOP >= 2) c[9] += (uint32_t)carry;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// add / sub / reduce
// ------------------------------------------------------------------------------------------------

// r = a + b (plain 256-bit, no reduction); returns carry-out
BJJ_HD uint32_t add256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if BJJ_DEVICE_CODE
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\taddc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\taddc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}

// r = a - b (plain 256-bit); returns 0 or 0xffffffff (borrow mask)
BJJ_HD uint32_t sub256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if BJJ_DEVICE_CODE
    uint32_t m;
    asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\tsubc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\tsubc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(m)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return m;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;   // 0 or 0xffffffff
#endif
}

#define BJJ_LIMBS8(P) {P##0, P##1, P##2, P##3, P##4, P##5, P##6, P##7}

// a in [0, 4Q)  ->  [0, 2Q)
BJJ_HD void fr_cond_sub_2q(Fr& a) {
    const uint32_t twoq[8] = BJJ_LIMBS8(BJJ_2Q);
    uint32_t t[8];
    uint32_t borrow = sub256(t, a.v, twoq);
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

// r = a + b  (lazy domain)
BJJ_HD void fr_add(Fr& r, const Fr& a, const Fr& b) {
    add256(r.v, a.v, b.v);
    fr_cond_sub_2q(r);
}

// r = a - b  (lazy domain)
BJJ_HD void fr_sub(Fr& r, const Fr& a, const Fr& b) {
    const uint32_t twoq[8] = BJJ_LIMBS8(BJJ_2Q);
    uint32_t t[8], m[8];
    uint32_t borrow = sub256(t, a.v, b.v);
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = twoq[i] & borrow;
    add256(r.v, t, m);
}

BJJ_HD void fr_dbl(Fr& r, const Fr& a) { fr_add(r, a, a); }

BJJ_HD void fr_zero(Fr& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}

BJJ_HD void fr_neg(Fr& r, const Fr& a) {
    Fr z;
    fr_zero(z);
    fr_sub(r, z, a);
}

// canonical value in [0, Q)
BJJ_HD void fr_reduce(Fr& a) {
    const uint32_t q[8] = BJJ_LIMBS8(BJJ_Q);
    uint32_t t[8];
    uint32_t borrow = sub256(t, a.v, q);
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

// plain 256-bit comparisons on canonical / raw integers
BJJ_HD bool u256_eq(const uint32_t* a, const uint32_t* b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a[i] ^ b[i];
    return d == 0;
}
BJJ_HD bool u256_is_zero(const uint32_t* a) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a[i];
    return d == 0;
}
// a < b
BJJ_HD bool u256_lt(const uint32_t* a, const uint32_t* b) {
    uint32_t t[8];
    return sub256(t, a, b) != 0;
}

// value == 0 mod Q, for a in the lazy domain [0, 2Q)
BJJ_HD bool fr_is_zero(const Fr& a) {
    const uint32_t q[8] = BJJ_LIMBS8(BJJ_Q);
    return u256_is_zero(a.v) || u256_eq(a.v, q);
}

// a == b mod Q (lazy inputs)
BJJ_HD bool fr_eq(const Fr& a, const Fr& b) {
    Fr x = a, y = b;
    fr_reduce(x);
    fr_reduce(y);
    return u256_eq(x.v, y.v);
}

// ------------------------------------------------------------------------------------------------
// Montgomery multiplication
// ------------------------------------------------------------------------------------------------

// m = t * (-Q^-1 mod 2^32).  Q = 1 - 2^28 (mod 2^32), so -Q^-1 = -(1 + 2^28) and the product is a shift, an add and a
// negation on the ALU pipe: the multiplier pipe, which bounds every kernel here, keeps its slots for the wide MACs
// (8 IMAD less per multiplication, 3 % of its multiplier-pipe time).
BJJ_HD uint32_t mont_m(uint32_t t) {
    static_assert(BJJ_NINV32 == 0xefffffffu, "mont_m is specialised for Q = 0xf0000001 (mod 2^32)");
#if BJJ_DEVICE_CODE
    uint32_t s, m;
    asm("shf.l.clamp.b32 %0, 0, %1, 28;" : "=r"(s) : "r"(t));      // t << 28 as a funnel shift (SHF: ALU pipe)
    asm("sub.u32 %0, 0, %1;\n\tsub.u32 %0, %0, %2;" : "=r"(m) : "r"(s), "r"(t));
    return m;
#else
    return 0u - (t + (t << 28));
#endif
}

// One interleaved CIOS step at absolute column I.  S is the accumulator whose chain starts at column
// I (X when I is even, Y when I is odd), N the other one.  FOLD: carry of column I-1 enters the chain.
//
// Column bound: after step I the running total is < 2^(32(I+1)) * (a + Q) < 2^(32(I+9)) because
// a + Q < 3Q < 2^256, and X, Y are each <= the total.  Hence the N chain (columns I+1..I+8) never
// carries out (TOP = 0) and the S chain (columns I..I+7) carries into a fresh S[I+8] (TOP = 1).
#define BJJ_MUL_STEP(S, N, I, FOLD)                                                                   \
    {                                                                                                   \
        mac4<FOLD, 1>(&S[I], a.v[0], a.v[2], a.v[4], a.v[6], b.v[I], X[(I) ? (I)-1 : 0], Y[(I) ? (I)-1 : 0]); \
        mac4<false, 0>(&N[I + 1], a.v[1], a.v[3], a.v[5], a.v[7], b.v[I]);                              \
        uint32_t m = mont_m(S[I] + N[I]);                                                        \
        mac4<false, 1>(&S[I], BJJ_Q0, BJJ_Q2, BJJ_Q4, BJJ_Q6, m);                                       \
        mac4<false, 0>(&N[I + 1], BJJ_Q1, BJJ_Q3, BJJ_Q5, BJJ_Q7, m);                                   \
    }

// result limbs = columns 8..15 of X + Y, plus the carry of column 7 (X[7] + Y[7] is 0 or 2^32)
BJJ_HD void fr_merge_xy(Fr& r, const uint32_t* X, const uint32_t* Y) {
#if BJJ_DEVICE_CODE
    asm("{\n\t.reg .u32 t;\n\t"
        "add.cc.u32 t, %8, %9;\n\t"
        "addc.cc.u32 %0, %10, %18;\n\taddc.cc.u32 %1, %11, %19;\n\taddc.cc.u32 %2, %12, %20;\n\t"
        "addc.cc.u32 %3, %13, %21;\n\taddc.cc.u32 %4, %14, %22;\n\taddc.cc.u32 %5, %15, %23;\n\t"
        "addc.cc.u32 %6, %16, %24;\n\taddc.u32 %7, %17, %25;\n\t}"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(X[7]), "r"(Y[7]),
          "r"(X[8]), "r"(X[9]), "r"(X[10]), "r"(X[11]), "r"(X[12]), "r"(X[13]), "r"(X[14]), "r"(X[15]),
          "r"(Y[8]), "r"(Y[9]), "r"(Y[10]), "r"(Y[11]), "r"(Y[12]), "r"(Y[13]), "r"(Y[14]), "r"(Y[15]));
#else
    uint64_t c = ((uint64_t)X[7] + Y[7]) >> 32;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)X[8 + i] + Y[8 + i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// r = a * b / 2^256 mod Q;  a, b in [0, 2Q)  ->  r in [0, 2Q)   (also valid for a < 2^256, b < Q)
BJJ_HD void fr_mul_inline(Fr& r, const Fr& a, const Fr& b) {
    uint32_t X[18], Y[18];
#pragma unroll
    for (int i = 0; i < 18; i++) X[i] = Y[i] = 0;
    BJJ_MUL_STEP(X, Y, 0, false)
    BJJ_MUL_STEP(Y, X, 1, true)
    BJJ_MUL_STEP(X, Y, 2, true)
    BJJ_MUL_STEP(Y, X, 3, true)
    BJJ_MUL_STEP(X, Y, 4, true)
    BJJ_MUL_STEP(Y, X, 5, true)
    BJJ_MUL_STEP(X, Y, 6, true)
    BJJ_MUL_STEP(Y, X, 7, true)
    fr_merge_xy(r, X, Y);
}

// (An out-of-line multiplier called with operands BY VALUE was measured twice and rejected: ~45 register moves per
// call, 22.5 vs 26.1 M mults/s in k_mul_scalar, no gain in k_verify_ec.  Out-of-line multiplication with operands
// in SHARED memory is a different design: see vm.cuh.)
BJJ_HD void fr_mul(Fr& r, const Fr& a, const Fr& b) { fr_mul_inline(r, a, b); }

// One step of a Montgomery dot product  sum_p A_p * B_p  at column I (see fr_dot).
#define BJJ_DOT_STEP(S, N, I, FOLD)                                                                   \
    {                                                                                                   \
        _Pragma("unroll") for (int p = 0; p < NP; p++) {                                                \
            if (FOLD && p == 0)                                                                         \
                mac4<true, 2>(&S[I], A[p].v[0], A[p].v[2], A[p].v[4], A[p].v[6], B[p].v[I], X[(I) ? (I)-1 : 0], Y[(I) ? (I)-1 : 0]); \
            else                                                                                        \
                mac4<false, 2>(&S[I], A[p].v[0], A[p].v[2], A[p].v[4], A[p].v[6], B[p].v[I]);           \
            mac4<false, 1>(&N[I + 1], A[p].v[1], A[p].v[3], A[p].v[5], A[p].v[7], B[p].v[I]);           \
        }                                                                                               \
        uint32_t m = mont_m(S[I] + N[I]);                                                        \
        mac4<false, 2>(&S[I], BJJ_Q0, BJJ_Q2, BJJ_Q4, BJJ_Q6, m);                                       \
        mac4<false, 1>(&N[I + 1], BJJ_Q1, BJJ_Q3, BJJ_Q5, BJJ_Q7, m);                                   \
    }

// r = (sum_{p<NP} A_p * B_p) / 2^256 mod Q  with ONE interleaved reduction (NP*64 + 72 IMADs instead
// of NP*136).  Requirements: every A_p canonical (< Q), every B_p in the lazy domain (< 2Q),
// NP <= 9.  Column bound: total after step I < 2^(32(I+1)) * (NP*Q + Q) < 2^(32(I+9)+2), so the
// tops ripple one limb further than in fr_mul.  The result is < (2*NP*Q/2^256 + 1) * Q <= 4.4Q
// < 2^256 and is brought back to [0, 2Q) by one (NP <= 7) or two conditional subtractions.
template <int NP>
BJJ_HD void fr_dot(Fr& r, const Fr* A, const Fr* B) {
    uint32_t X[19], Y[19];
#pragma unroll
    for (int i = 0; i < 19; i++) X[i] = Y[i] = 0;
    BJJ_DOT_STEP(X, Y, 0, false)
    BJJ_DOT_STEP(Y, X, 1, true)
    BJJ_DOT_STEP(X, Y, 2, true)
    BJJ_DOT_STEP(Y, X, 3, true)
    BJJ_DOT_STEP(X, Y, 4, true)
    BJJ_DOT_STEP(Y, X, 5, true)
    BJJ_DOT_STEP(X, Y, 6, true)
    BJJ_DOT_STEP(Y, X, 7, true)
    fr_merge_xy(r, X, Y);
    fr_cond_sub_2q(r);
    if (NP > 7) fr_cond_sub_2q(r);
}

// ------------------------------------------------------------------------------------------------
// Montgomery squaring: 36 + 64 wide MACs instead of 128
// ------------------------------------------------------------------------------------------------

// c[0..2N-1] += {a0..a(N-1)} * b (lo -> c[2k], hi -> c[2k+1]); TOP = 1: carry-out added into c[2N], 0: provably none.
template <int N, int TOP>
BJJ_HD void macn(uint32_t* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    static_assert(N >= 1 && N <= 4 && TOP >= 0 && TOP <= 1, "macn");
#if BJJ_DEVICE_CODE
    uint32_t t = TOP ? c[2 * N] : 0u;       // TOP = 0: a dead zero temporary, ptxas drops the addc
    if (N == 1) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(t) : "r"(a0), "r"(b));
    } else if (N == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\taddc.u32 %4, %4, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(t) : "r"(a0), "r"(a1), "r"(b));
    } else if (N == 3) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\taddc.u32 %6, %6, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(t)
            : "r"(a0), "r"(a1), "r"(a2), "r"(b));
    } else {
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]), "+r"(t)
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
    }
    if (TOP) c[2 * N] = t;
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
    if (TOP) c[2 * N] += (uint32_t)carry;
#endif
}

// s + n + carry(sx + sy): the value of column I once the carry of column I-1 (sx + sy is 0 or 2^32) is folded in
BJJ_HD uint32_t fold_sum(uint32_t s, uint32_t n, uint32_t sx, uint32_t sy) {
#if BJJ_DEVICE_CODE
    uint32_t r;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %3, %4;\n\taddc.u32 %0, %1, %2;\n\t}" : "=r"(r) : "r"(s), "r"(n), "r"(sx), "r"(sy));
    return r;
#else
    return s + n + (uint32_t)(((uint64_t)sx + sy) >> 32);
#endif
}

// Row I of the squaring triangle,  a_I * (a_I + 2 * (a >> 32(I+1)) * 2^32)  at absolute column 2I, followed by the
// reduction step of column I.  d[j] = limb j of 2a (j >= 2), c = a_(I+1) << 1 (the limb of 2 * (a >> 32(I+1)) that
// takes no bit from below).  Products with j = I (mod 2) start at the even column 2I: always the X chain; the others
// start at 2I + 1: always Y.  Column bound: after step I the total is < 2^(32(I+1)) * (2a + Q) < 2^(32(I+9)) because
// 2a + Q < 5Q < 2^256, and X, Y are each <= the total.  The row's X chain ends at column I+7 (I even: its carry opens
// the fresh X[I+8]) or I+8 (I odd: no carry-out); the Y chain the other way round.  The reduction chains are those of
// BJJ_MUL_STEP, except that the carry of column I-1 enters with the reduction (nothing else starts at column I).
#define BJJ_SQR_ROW(I, NX, NY, x1, x2, x3, y0, y1, y2, y3)                                              \
    {                                                                                                   \
        macn<NX, ((I) & 1) ? 0 : 1>(&X[2 * (I)], a.v[I], x1, x2, x3, a.v[I]);                           \
        if (NY > 0) macn<(NY > 0 ? NY : 1), ((I) & 1) ? 1 : 0>(&Y[2 * (I) + 1], y0, y1, y2, y3, a.v[I]); \
    }
#define BJJ_SQR_REDC(S, N, I)                                                                           \
    {                                                                                                   \
        const uint32_t m = mont_m((I) ? fold_sum(S[I], N[I], X[(I) ? (I)-1 : 0], Y[(I) ? (I)-1 : 0]) : S[I] + N[I]); \
        mac4<((I) > 0), 1>(&S[I], BJJ_Q0, BJJ_Q2, BJJ_Q4, BJJ_Q6, m, X[(I) ? (I)-1 : 0], Y[(I) ? (I)-1 : 0]); \
        mac4<false, 0>(&N[I + 1], BJJ_Q1, BJJ_Q3, BJJ_Q5, BJJ_Q7, m);                                   \
    }

// r = a * a / 2^256 mod Q;  a in [0, 2Q)  ->  r in [0, 2Q)  (r < 4Q^2 / 2^256 + Q < 1.76 Q).  The same integer as
// fr_mul_inline(r, a, a) up to a multiple of Q -- in fact identical: both return (a^2 + m Q) / 2^256 with the same m.
BJJ_HD void fr_sqr_inline(Fr& r, const Fr& a) {
    uint32_t X[18], Y[18], d[8], c[8];
#pragma unroll
    for (int i = 0; i < 18; i++) X[i] = Y[i] = 0;
#pragma unroll
    for (int j = 1; j < 8; j++) {
        c[j] = a.v[j] << 1;
        d[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
    }
    BJJ_SQR_ROW(0, 4, 4, d[2], d[4], d[6], c[1], d[3], d[5], d[7])
    BJJ_SQR_REDC(X, Y, 0)
    BJJ_SQR_ROW(1, 4, 3, d[3], d[5], d[7], c[2], d[4], d[6], 0u)
    BJJ_SQR_REDC(Y, X, 1)
    BJJ_SQR_ROW(2, 3, 3, d[4], d[6], 0u, c[3], d[5], d[7], 0u)
    BJJ_SQR_REDC(X, Y, 2)
    BJJ_SQR_ROW(3, 3, 2, d[5], d[7], 0u, c[4], d[6], 0u, 0u)
    BJJ_SQR_REDC(Y, X, 3)
    BJJ_SQR_ROW(4, 2, 2, d[6], 0u, 0u, c[5], d[7], 0u, 0u)
    BJJ_SQR_REDC(X, Y, 4)
    BJJ_SQR_ROW(5, 2, 1, d[7], 0u, 0u, c[6], 0u, 0u, 0u)
    BJJ_SQR_REDC(Y, X, 5)
    BJJ_SQR_ROW(6, 1, 1, 0u, 0u, 0u, c[7], 0u, 0u, 0u)
    BJJ_SQR_REDC(X, Y, 6)
    BJJ_SQR_ROW(7, 1, 0, 0u, 0u, 0u, 0u, 0u, 0u, 0u)
    BJJ_SQR_REDC(Y, X, 7)
    fr_merge_xy(r, X, Y);
}

#ifndef BJJ_DEDICATED_SQR
#define BJJ_DEDICATED_SQR 1
#endif
BJJ_HD void fr_sqr(Fr& r, const Fr& a) {
#if BJJ_DEDICATED_SQR
    fr_sqr_inline(r, a);
#else
    fr_mul(r, a, a);
#endif
}

// ------------------------------------------------------------------------------------------------
// conversions and helpers
// ------------------------------------------------------------------------------------------------

BJJ_HD void fr_set(Fr& r, const uint32_t* w) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
}

BJJ_HD Fr fr_const(const uint32_t* w) {
    Fr r;
    fr_set(r, w);
    return r;
}

// any 256-bit integer -> Montgomery form of (x mod Q), lazy domain.  R2 must be the FIRST operand:
// the column bound of fr_mul needs a + Q < 2^256 for the multiplicand `a`, while `b` may be any
// 256-bit value as long as a*b < Q*2^256.
BJJ_HD void fr_to_mont(Fr& r, const Fr& x) {
    Fr r2 = fr_const(BJJ_R2);
    fr_mul(r, r2, x);
}

// Montgomery form -> canonical integer in [0, Q)
BJJ_HD void fr_from_mont(Fr& r, const Fr& a) {
    Fr one;
    fr_zero(one);
    one.v[0] = 1;
    fr_mul(r, a, one);
    fr_reduce(r);
}

BJJ_HD void fr_cmov(Fr& r, const Fr& a, bool take) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = take ? a.v[i] : r.v[i];
}

// r = a^e for a public 256-bit exponent (plain binary, MSB first).  Not unrolled: one sqr + one mul
// body in the instruction stream.
BJJ_HD void fr_pow(Fr& r, const Fr& a, const uint32_t* e, int nbits) {
    Fr acc = fr_const(BJJ_ONE_M);
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        fr_sqr(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) fr_mul(acc, acc, a);
    }
    r = acc;
}

// r = a^e for one of the library's two fixed exponents, by the sliding-window schedule generated at build time
// (tools/gen_constants.py::pow_schedule: window of 4 bits over the odd powers a, a^3, .., a^15).  The schedule is the
// same for every lane, so the walk does not diverge; the table of odd powers is indexed dynamically and lives in local
// memory (256 B per thread, read once per step).  One squaring body and one multiplication body in the instruction stream.
BJJ_HD void fr_pow_sched(Fr& r, const Fr& a, const uint8_t (*steps)[2], int nsteps, int tail) {
    Fr tab[BJJ_POW_TABLE], a2;
    tab[0] = a;
    fr_sqr(a2, a);
#pragma unroll 1
    for (int i = 1; i < BJJ_POW_TABLE; i++) fr_mul(tab[i], tab[i - 1], a2);
    Fr acc = tab[steps[0][1]];
#pragma unroll 1
    for (int k = 1; k < nsteps; k++) {
#pragma unroll 1
        for (int s = steps[k][0]; s > 0; s--) fr_sqr(acc, acc);
        fr_mul(acc, acc, tab[steps[k][1]]);
    }
#pragma unroll 1
    for (int s = tail; s > 0; s--) fr_sqr(acc, acc);
    r = acc;
}

// Fermat inverse a^(Q-2); 0 -> 0.  (BJJ_POW_SCHED=0: plain binary exponentiation, 254 squarings + 126 multiplications.)
#ifndef BJJ_POW_SCHED
#define BJJ_POW_SCHED 1
#endif
BJJ_HD void fr_inv(Fr& r, const Fr& a) {
#if BJJ_POW_SCHED
    fr_pow_sched(r, a, BJJ_POW_QM2, BJJ_POW_QM2_STEPS, BJJ_POW_QM2_TAIL);
#else
    fr_pow(r, a, BJJ_EXP_QM2, 254);
#endif
}

// ---- pipe-selection ballast ----------------------------------------------------------------------------
// ptxas decides per KERNEL, from static instruction counts, whether integer adds, moves and negations go to the ALU
// pipe (IADD3, MOV) or ride the fma pipe as IMAD.IADD / IMAD.MOV / IMAD.X -- and it counts a wide multiply as one
// slot, although IMAD.WIDE holds the fma pipe for two.  A kernel with ALU-heavy stretches (scalar recoding, digit
// extraction, BLAKE-512, address arithmetic) therefore gets "passengers" inside its multiplications: 16 per product
// in k_verify_ec_vm (7 % of the pipe that bounds the kernel), 14 % of the multiplier pipe in k_public
// (profiles/r2_ncu_kernels_summary.txt).  The ballast is a block of N plain multiply-adds behind a condition that is
// never true at run time (never fetched, never executed): it tips the static balance so that the compiler keeps
// those instructions on the ALU pipe.  `cond` must be opaque to the compiler; the result is stored so that the block
// is not dead code.  (tools/microbench has the experiment that shows the mechanism.)
#if defined(__CUDACC__) && !defined(BJJ_HOST_EMU)
#ifndef BJJ_BALLAST
#define BJJ_BALLAST 1500
#endif
__device__ __forceinline__ void fma_ballast(bool cond, uint32_t seed, uint8_t* sink) {
#if BJJ_BALLAST > 0
    if (cond) {
        uint32_t g = seed, h = seed | 3u;
#pragma unroll
        for (int j = 0; j < BJJ_BALLAST; j++) g = g * h + (uint32_t)j;
        sink[0] = (uint8_t)g;
    }
#endif
}
#endif

}  // namespace bjj
