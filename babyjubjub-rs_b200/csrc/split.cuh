// Half-size scalars for EdDSA verification.
//
// The reference's verify (src/lib.rs:395-412) accepts iff  S*B8 == R8 + (8*hm)*A.  With both input points on
// the curve this is the group equation  P := S*B8 - R8 - hm*(8A) == O.  The group has order 8*l
// (l = SUBORDER, prime), so for any ODD v with 0 < |v| < l the map P -> v*P is a bijection and
//     P == O   <=>   (v*S mod l)*B8 - v*R8 - u*(8A) == O,      u = v*hm mod l
// (B8 and 8A have order dividing l, so their scalars may be reduced mod l; R8 may carry a torsion component,
// which is why v must also be odd).  The extended Euclidean algorithm on (l, hm) walks through relations
// u_i = v_i*hm (mod l) with u_i * |v_i| < l; stopping half-way gives |u|, |v| ~ sqrt(l) = 2^125, so the
// double-scalar pass over the two per-lane points needs 32-33 radix-16 windows instead of 64, while the
// full-width scalar (v*S mod l) only meets the precomputed B8 tables.  (Antipa, Brown, Gallant, Lambert,
// Struik, Vanstone: "Accelerated verification of ECDSA signatures", SAC 2005.)
//
// Nothing here is approximate: any (u, v) this file returns satisfies the relation exactly, v is odd and
// non-zero, and verify_fast runs as many windows as the wider of the two needs -- a lane whose lattice has no
// short odd vector is only slower, never different.
#pragma once
#include "fr.cuh"

namespace bjj {

#if BJJ_DEVICE_CODE
#define BJJ_CLZ32(x) __clz((int)(x))
#else
#define BJJ_CLZ32(x) ((x) ? __builtin_clz(x) : 32)
#endif

BJJ_HD int u256_bitlen(const uint32_t* a) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (a[i]) n = 32 * i + 32 - BJJ_CLZ32(a[i]);
    return n;
}

// r = a << k, 0 <= k < 32 (bits shifted out of limb 7 are dropped)
BJJ_HD void u256_shl(uint32_t* r, const uint32_t* a, int k) {
#if BJJ_DEVICE_CODE
#pragma unroll
    for (int i = 7; i > 0; i--) r[i] = __funnelshift_l(a[i - 1], a[i], k);
    r[0] = a[0] << k;
#else
#pragma unroll
    for (int i = 7; i > 0; i--) r[i] = k ? ((a[i] << k) | (a[i - 1] >> (32 - k))) : a[i];
    r[0] = a[0] << k;
#endif
}

BJJ_HD void u256_shr1(uint32_t* r, const uint32_t* a) {
#pragma unroll
    for (int i = 0; i < 7; i++) r[i] = (a[i] >> 1) | (a[i + 1] << 31);
    r[7] = a[7] >> 1;
}

// x < 0x70000000 * 2^96  (~2^126.8): the largest scalars whose signed radix-16 recoding fits 32 windows
BJJ_HD bool split_small(const uint32_t* x) {
    return (x[4] | x[5] | x[6] | x[7]) == 0 && x[3] < 0x70000000u;
}

// One shift-and-subtract step on the unshifted pair (a, ta) against (b, tb), a >= b:  a -= 2^k b, |ta| += 2^k |tb|
// with the largest k that keeps a >= 0.  Slow but simple; only the (short) second phase below uses it.
BJJ_HD void split_step_pow2(uint32_t* a, uint32_t* ta, const uint32_t* b, const uint32_t* tb) {
    int k = u256_bitlen(a) - u256_bitlen(b);
    uint32_t c[8], tc[8], d[8], half[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c[i] = b[i];
        tc[i] = tb[i];
    }
#pragma unroll 1
    for (; k >= 32; k -= 32) {          // whole-limb shifts: degenerate h only
#pragma unroll
        for (int i = 7; i > 0; i--) {
            c[i] = c[i - 1];
            tc[i] = tc[i - 1];
        }
        c[0] = 0;
        tc[0] = 0;
    }
    u256_shl(c, c, k);
    u256_shl(tc, tc, k);
    const uint32_t over = sub256(d, a, c);       // borrow: c > a (then k >= 1) and half of c is used
    u256_shr1(half, c);
#pragma unroll
    for (int i = 0; i < 8; i++) half[i] = over ? half[i] : 0u;
    add256(a, d, half);
    u256_shr1(half, tc);
#pragma unroll
    for (int i = 0; i < 8; i++) tc[i] = over ? half[i] : tc[i];
    add256(ta, ta, tc);
}

// x -= q * y  (x >= q * y)
BJJ_HD void u256_mulsub(uint32_t* x, const uint32_t* y, uint32_t q) {
    uint64_t carry = 0;
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        carry += (uint64_t)q * y[i];
        const uint32_t lo = (uint32_t)carry;
        carry >>= 32;
        const uint64_t d = (uint64_t)x[i] - lo - borrow;
        x[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 63);
    }
}
// x += q * y  (no overflow)
BJJ_HD void u256_muladd(uint32_t* x, const uint32_t* y, uint32_t q) {
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        carry += (uint64_t)q * y[i] + x[i];
        x[i] = (uint32_t)carry;
        carry >>= 32;
    }
}

// u, |v| and the sign of v (vneg = 1: v < 0) with  u = v * h (mod l),  v odd,  0 <= u,  0 < |v| < l.
// h is any 256-bit integer.
BJJ_HD void split_scalars(uint32_t* u, uint32_t* v, uint32_t& vneg, const uint32_t* h) {
    // Invariants:  a = ta*h, b = tb*h (mod l);  a >= b >= 0;  ta and tb have opposite signs (tb carries
    // (-1)^bneg, a zero has no sign);  a*|tb| + b*|ta| = l.  A step replaces (a, ta) by (a - q b, ta - q tb)
    // -- in magnitudes |ta| + q |tb| -- for ANY q with q b <= a, so every intermediate pair is a valid relation.
    // Phase 1 keeps A = a 2^s and B = b 2^s with the top bit of A set, so a 32-bit quotient estimate needs no
    // limb indexing:  q = max(1, A[7] / (B[7] + 1)) never exceeds a / b.
    uint32_t A[8], B[8], ta[8], tb[8];
    const bool h_ge_l = !u256_lt(h, BJJ_SUBORDER);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        A[i] = h_ge_l ? h[i] : BJJ_SUBORDER[i];
        B[i] = h_ge_l ? BJJ_SUBORDER[i] : h[i];
        ta[i] = 0;
        tb[i] = 0;
    }
    ta[0] = h_ge_l ? 1u : 0u;
    tb[0] = h_ge_l ? 0u : 1u;
    uint32_t bneg = h_ge_l ? 1u : 0u;
    int s = BJJ_CLZ32(A[7]);           // a >= l > 2^250: a bit shift normalises it
    u256_shl(A, A, s);
    u256_shl(B, B, s);
    // phase 1: Euclid until the smaller remainder b is below 2^126
#pragma unroll 1
    for (;;) {
        const int nb = (B[7] ? 256 - BJJ_CLZ32(B[7]) : u256_bitlen(B)) - s;      // bitlen(b); <= 0 for b == 0
        if (nb <= 126) break;
        const uint32_t bh = B[7];
        if (bh == 0) {
            // b is 2^31 times shorter than a (degenerate h): the estimate below would crawl; halve a instead
            split_step_pow2(A, ta, B, tb);
        } else {
            uint32_t q = bh == 0xFFFFFFFFu ? 1u : A[7] / (bh + 1u);
            q = q ? q : 1u;
            u256_mulsub(A, B, q);
            u256_muladd(ta, tb, q);
        }
        if (u256_lt(A, B)) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t t = A[i];
                A[i] = B[i];
                B[i] = t;
                t = ta[i];
                ta[i] = tb[i];
                tb[i] = t;
            }
            bneg ^= 1u;
        }
        // renormalise (A >= B, A != 0: the pair has gcd 1 or l)
#pragma unroll 1
        while (A[7] == 0) {              // rare: a lost 32 bits or more in one step
#pragma unroll
            for (int i = 7; i > 0; i--) {
                A[i] = A[i - 1];
                B[i] = B[i - 1];
            }
            A[0] = 0;
            B[0] = 0;
            s += 32;
        }
        const int lz = BJJ_CLZ32(A[7]);
        u256_shl(A, A, lz);
        u256_shl(B, B, lz);
        s += lz;
    }
    // back to plain integers
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = A[i];
        b[i] = B[i];
    }
#pragma unroll 1
    for (; s >= 32; s -= 32) {
#pragma unroll
        for (int i = 0; i < 7; i++) {
            a[i] = a[i + 1];
            b[i] = b[i + 1];
        }
        a[7] = 0;
        b[7] = 0;
    }
#if BJJ_DEVICE_CODE
#pragma unroll
    for (int i = 0; i < 7; i++) {
        a[i] = __funnelshift_r(a[i], a[i + 1], s);
        b[i] = __funnelshift_r(b[i], b[i + 1], s);
    }
    a[7] >>= s;
    b[7] >>= s;
#else
    if (s) {
        for (int i = 0; i < 7; i++) {
            a[i] = (a[i] >> s) | (a[i + 1] << (32 - s));
            b[i] = (b[i] >> s) | (b[i + 1] << (32 - s));
        }
        a[7] >>= s;
        b[7] >>= s;
    }
#endif
    // b < 2^126 <= a (or b = h was short from the start).  If its cofactor tb is odd, (b, tb) is the answer:
    // |tb| <= l / a < 2^125.  Otherwise ta is odd (the invariant sum is odd) and stays odd under ta += 2^k tb:
    // phase 2 keeps reducing a by b, without swapping, until a is short too.
    const bool second = !(tb[0] & 1u);
    if (second) {
#pragma unroll 1
        while (!split_small(a)) split_step_pow2(a, ta, b, tb);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u[i] = second ? a[i] : b[i];
        v[i] = second ? ta[i] : tb[i];
    }
    vneg = second ? (bneg ^ 1u) : bneg;
}

// r = a * b * 2^-256 mod l, not fully reduced: r < 2l whenever a*b < 2^256 * l.  Plain word-serial Montgomery
// (two calls per verification; speed is irrelevant here).
BJJ_HD void montmul_suborder(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a[j] * b[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        const uint32_t m = t[0] * BJJ_L_NINV32;
        c = (uint64_t)m * BJJ_SUBORDER[0] + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * BJJ_SUBORDER[j] + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = t[i];
}

// w = |v| * s mod l (any representative below 2l): s is any 256-bit integer, |v| <= l.
BJJ_HD void split_scale_s(uint32_t* w, const uint32_t* s, const uint32_t* v) {
    uint32_t t[8];
    montmul_suborder(t, s, BJJ_L_R2);     // s * 2^256 mod l   (s * R2 < 2^256 * l)
    montmul_suborder(w, t, v);            // s * |v| mod l     (t * |v| < 2l * l)
}

}  // namespace bjj
