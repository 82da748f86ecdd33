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

// One shift-and-subtract step of the division of d by b (d >= b), mirrored on the relation (a, ta):
// d -= 2^k b, a -= 2^k b, |ta| += 2^k |tb|  with the largest k that keeps d >= 0.
BJJ_HD void split_step_div(uint32_t* d, uint32_t* a, uint32_t* ta, const uint32_t* b, const uint32_t* tb) {
    int k = u256_bitlen(d) - u256_bitlen(b);
    uint32_t c[8], tc[8], t[8], half[8];
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
    const uint32_t over = sub256(t, d, c);       // borrow: c > d (then the shift was >= 1) and half of c is used
    u256_shr1(half, c);
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = over ? half[i] : c[i];
    u256_shr1(half, tc);
#pragma unroll
    for (int i = 0; i < 8; i++) tc[i] = over ? half[i] : tc[i];
    sub256(d, d, c);
    sub256(a, a, c);
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

// r = x * a - y * b   (the caller guarantees 0 <= x a - y b < 2^256; x, y < 2^32)
BJJ_HD void u256_lin_sub(uint32_t* r, uint32_t x, const uint32_t* a, uint32_t y, const uint32_t* b) {
    uint64_t c1 = 0, c2 = 0;
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c1 += (uint64_t)x * a[i];
        c2 += (uint64_t)y * b[i];
        const uint64_t d = (uint64_t)(uint32_t)c1 - (uint32_t)c2 - borrow;
        r[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 63);
        c1 >>= 32;
        c2 >>= 32;
    }
}
// r = x * a + y * b   (x, y < 2^31; no overflow of 256 bits)
BJJ_HD void u256_lin_add(uint32_t* r, uint32_t x, const uint32_t* a, uint32_t y, const uint32_t* b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x * a[i] + (uint64_t)y * b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
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
#ifndef BJJ_SPLIT_LEHMER
#define BJJ_SPLIT_LEHMER 1
#endif
#if BJJ_SPLIT_LEHMER
        // Lehmer: run Euclid on the leading words (A[7], B[7]) for as long as the TRUE remainders are certainly
        // non-negative and still >= 2^126, then apply the accumulated 2x2 cofactor matrix to the long numbers once --
        // ~8 quotients per multi-precision pass instead of one.  With r_0 = a, r_1 = b the remainders are
        // r_j = (-1)^j (x_j a - y_j b), x_j, y_j >= 0; cutting a and b to their leading words changes r_j by less than
        // max(x_j, y_j) units of the leading word, so  r^_j >= x_j + y_j (+ the 2^126 mark in those units)  keeps the
        // true r_j >= 0 (resp. >= 2^126).  The quotients need not be the true ones: ANY sequence of steps with
        // non-negative remainders preserves the invariants above, and the pair is re-ordered after the pass.
        {
            // true r >= 2^126  <=  (r << s) >= 2^(126+s)  <=  r^ >= 2^(s-98) + error   (r^ in units of 2^224)
            const uint32_t mark = s > 98 ? (s - 98 >= 32 ? 0xFFFFFFFFu : (1u << (s - 98))) : 1u;
            uint32_t rp = A[7], rc = bh, x0 = 1, y0 = 0, x1 = 0, y1 = 1;
            int steps = 0;
#pragma unroll 1
            while (rc != 0) {
                const uint32_t q = rp / rc, rn = rp - q * rc;
                const uint64_t xn = x0 + (uint64_t)q * x1, yn = y0 + (uint64_t)q * y1;
                if ((xn | yn) >> 31) break;
                if ((uint64_t)rn < xn + yn + mark) break;
                rp = rc;
                rc = rn;
                x0 = x1;
                y0 = y1;
                x1 = (uint32_t)xn;
                y1 = (uint32_t)yn;
                steps++;
            }
            if (steps > 0) {
                // (a, b) <- (r_steps, r_steps+1);  cofactor magnitudes x |ta| + y |tb|;  the signs alternate per step
                uint32_t na[8], nb2[8], nta[8], ntb[8];
                if (steps & 1) {
                    u256_lin_sub(na, y0, B, x0, A);
                    u256_lin_sub(nb2, x1, A, y1, B);
                } else {
                    u256_lin_sub(na, x0, A, y0, B);
                    u256_lin_sub(nb2, y1, B, x1, A);
                }
                u256_lin_add(nta, x0, ta, y0, tb);
                u256_lin_add(ntb, x1, ta, y1, tb);
                bneg ^= (uint32_t)(steps & 1);
                const bool sw = u256_lt(na, nb2);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    A[i] = sw ? nb2[i] : na[i];
                    B[i] = sw ? na[i] : nb2[i];
                    ta[i] = sw ? ntb[i] : nta[i];
                    tb[i] = sw ? nta[i] : ntb[i];
                }
                bneg ^= sw ? 1u : 0u;
                goto renormalise;
            }
        }
#endif
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
#if BJJ_SPLIT_LEHMER
    renormalise:
#endif
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
    if (second && !split_small(a)) {
        // The relations (a - j b, |ta| + j |tb|), j >= 0, all have an odd cofactor.  The smallest j that brings a - j b
        // under the 32-window limit LIM is  j = floor((a - LIM) / b) + 1  -- it leaves the SMALLEST cofactor among the
        // short ones (halving a greedily, as the first version did, overshoots j by up to 2x and with it |v|; 3.8 % of
        // random hm then needed a 33rd window, which its whole warp pays for).  d = a - LIM is divided by b by
        // shift-and-subtract, every step mirrored on (a, ta); then one more b comes off.
        uint32_t d[8];
        const uint32_t lim[8] = {0u, 0u, 0u, 0x70000000u, 0u, 0u, 0u, 0u};
        sub256(d, a, lim);
#pragma unroll 1
        while (!u256_lt(d, b)) split_step_div(d, a, ta, b, tb);
        sub256(a, a, b);               // a = LIM + d - b  in (0, LIM)
        add256(ta, ta, tb);
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
