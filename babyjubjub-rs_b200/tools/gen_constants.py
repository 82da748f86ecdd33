#!/usr/bin/env python3
"""Build-time generator for csrc/generated/bjj_consts.inc (Montgomery-form constants for the GPU).

Stand-alone on purpose: the product build must not import anything under oracle/.  tests/
cross-check the tables emitted here against the oracle's independent regeneration.

What the constants replace in the reference (paths into /root/reference):
  * Q, D, A, B8, SUBORDER                      src/lib.rs:28-58
  * Fr Montgomery parameters (ff_ce derive)    Cargo.toml:12, src/lib.rs:7
  * Poseidon round constants / MDS             poseidon-rs 0.0.8 tables behind src/lib.rs:59
"""
import functools
import os
import sys

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
A = 168700
D = 168696
B8X = 5299619240641551281634865583518297030282874472190772894086521144482721001553
B8Y = 16950150798460657717958625567821834550301663161624707787222815936182638968203
ORDER = 21888242871839275222246405745257275088614511777268538073601725287587578984328
SUBORDER = ORDER >> 3
R = (1 << 256) % Q
R_F = 8
R_P = {2: 56, 3: 57, 4: 56, 5: 60, 6: 60, 7: 63, 8: 64, 9: 63}


def inv(x):
    return pow(x % Q, Q - 2, Q)


def sqrt_mod_q(a):
    """Tonelli-Shanks for Q (Q-1 = 2^28 * odd); used only to derive sqrt(-a) at build time."""
    a %= Q
    assert pow(a, (Q - 1) // 2, Q) == 1
    s, e = Q - 1, 0
    while s % 2 == 0:
        s //= 2
        e += 1
    n = 2
    while pow(n, (Q - 1) // 2, Q) != Q - 1:
        n += 1
    x = pow(a, (s + 1) // 2, Q)
    b = pow(a, s, Q)
    g = pow(n, s, Q)
    r = e
    while b != 1:
        m, t = 0, b
        while t != 1:
            t = t * t % Q
            m += 1
        gs = pow(g, 1 << (r - m - 1), Q)
        g = gs * gs % Q
        x = x * gs % Q
        b = b * g % Q
        r = m
    return x


# ---- Poseidon constants: Grain LFSR, 80-bit state held in one Python int (bit 79 = b[0]) --------
class Grain:
    MASK = (1 << 80) - 1

    def __init__(self, t, r_p):
        st = 0
        for value, width in ((1, 2), (0, 4), (254, 12), (t, 12), (R_F, 10), (r_p, 10), ((1 << 30) - 1, 30)):
            st = (st << width) | value
        self.st = st
        for _ in range(160):
            self.clock()

    def clock(self):
        st = self.st
        tap = lambda k: (st >> (79 - k)) & 1
        nb = tap(62) ^ tap(51) ^ tap(38) ^ tap(23) ^ tap(13) ^ tap(0)
        self.st = ((st << 1) & self.MASK) | nb
        return nb

    def shrunk_bit(self):
        while True:
            keep = self.clock()
            out = self.clock()
            if keep:
                return out

    def draw(self, nbits=254):
        v = 0
        for _ in range(nbits):
            v = (v << 1) | self.shrunk_bit()
        return v


def poseidon_tables(t):
    g = Grain(t, R_P[t])
    n_rc = (R_F + R_P[t]) * t
    rc = []
    while len(rc) < n_rc:
        v = g.draw()
        if v < Q:
            rc.append(v)
    while True:
        seeds = [g.draw() % Q for _ in range(2 * t)]
        if len(set(seeds)) == 2 * t and all((x + y) % Q for x in seeds[:t] for y in seeds[t:]):
            break
    mds = [[inv(seeds[i] + seeds[t + j]) for j in range(t)] for i in range(t)]
    return rc, mds


# ---- optimised schedule: sparse partial-round matrices ----------------------------------------------
# The dense permutation  x <- M * sbox(x + c_r)  is algebraically rewritten (exact field arithmetic, so
# the hash is bit-identical):  in a partial round only lane 0 goes through x^5, so the block-diagonal
# part A' = diag(1, A_hat) of a round matrix A = S * A' commutes with the S-box layer and is pushed
# into the PREVIOUS round.  Working backwards from the last partial round leaves, per partial round, a
# sparse matrix S_r = [[a00, what], [v, I]] (2t-1 products instead of t^2) and ONE scalar constant k_r;
# the leftovers are a dense matrix P = A'_1 * M used by full round 3 and one vector added before the
# first partial round.  (Technique: Poseidon paper, appendix "optimised implementation".)
def _matmul(a, b):
    return [[sum(a[i][k] * b[k][j] for k in range(len(b))) % Q for j in range(len(b[0]))] for i in range(len(a))]


def _matvec(a, v):
    return [sum(x * y for x, y in zip(row, v)) % Q for row in a]


def _inverse(a):
    n = len(a)
    m = [row[:] + [int(i == j) for j in range(n)] for i, row in enumerate(a)]
    for c in range(n):
        p = next(r for r in range(c, n) if m[r][c] % Q)
        m[c], m[p] = m[p], m[c]
        iv = inv(m[c][c])
        m[c] = [x * iv % Q for x in m[c]]
        for r in range(n):
            if r != c and m[r][c]:
                f = m[r][c]
                m[r] = [(x - f * y) % Q for x, y in zip(m[r], m[c])]
    return [row[n:] for row in m]


@functools.lru_cache(maxsize=None)
def poseidon_optimized(t):
    rc, mds = poseidon_tables(t)
    rp = R_P[t]
    rounds = [rc[r * t:(r + 1) * t] for r in range(R_F + rp)]
    part = rounds[R_F // 2:R_F // 2 + rp]
    sparse = [None] * (rp + 1)
    chat = [None] * (rp + 2)
    a = [row[:] for row in mds]
    for r in range(rp, 0, -1):
        ahat = [row[1:] for row in a[1:]]
        ainv = _inverse(ahat)
        wrow = a[0][1:]
        what = [sum(wrow[k] * ainv[k][j] for k in range(t - 1)) % Q for j in range(t - 1)]
        sparse[r] = (a[0][0], what, [a[i][0] for i in range(1, t)])
        aprime = [[1] + [0] * (t - 1)] + [[0] + ahat[i] for i in range(t - 1)]
        chat[r] = _matvec(aprime, part[r - 1])
        a = _matmul(aprime, mds)
    # constants: lanes 1.. are pushed back to one vector; lane 0 keeps one scalar per round
    d = [0] * (t - 1)
    k = [0] * (rp + 2)
    for r in range(rp, 0, -1):
        if r + 1 <= rp:
            k[r + 1] = (chat[r + 1][0] - sum(x * y for x, y in zip(sparse[r][1], d))) % Q
        d = [(chat[r][1 + i] + d[i]) % Q for i in range(t - 1)]
    pre = [chat[1][0]] + d
    partial = [[k[r] if r > 1 else 0, sparse[r][0]] + sparse[r][1] + sparse[r][2] for r in range(1, rp + 1)]
    return {"full_rc": rounds[:R_F // 2] + rounds[R_F // 2 + rp:], "pre": pre, "partial": partial, "M": mds, "P": a}


def poseidon_optimized_eval(inputs):
    """reference evaluation of the optimised schedule (used by tests to pin it on the dense oracle)"""
    t = len(inputs) + 1
    o = poseidon_optimized(t)
    x = [0] + [v % Q for v in inputs]
    for r in range(R_F // 2):
        x = [pow((p + c) % Q, 5, Q) for p, c in zip(x, o["full_rc"][r])]
        x = _matvec(o["P"] if r == R_F // 2 - 1 else o["M"], x)
    x = [(p + c) % Q for p, c in zip(x, o["pre"])]
    for row in o["partial"]:
        x0 = pow((x[0] + row[0]) % Q, 5, Q)
        what, v = row[2:2 + t - 1], row[2 + t - 1:]
        n0 = (row[1] * x0 + sum(p * q for p, q in zip(what, x[1:]))) % Q
        x = [n0] + [(x[i] + v[i - 1] * x0) % Q for i in range(1, t)]
    for r in range(R_F // 2, R_F):
        x = [pow((p + c) % Q, 5, Q) for p, c in zip(x, o["full_rc"][r])]
        x = _matvec(o["M"], x)
    return x[0]


# ---- grouped partial rounds ---------------------------------------------------------------------
# In the sparse schedule the lanes 1..t-1 only ACCUMULATE between partial rounds:  x_i <- x_i + v_i * x0.  Over a
# group of K rounds they are therefore  S_i + sum_m v_(r+m),i * x0_m  with S_i the value at the start of the group, and
# lane 0 of round r+j is
#     a_(r+j) * x0_j + sum_(m<j) c_(j,m) * x0_m + sum_i what_(r+j),i * S_i,    c_(j,m) = sum_i what_(r+j),i * v_(r+m),i
# -- one dot product of t + j terms over values that are already reduced -- and the lanes 1..t-1 are brought up to date
# once per group by a K-term dot product each.  Per round (t = 6) that is 14.7 limb-products-and-reductions of 64 MACs
# instead of 17 (K = 3; K = 4 is no better).  Layout of one group of K rounds starting at round r:
#     for j in 0..K-1:  k_(r+j), a_(r+j), c_(j,j-1), .., c_(j,0), what_(r+j),1..t-1        (t + 1 + j elements)
#     for i in 1..t-1:  v_(r),i, .., v_(r+K-1),i                                           (K elements)
# A group of one round is the plain sparse round (k, a00, what, v).
POSEIDON_GROUP = 3
# Measured (profiles/r2_ab_poseidon_group.txt): the grouped rounds win for t >= 5 (-8..-10 % time); for t <= 4 the
# saving is small or negative in MACs (2t + (K-1)/2 + (t-1)/K against 3t - 1 units per round) and the three-round loop
# body costs more in instruction fetch than it saves, so those widths keep the plain sparse round.
POSEIDON_GROUP_MIN_T = 5


def poseidon_groups(t, group=POSEIDON_GROUP):
    rows = poseidon_optimized(t)["partial"]
    rp = len(rows)
    sizes = [group] * (rp // group) + ([rp % group] if rp % group else [])
    out, r = [], 0
    for K in sizes:
        g = []
        ks = [rows[r + j][0] for j in range(K)]
        a = [rows[r + j][1] for j in range(K)]
        what = [rows[r + j][2:2 + t - 1] for j in range(K)]
        v = [rows[r + j][2 + t - 1:] for j in range(K)]
        for j in range(K):
            c = [sum(w * x for w, x in zip(what[j], v[m])) % Q for m in range(j - 1, -1, -1)]
            g += [ks[j], a[j]] + c + what[j]
        for i in range(t - 1):
            g += [v[m][i] for m in range(K)]
        assert len(g) == 2 * K * t + K * (K - 1) // 2
        out.append((K, g))
        r += K
    return out


def poseidon_grouped_eval(inputs, group=POSEIDON_GROUP):
    """reference evaluation of the grouped schedule, reading the table exactly as the device does"""
    t = len(inputs) + 1
    o = poseidon_optimized(t)
    x = [0] + [v % Q for v in inputs]
    for r in range(R_F // 2):
        x = [pow((p + c) % Q, 5, Q) for p, c in zip(x, o["full_rc"][r])]
        x = _matvec(o["P"] if r == R_F // 2 - 1 else o["M"], x)
    x = [(p + c) % Q for p, c in zip(x, o["pre"])]
    for K, g in poseidon_groups(t, group):
        x0s, p = [], 0
        for j in range(K):
            x0 = pow((x[0] + g[p]) % Q, 5, Q)
            x0s.append(x0)
            ops = x0s[::-1] + x[1:]
            coef = g[p + 1:p + 1 + t + j]
            x[0] = sum(c * y for c, y in zip(coef, ops)) % Q
            p += t + 1 + j
        for i in range(1, t):
            x[i] = (x[i] + sum(c * y for c, y in zip(g[p:p + K], x0s))) % Q
            p += K
    for r in range(R_F // 2, R_F):
        x = [pow((p + c) % Q, 5, Q) for p, c in zip(x, o["full_rc"][r])]
        x = _matvec(o["M"], x)
    return x[0]


# ---- fixed exponents: sliding-window schedules ------------------------------------------------------
# a^e for a PUBLIC exponent, left to right with a window of 4 bits over the odd powers a, a^3, .., a^15: every step is
# "square n times, multiply by a^(2 idx + 1)"; the first step only selects its power.  For Q - 2 (Fermat inversion) that
# is 254 squarings + 8 + ~50 multiplications instead of 254 + 126; for (T - 1) / 2 (the square root) 225 + 8 + ~45
# instead of 225 + 99.
POW_WINDOW = 4


def pow_schedule(e, w=POW_WINDOW):
    bits = bin(e)[2:]
    steps, i, pending = [], 0, 0
    while i < len(bits):
        if bits[i] == "0":
            pending += 1
            i += 1
            continue
        j = min(i + w, len(bits))
        while bits[j - 1] == "0":
            j -= 1
        val = int(bits[i:j], 2)
        steps.append((pending + (j - i), (val - 1) // 2))
        pending = 0
        i = j
    return steps, pending


def pow_schedule_eval(a, e):
    """evaluates a^e mod Q the way csrc/fr.cuh::fr_pow_sched walks the schedule"""
    steps, tail = pow_schedule(e)
    a2 = a * a % Q
    tab = [a % Q]
    for _ in range((1 << (POW_WINDOW - 1)) - 1):
        tab.append(tab[-1] * a2 % Q)
    acc = tab[steps[0][1]]
    for nsq, idx in steps[1:]:
        for _ in range(nsq):
            acc = acc * acc % Q
        acc = acc * tab[idx] % Q
    for _ in range(tail):
        acc = acc * acc % Q
    return acc


# ---- emit ---------------------------------------------------------------------------------------
def limbs(x):
    assert 0 <= x < (1 << 256)
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def mont(x):
    return (x % Q) * R % Q


def fmt(x):
    return "{" + ",".join("0x%08xu" % w for w in limbs(x)) + "}"


def emit(out):
    w = out.write
    w("// GENERATED by babyjubjub-rs_b200/tools/gen_constants.py -- do not edit.\n")
    w("// All field elements: 8 x u32 little-endian limbs; *_M = Montgomery form (x * 2^256 mod Q).\n")
    w("#pragma once\n\n")
    for name, val in (("BJJ_Q", Q), ("BJJ_2Q", 2 * Q)):
        for i, l in enumerate(limbs(val)):
            w("#define %s%d 0x%08xu\n" % (name, i, l))
    w("#define BJJ_NINV32 0x%08xu   // -Q^-1 mod 2^32\n\n" % ((-pow(Q, -1, 1 << 32)) % (1 << 32)))

    s = sqrt_mod_q(-A)              # x' = s*x maps a*x^2+y^2=1+d*x^2*y^2 onto -x'^2+y^2=1+d'*x'^2*y^2
    s = min(s, Q - s)
    dp = (-D * inv(A)) % Q          # d' = -d/a
    assert (s * s + A) % Q == 0
    # sanity: B8 maps onto the a=-1 curve
    xb = s * B8X % Q
    assert (-xb * xb + B8Y * B8Y - 1 - dp * xb * xb % Q * B8Y * B8Y) % Q == 0
    singles = [
        ("Q", Q), ("TWOQ", 2 * Q), ("QHALF", Q >> 1), ("ONE_M", mont(1)), ("R2", R * R % Q),
        ("A_M", mont(A)), ("D_M", mont(D)),
        ("SQRT_NEG_A_M", mont(s)), ("INV_SQRT_NEG_A_M", mont(inv(s))),
        ("DP_M", mont(dp)), ("TWO_DP_M", mont(2 * dp)),
        ("B8X_M", mont(B8X)), ("B8Y_M", mont(B8Y)),
        ("SUBORDER", SUBORDER), ("ORDER", ORDER),
    ]
    for name, val in singles:
        w("BJJ_CONST uint32_t BJJ_%s[8] = %s;\n" % (name, fmt(val)))
    # Montgomery arithmetic modulo l = SUBORDER (verify's half-size scalar split)
    w("#define BJJ_L_NINV32 0x%08xu   // -l^-1 mod 2^32\n" % ((-pow(SUBORDER, -1, 1 << 32)) % (1 << 32)))
    w("BJJ_CONST uint32_t BJJ_L_R2[8] = %s;   // 2^512 mod l\n" % fmt(pow(2, 512, SUBORDER)))
    w("BJJ_CONST uint32_t BJJ_L_R1[8] = %s;   // 2^256 mod l\n" % fmt(pow(2, 256, SUBORDER)))
    w("BJJ_CONST uint32_t BJJ_L_R3[8] = %s;   // 2^768 mod l\n" % fmt(pow(2, 768, SUBORDER)))
    w("\n")

    # exponent bits for Fermat inversion (Q-2) and sqrt ((T-1)/2 with Q-1 = 2^28*T)
    T = (Q - 1) >> 28
    assert T % 2 == 1 and T << 28 == Q - 1
    w("BJJ_CONST uint32_t BJJ_EXP_QM2[8] = %s;\n" % fmt(Q - 2))
    w("BJJ_CONST uint32_t BJJ_EXP_TM1H[8] = %s;   // (T-1)/2, %d bits\n" % (fmt((T - 1) // 2), ((T - 1) // 2).bit_length()))
    w("#define BJJ_EXP_TM1H_BITS %d\n\n" % ((T - 1) // 2).bit_length())
    w("#define BJJ_POW_TABLE %d   // odd powers a, a^3, .. kept by fr_pow_sched\n" % (1 << (POW_WINDOW - 1)))
    for name, e in (("QM2", Q - 2), ("TM1H", (T - 1) // 2)):
        steps, tail = pow_schedule(e)
        assert all(n < 256 for n, _ in steps) and pow_schedule_eval(3, e) == pow(3, e, Q)
        w("// a^%s: %d steps of (squarings, index of the odd power), then %d squarings: %d squarings, %d multiplications\n"
          % (name, len(steps), tail, sum(n for n, _ in steps[1:]) + tail, len(steps) - 1 + (1 << (POW_WINDOW - 1))))
        w("#define BJJ_POW_%s_STEPS %d\n#define BJJ_POW_%s_TAIL %d\n" % (name, len(steps), name, tail))
        w("BJJ_CONST uint8_t BJJ_POW_%s[%d][2] = {%s};\n\n" % (name, len(steps), ",".join("{%d,%d}" % st for st in steps)))

    # 2-adic part of the square root: g = 5^T generates the order-2^28 subgroup.
    g = pow(5, T, Q)
    assert pow(g, 1 << 27, Q) == Q - 1
    # Pohlig-Hellman with four 7-bit digits.  For digit j (value v): table PH_NEG[j][v] = g^(-(v << 7j)),
    # table PH_HALF[j][v] = g^(-(v << 7j)/2) for the final root correction (only used when the dlog is
    # even; for j == 0 the entry uses v>>1), and a perfect-hash lookup from h^v (h = g^(2^21)) to v.
    h = pow(g, 1 << 21, Q)
    hv = [pow(h, v, Q) for v in range(128)]
    keys = [limbs(mont(x))[0] for x in hv]
    # find (mult, shift) with 512 slots such that ((key * mult) >> 23) is collision free
    mult = None
    for cand in range(1, 1 << 20, 2):
        slots = {((k * cand) & 0xFFFFFFFF) >> 23 for k in keys}
        if len(slots) == 128:
            mult = cand
            break
    assert mult is not None
    lut = [0xFF] * 512
    for v, k in enumerate(keys):
        lut[((k * mult) & 0xFFFFFFFF) >> 23] = v
    w("#define BJJ_PH_HASH_MULT 0x%08xu\n" % mult)
    w("BJJ_TABLE uint8_t BJJ_PH_LUT[512] = {%s};\n" % ",".join(str(x) for x in lut))
    w("BJJ_TABLE uint32_t BJJ_PH_NEG[4][128][8] = {\n")
    for j in range(4):
        w(" {\n")
        for v in range(128):
            w("  %s,\n" % fmt(mont(pow(g, (-(v << (7 * j))) % (1 << 28), Q))))
        w(" },\n")
    w("};\n")
    w("BJJ_TABLE uint32_t BJJ_PH_HALF[4][128][8] = {\n")
    for j in range(4):
        w(" {\n")
        for v in range(128):
            e = (v << (7 * j)) >> 1
            w("  %s,\n" % fmt(mont(pow(g, (-e) % (1 << 28), Q))))
        w(" },\n")
    w("};\n\n")

    w("#ifndef BJJ_POSEIDON_GROUP\n#define BJJ_POSEIDON_GROUP %d\n#endif\n" % POSEIDON_GROUP)
    w("#if BJJ_POSEIDON_GROUP != 1 && BJJ_POSEIDON_GROUP != %d\n#error \"the tables are generated for groups of 1 or %d partial rounds\"\n#endif\n\n"
      % (POSEIDON_GROUP, POSEIDON_GROUP))
    for t in range(2, 8):        # poseidon-rs 0.0.8 accepts 1..6 inputs
        o = poseidon_optimized(t)
        w("#define BJJ_POSEIDON_RP_%d %d\n" % (t, R_P[t]))
        space = "BJJ_CONST" if t == 6 else "BJJ_TABLE"   # t=6 (verify) sits in the constant bank

        def table(name, elems, note=""):
            w("%s uint32_t BJJ_POSEIDON_%s%d[%d][8] = {%s\n" % (space, name, t, len(elems), ("   // " + note) if note else ""))
            for e in elems:
                w(" %s,\n" % fmt(mont(e)))
            w("};\n")
        table("FC", [c for r in o["full_rc"] for c in r], "round constants of the 8 full rounds, [round][lane]")
        table("PRE", o["pre"], "vector added once before the partial rounds")
        w("#define BJJ_POSEIDON_GROUP_%d %s\n" % (t, "BJJ_POSEIDON_GROUP" if t >= POSEIDON_GROUP_MIN_T else "1"))
        w("#if BJJ_POSEIDON_GROUP_%d == 1\n" % t)
        table("PR", [e for r in o["partial"] for e in r],
              "per partial round: k, a00, what[1..t-1], v[1..t-1]  (2t elements)")
        w("#else\n")
        table("PR", [e for _, g in poseidon_groups(t) for e in g],
              "partial rounds in groups of %d (+ one shorter group): see gen_constants.py::poseidon_groups" % POSEIDON_GROUP)
        w("#endif\n")
        table("M", [o["M"][i][j] for i in range(t) for j in range(t)], "row-major MDS")
        table("P", [o["P"][i][j] for i in range(t) for j in range(t)], "row-major matrix of full round 3 (= A'_1 * M)")
        w("\n")


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "csrc", "generated", "bjj_consts.inc")
    if len(sys.argv) > 1:
        dst = sys.argv[1]
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    tmp = dst + ".tmp"
    with open(tmp, "w") as f:
        emit(f)
    os.replace(tmp, dst)
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
