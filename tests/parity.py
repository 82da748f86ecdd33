"""Parity checks shared by the host-emulation tests (CPU) and the GPU tests: a back end under test is
compared, bit for bit, with the oracle on the same seeded inputs.  Sizes are parameters so the CPU
suite stays small while the GPU suite runs the same cases wider."""
import random

import numpy as np

from common import (FIELD_EDGES, SCALAR_EDGES, O, Q, cases_to_arrays, oracle_verify, pack, random_points,
                    signature_cases, special_points, unpack)

STATUS = {O.ERR_Y_RANGE: 1, O.ERR_NO_INV: 2, O.ERR_NOT_SQUARE: 3}


def _eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        bad = np.nonzero((a.reshape(len(a), -1) != b.reshape(len(b), -1)).any(axis=1))[0]
        raise AssertionError("%s: %d/%d lanes differ, first at %s" % (what, len(bad), len(a), bad[:8]))


def check_fr(be, ref, n, seed=1):
    rnd = random.Random(seed)
    a = FIELD_EDGES + [rnd.randrange(Q) for _ in range(n - len(FIELD_EDGES))]
    b = FIELD_EDGES[::-1] + [rnd.randrange(Q) for _ in range(n - len(FIELD_EDGES))]
    A, B = pack(a), pack(b)
    for op in (0, 1, 2, 4):
        _eq(be.fr_op(op, A, B), ref.fr_op(op, A, B), "fr op %d" % op)
    # op 5: the dedicated squaring on inputs in [Q, 2Q) (lazy domain) must agree with the plain square
    _eq(be.fr_op(5, A, B), ref.fr_op(4, A, B), "fr sqr on the lazy domain")
    m = min(n, 256)
    _eq(be.fr_op(3, A[:m], B[:m]), ref.fr_op(3, A[:m], B[:m]), "fr inverse")
    # spot-check the C oracle itself against Python integers
    got = unpack(be.fr_op(0, A[:64], B[:64]))
    assert got == [x * y % Q for x, y in zip(a[:64], b[:64])]


def check_add(be, ref, n, seed=2):
    rnd = random.Random(seed)
    on, off = special_points()
    pts = [(p[0], p[1], 1) for p in on + off] + [(1, 2, 0), (0, 0, 0), (0, 1, 0)]
    pts += [(rnd.randrange(Q), rnd.randrange(Q), rnd.randrange(Q)) for _ in range(max(0, n - len(pts)))]
    qts = pts[::-1]
    ins = [pack([p[i] for p in pts]) for i in range(3)] + [pack([q[i] for q in qts]) for i in range(3)]
    got, exp = be.add(*ins), ref.add(*ins)
    for g, e, nm in zip(got, exp, "xyz"):
        _eq(g, e, "add " + nm)
    # reference KATs src/lib.rs:420-499 through add + affine
    assert list(zip(*[unpack(v) for v in got]))[0] == O.proj_add(pts[0], qts[0])
    ga, ea = be.affine(*got), ref.affine(*exp)
    for g, e, nm in zip(ga, ea, "xy"):
        _eq(g, e, "affine " + nm)


def check_mul_scalar(be, ref, n_random, seed=3):
    rnd = random.Random(seed)
    on, off = special_points()
    pts = on + off
    cases = [(p, k) for p in pts for k in SCALAR_EDGES[:8]]
    cases += [(on[0], k) for k in SCALAR_EDGES] + [(off[2], k) for k in SCALAR_EDGES]
    cases += [(on[0], 14035240266687799601661095864649209771790948434046947201833777492504781204499)]   # src/lib.rs:533
    rp = random_points(rnd, max(1, n_random // 4))
    cases += [(rnd.choice(rp + pts), rnd.randrange(1 << rnd.choice((3, 64, 251, 254, 256)))) for _ in range(n_random)]
    px, py, k = pack([c[0][0] for c in cases]), pack([c[0][1] for c in cases]), pack([c[1] for c in cases])
    got, exp = be.mul_scalar(px, py, k), ref.mul_scalar(px, py, k)
    _eq(got[0], exp[0], "mul_scalar x")
    _eq(got[1], exp[1], "mul_scalar y")
    # pin the checker itself on the Python oracle for a few lanes
    gx, gy = unpack(got[0]), unpack(got[1])
    for i in list(range(0, len(cases), max(1, len(cases) // 24)))[:24]:
        assert (gx[i], gy[i]) == O.mul_scalar(*cases[i]), i


def check_fixed_base(be, ref, n_random, seed=4):
    rnd = random.Random(seed)
    ks = SCALAR_EDGES + [rnd.randrange(1 << rnd.choice((8, 128, 251, 256))) for _ in range(n_random)]
    K = pack(ks)
    got, exp = be.fixed_base(K), ref.fixed_base(K)
    _eq(got[0], exp[0], "fixed_base x")
    _eq(got[1], exp[1], "fixed_base y")
    keys = np.frombuffer(b"".join([bytes.fromhex("0001020304050607080900010203040506070809000102030405060708090001")] +
                                  [rnd.randbytes(32) for _ in range(max(3, n_random // 4))] + [bytes(32), b"\xff" * 32]),
                         dtype=np.uint8).reshape(-1, 32)
    _eq(be.scalar_key(keys), ref.scalar_key(keys), "scalar_key")
    assert unpack(be.scalar_key(keys))[0] == 6466070937662820620902051049739362987537906109895538826186780010858059362905
    got, exp = be.public(keys), ref.public(keys)
    _eq(got[0], exp[0], "public x")
    _eq(got[1], exp[1], "public y")
    assert (unpack(got[0])[0], unpack(got[1])[0]) == (
        0x1d5ac1f31407018b7d413a4f52c8f74463b30e6ac2238220ad8b254de4eaa3a2,
        0x1e1de8a908826c3f9ac2e0ceee929ecd0caf3b99b3ef24523aaab796a6f733c4)        # src/lib.rs:710-718


def check_compress_decompress(be, ref, n_random, seed=5):
    rnd = random.Random(seed)
    on, off = special_points()
    pts = on + off + random_points(rnd, 8)
    px, py = pack([p[0] for p in pts]), pack([p[1] for p in pts])
    comp = be.compress(px, py)
    _eq(comp, ref.compress(px, py), "compress")
    assert comp[0].tobytes().hex() == "53b81ed5bffe9545b54016234682e7b2f699bd42a5e9eae27ff4051bc698ce85"   # src/lib.rs:587-590
    blobs = [O.compress(p) for p in on + random_points(rnd, 8)]
    blobs += [bytes.fromhex("b5328f8791d48f20bec6e481d91c7ada235f1facf22547901c18656b6c3e042f"),            # src/lib.rs:598
              bytes.fromhex("70552d3ff548e09266ded29b33ce75139672b062b02aa66bb0d9247ffecf1d0b")]            # src/lib.rs:617
    blobs += [int(v).to_bytes(32, "little") for v in
              (0, 1, Q - 1, Q, Q + 1, 2, Q - 2, (1 << 255) | 1, (1 << 255) | (Q - 1), (1 << 255), (1 << 256) - 1,
               (1 << 255) | 5, 5)]
    blobs += [rnd.randbytes(32) for _ in range(n_random)]
    C = np.frombuffer(b"".join(blobs), dtype=np.uint8).reshape(-1, 32)
    got, exp = be.decompress(C), ref.decompress(C)
    _eq(got[2], exp[2], "decompress status")
    _eq(got[0], exp[0], "decompress x")
    _eq(got[1], exp[1], "decompress y")
    # checker vs the Python oracle on the named edge cases
    gx, gy, st = unpack(got[0]), unpack(got[1]), list(got[2])
    for i in range(len(blobs) - n_random):
        try:
            e, s = O.decompress_point(blobs[i]), 0
        except ValueError as ex:
            e, s = (0, 0), STATUS[str(ex)]
        assert (gx[i], gy[i], st[i]) == (e[0], e[1], s), i
    assert set(st) >= {0, 1, 3}


def check_poseidon(be, ref, n, seed=6, widths=range(1, 7)):
    rnd = random.Random(seed)
    for nin in widths:
        cols = [[rnd.randrange(Q) for _ in range(n)] for _ in range(nin)]
        for j in range(nin):
            cols[j][0] = j + 1
            cols[j][1] = Q - 1
            cols[j][2] = 0
        ins = [pack(c) for c in cols]
        got = be.poseidon(ins)
        _eq(got, ref.poseidon(ins), "poseidon nin=%d" % nin)
        assert unpack(got)[0] == O.poseidon(list(range(1, nin + 1)))
    if 1 in widths:
        assert unpack(be.poseidon([pack([1])]))[0] == 0x29176100eaa962bdc1fe6c654d6a3c130e96a4d1168b33848b897dc502820133


def check_verify(be, ref, n_valid, seed=7):
    rnd = random.Random(seed)
    cases = signature_cases(rnd, n_valid)
    arrs = cases_to_arrays(cases)
    got = be.verify(*arrs)
    exp = np.array([oracle_verify(c) for c in cases], dtype=np.uint8)
    _eq(got, exp, "verify vs python oracle")
    _eq(got, ref.verify(*arrs), "verify vs C oracle")
    assert exp[0] == 1 and 0 in exp
    # compressed pipeline
    sig64 = np.frombuffer(b"".join(O.compress((c[0], c[1])) + (c[2] & ((1 << 256) - 1)).to_bytes(32, "little") for c in cases),
                          dtype=np.uint8).reshape(-1, 64)
    pk32 = np.frombuffer(b"".join(O.compress((c[3], c[4])) for c in cases), dtype=np.uint8).reshape(-1, 32)
    # extra lanes with undecodable R8 / A
    bad = [int(v).to_bytes(32, "little") for v in (Q, 1, (1 << 255) | 7)]
    extra_sig = np.frombuffer(b"".join(b + bytes(32) for b in bad), dtype=np.uint8).reshape(-1, 64)
    sig64 = np.concatenate([sig64, extra_sig, sig64[:3]])
    pk32 = np.concatenate([pk32, pk32[:3], np.frombuffer(b"".join(bad), dtype=np.uint8).reshape(-1, 32)])
    msg = np.concatenate([arrs[5], arrs[5][:3], arrs[5][:3]])
    gok, gst = be.verify_compressed(sig64, pk32, msg)
    eok, est = ref.verify_compressed(sig64, pk32, msg)
    _eq(gst, est, "verify_compressed status")
    _eq(gok, eok, "verify_compressed ok")
    assert gok[0] == 1 and set(gst) >= {0, 1, 3}


def check_sign(be, ref, n_random, seed=8):
    """PrivateKey::sign (src/lib.rs:308-342) incl. the circomlib vector (src/lib.rs:720-735) and msg > Q"""
    from common import KEY_KAT, MSG_KAT
    rnd = random.Random(seed)
    keys = [KEY_KAT] + [rnd.randbytes(32) for _ in range(n_random)] + [bytes(32), b"\xff" * 32, KEY_KAT, KEY_KAT, KEY_KAT]
    msgs = [MSG_KAT] + [rnd.randrange(Q) for _ in range(n_random)] + [0, Q - 1, Q, Q + 1, (1 << 256) - 1]
    K = np.frombuffer(b"".join(keys), dtype=np.uint8).reshape(-1, 32)
    M = pack(msgs)
    got, exp = be.sign(K, M), ref.sign(K, M)
    for g, e, nm in zip(got, exp, ("r8x", "r8y", "s", "status")):
        _eq(g, e, "sign " + nm)
    assert unpack(got[0])[0] == 0x192b4e51adf302c8139d356d0e08e2404b5ace440ef41fc78f5c4f2428df0765
    assert unpack(got[2])[0] == 1672775540645840396591609181675628451599263765380031905495115170613215233181
    assert list(got[3][-3:]) == [0, 4, 4]
    # sign -> verify round trip on the back end under test (reference tests :554-572)
    ok = [i for i in range(len(keys)) if got[3][i] == 0]
    ax, ay = be.public(K[ok])
    v = be.verify(got[0][ok], got[1][ok], got[2][ok], ax, ay, M[ok])
    assert v.all()


def check_schnorr(be, ref, n_valid, seed=9):
    """verify_schnorr / schnorr_hash (src/lib.rs:364-385): round trips as in the reference's
    test_schnorr_signature (:677-686, random nonce, so no KAT) plus the negative and off-curve cases"""
    rnd = random.Random(seed)
    on, off = special_points()
    cases = []          # [pkx, pky, msg, rx, ry, s (reduced mod SUBORDER for the 256-bit ABI)]
    full = []           # the same with the reference's unreduced s, for the Python oracle
    for _ in range(n_valid):
        key, m, k = rnd.randbytes(32), rnd.randrange(Q), rnd.getrandbits(1024)
        r, s = O.sign_schnorr(key, m, k)
        pk = O.public(key)
        full.append((pk, m, r, s))
        cases.append([pk[0], pk[1], m, r[0], r[1], s % O.SUBORDER])
    base, bfull = cases[0], full[0]
    def mod(i, v):
        c = list(base)
        c[i] = v
        return c
    cases.append(mod(5, base[5] ^ 1))                           # wrong s
    cases.append(mod(2, base[2] ^ 4))                           # wrong msg
    cases.append(mod(5, base[5] + O.SUBORDER))                  # s + SUBORDER: still valid
    cases.append(mod(2, Q + 1))                                 # Err: msg outside the field
    cases.append(mod(2, (1 << 256) - 1))
    cases.append(mod(0, (base[0] + 1) % Q))                     # off-curve pk  -> literal ladder
    cases.append(mod(3, (base[3] + 1) % Q))                     # off-curve r   -> literal final add
    c = list(base); c[0], c[1] = 0, 0; cases.append(c)
    c = list(base); c[3], c[4] = 0, 0; cases.append(c)
    c = list(base); c[0], c[1] = cases[1][0], cases[1][1]; cases.append(c)      # foreign pk
    s0 = rnd.randrange(O.SUBORDER)
    R = O.mul_scalar(O.B8, s0)
    cases.append([0, 1, 5, R[0], R[1], s0])                     # pk = identity: s*B8 == r
    arrs = cases_to_arrays(cases)
    gok, gst = be.verify_schnorr(*arrs)
    eok, est = ref.verify_schnorr(*arrs)
    _eq(gst, est, "verify_schnorr status")
    _eq(gok, eok, "verify_schnorr ok")
    assert list(gok[:n_valid]) == [1] * n_valid and 0 in gok and 4 in gst
    # the checker against the Python oracle, with the reference's unreduced s for the valid ones
    for i, (pk, m, r, s) in enumerate(full):
        assert O.verify_schnorr(pk, m, r, s) is True and gok[i] == 1
    for i in range(n_valid, len(cases)):
        c = cases[i]
        try:
            e, st = int(O.verify_schnorr((c[0], c[1]), c[2], (c[3], c[4]), c[5])), 0
        except ValueError:
            e, st = 0, 4
        assert (gok[i], gst[i]) == (e, st), i
