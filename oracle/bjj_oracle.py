"""CPU oracle (pure Python) for the babyjubjub-rs hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product path (babyjubjub-rs_b200/, libbjj_cuda.so) never does.  It is a *restatement* of the
reference algorithm on Python integers, each function citing the reference lines it follows
(paths are into /root/reference).  Parity is PINNED: `self_check()` replays every known-answer
test the reference holds for this path (src/lib.rs:420-552, 574-632, 688-738; src/utils.rs:229-260).

Third-party arithmetic that is not under /root/reference (no Cargo.lock in the tree):
  * ff_ce ^0.11 `Fr`      -> exact integers mod Q (canonical values are bit-identical by definition)
  * poseidon-rs =0.0.8     -> `poseidon()` below + oracle/poseidon_constants.py
  * blake-hash ^0.4.0      -> `blake512()` below (original BLAKE-512)
  * num-bigint ^0.4        -> Python int
"""
import struct

from . import poseidon_constants as _pc

# ---- constants: src/lib.rs:28-58 ------------------------------------------------------------
Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
D = 168696
A = 168700
B8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
      16950150798460657717958625567821834550301663161624707787222815936182638968203)
ORDER = 21888242871839275222246405745257275088614511777268538073601725287587578984328
SUBORDER = ORDER >> 3
Q_HALF = Q >> 1

ERR_Y_RANGE = "y outside the Finite Field over R"      # src/lib.rs:202
ERR_NO_INV = "no mod inv of Zero"                      # src/utils.rs:14
ERR_NOT_SQUARE = "not a mod p square"                  # src/utils.rs:119


# ---- src/utils.rs -----------------------------------------------------------------------------
def modulus(a, m):                                      # src/utils.rs:7-9
    return ((a % m) + m) % m


def modinv(a, q):                                       # src/utils.rs:11-29
    if a == 0:
        raise ValueError(ERR_NO_INV)
    mn = (q, a)
    xy = (0, 1)
    while mn[1] != 0:
        # Rust BigInt `/` truncates toward zero; operands here are non-negative so // is identical
        xy = (xy[1], xy[0] - (mn[0] // mn[1]) * xy[1])
        mn = (mn[1], modulus(mn[0], mn[1]))
    x = xy[0]
    while x < 0:
        x = modulus(x, q)
    return x


def legendre_symbol(a, q):                              # src/utils.rs:215-223
    ls = pow(a, (q - 1) >> 1, q)
    return -1 if ls == q - 1 else 1


def modsqrt(a, q):                                      # src/utils.rs:109-160
    if legendre_symbol(a, q) != 1 or a == 0 or q == 2:
        raise ValueError(ERR_NOT_SQUARE)
    if q % 4 == 3:
        return pow(a, (q + 1) // 4, q)
    s = q - 1
    e = 0
    while s % 2 == 0:
        s >>= 1
        e += 1
    n = 2
    while legendre_symbol(n, q) != -1:
        n += 1
    y = pow(a, (s + 1) >> 1, q)
    b = pow(a, s, q)
    g = pow(n, s, q)
    r = e
    while True:
        t = b
        m = 0
        while t != 1:
            t = modulus(t * t, q)
            m += 1
        if m == 0:
            return y
        t = pow(g, pow(2, r - m - 1, q), q)
        g = pow(g, pow(2, r - m, q), q)
        y = modulus(y * t, q)
        b = modulus(b * g, q)
        r = m


# ---- curve: src/lib.rs:62-190 -------------------------------------------------------------------
def proj_add(p, q):
    """PointProjective::add, EFD add-2008-bbjlp, also used for doubling.  src/lib.rs:88-131."""
    x1, y1, z1 = p
    x2, y2, z2 = q
    a = z1 * z2 % Q
    b = a * a % Q
    c = x1 * x2 % Q
    d = y1 * y2 % Q
    e = D * c % Q * d % Q
    f = (b - e) % Q
    g = (b + e) % Q
    aux = ((x1 + y1) * (x2 + y2) - c - d) % Q
    x3 = a * f % Q * aux % Q
    y3 = a * g % Q * ((d - A * c) % Q) % Q
    z3 = f * g % Q
    return (x3, y3, z3)


def proj_affine(p):
    """PointProjective::affine; Z == 0 -> (0, 0).  src/lib.rs:70-85."""
    x, y, z = p
    if z % Q == 0:
        return (0, 0)
    zinv = pow(z, Q - 2, Q)
    return (x * zinv % Q, y * zinv % Q)


def mul_scalar(p, n):
    """Point::mul_scalar: LSB-first double-and-add on |n|, unreduced.  src/lib.rs:149-164."""
    n = abs(n)                                          # :156 drops the sign
    r = (0, 1, 1)
    exp = (p[0], p[1], 1)
    for i in range(n.bit_length()):
        if (n >> i) & 1:
            r = proj_add(r, exp)
        exp = proj_add(exp, exp)
    return proj_affine(r)


def compress(p):
    """Point::compress: y little-endian, bit 255 = (x > Q>>1).  src/lib.rs:166-178."""
    b = bytearray(p[1].to_bytes(32, "little"))
    if p[0] > Q_HALF:
        b[31] |= 0x80
    return bytes(b)


def decompress_point(bb):
    """decompress_point.  src/lib.rs:192-224.  Raises ValueError(<reference error string>)."""
    b = bytearray(bb)
    sign = False
    if b[31] & 0x80:
        sign = True
        b[31] &= 0x7F
    y = int.from_bytes(b, "little")
    if y >= Q:
        raise ValueError(ERR_Y_RANGE)
    den = modinv(modulus(A - modulus(D * (y * y), Q), Q), Q)
    x = modulus((1 - modulus(y * y, Q)) * den, Q)
    x = modsqrt(x, Q)
    if (sign and x <= Q_HALF) or ((not sign) and x > Q_HALF):
        x = -x
    x = modulus(x, Q)
    return (x, y)


def on_curve(p):
    x, y = p
    return (A * x * x + y * y - 1 - D * x * x % Q * y * y) % Q == 0


# ---- BLAKE-512 (blake-hash 0.4.0; reference call sites src/lib.rs:226-237) ---------------------
_BL_IV = [0x6A09E667F3BCC908, 0xBB67AE8584CAA73B, 0x3C6EF372FE94F82B, 0xA54FF53A5F1D36F1,
          0x510E527FADE682D1, 0x9B05688C2B3E6C1F, 0x1F83D9ABFB41BD6B, 0x5BE0CD19137E2179]
_BL_C = [0x243F6A8885A308D3, 0x13198A2E03707344, 0xA4093822299F31D0, 0x082EFA98EC4E6C89,
         0x452821E638D01377, 0xBE5466CF34E90C6C, 0xC0AC29B7C97C50DD, 0x3F84D5B5B5470917,
         0x9216D5D98979FB1B, 0xD1310BA698DFB5AC, 0x2FFD72DBD01ADFB7, 0xB8E1AFED6A267E96,
         0xBA7C9045F12C7F99, 0x24A19947B3916CF7, 0x0801F2E2858EFC16, 0x636920D871574E69]
_BL_SIGMA = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15],
    [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
    [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4],
    [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
    [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13],
    [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
    [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11],
    [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
    [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5],
    [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0],
]
_M64 = (1 << 64) - 1


def _ror64(x, n):
    return ((x >> n) | (x << (64 - n))) & _M64


def _blake512_compress(h, block, t):
    m = struct.unpack(">16Q", block)
    v = list(h) + [_BL_C[0], _BL_C[1], _BL_C[2], _BL_C[3],
                   _BL_C[4] ^ (t & _M64), _BL_C[5] ^ (t & _M64),
                   _BL_C[6] ^ (t >> 64), _BL_C[7] ^ (t >> 64)]

    def g(a, b, c, d, r, i):
        s = _BL_SIGMA[r % 10]
        v[a] = (v[a] + v[b] + (m[s[2 * i]] ^ _BL_C[s[2 * i + 1]])) & _M64
        v[d] = _ror64(v[d] ^ v[a], 32)
        v[c] = (v[c] + v[d]) & _M64
        v[b] = _ror64(v[b] ^ v[c], 25)
        v[a] = (v[a] + v[b] + (m[s[2 * i + 1]] ^ _BL_C[s[2 * i]])) & _M64
        v[d] = _ror64(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & _M64
        v[b] = _ror64(v[b] ^ v[c], 11)

    for r in range(16):
        g(0, 4, 8, 12, r, 0)
        g(1, 5, 9, 13, r, 1)
        g(2, 6, 10, 14, r, 2)
        g(3, 7, 11, 15, r, 3)
        g(0, 5, 10, 15, r, 4)
        g(1, 6, 11, 12, r, 5)
        g(2, 7, 8, 13, r, 6)
        g(3, 4, 9, 14, r, 7)
    return [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]


def blake512(data):
    """Original BLAKE-512 (SHA-3 finalist, zero salt) of `data`."""
    data = bytes(data)
    bitlen = len(data) * 8
    h = list(_BL_IV)
    nfull = len(data) // 128
    t = 0
    for i in range(nfull):
        t += 1024
        h = _blake512_compress(h, data[128 * i:128 * (i + 1)], t)
    rem = data[128 * nfull:]
    length = bitlen.to_bytes(16, "big")
    if len(rem) <= 111:
        pad = bytearray(rem) + bytearray(112 - len(rem))
        pad[len(rem)] |= 0x80
        pad[111] |= 0x01
        # a block that holds only padding is compressed with counter 0
        h = _blake512_compress(h, bytes(pad) + length, bitlen if len(rem) else 0)
    else:
        pad = bytearray(rem) + bytearray(128 - len(rem))
        pad[len(rem)] |= 0x80
        h = _blake512_compress(h, bytes(pad), bitlen)
        pad2 = bytearray(112)
        pad2[111] |= 0x01
        h = _blake512_compress(h, bytes(pad2) + length, 0)
    return struct.pack(">8Q", *h)


# ---- Poseidon (poseidon-rs 0.0.8; call sites src/lib.rs:333,370,401) ---------------------------
def poseidon(inputs):
    n = len(inputs)
    # poseidon-rs 0.0.8 hash(): `if inp.is_empty() || inp.len() >= self.constants.n_rounds_p.len() - 1` with an 8-entry
    # round table, i.e. 1..6 inputs (t <= 7) [memory; the crate is not vendored -- ADVICE round 1]
    if n == 0 or n > 6:
        raise ValueError("Wrong inputs length")
    t = n + 1
    C, M = _pc.constants(t)
    r_p = _pc.R_P_TABLE[t - 2]
    state = [0] + [x % Q for x in inputs]
    for r in range(_pc.R_F + r_p):
        state = [(state[i] + C[r * t + i]) % Q for i in range(t)]
        if r < _pc.R_F // 2 or r >= _pc.R_F // 2 + r_p:
            state = [pow(s, 5, Q) for s in state]
        else:
            state[0] = pow(state[0], 5, Q)
        state = [sum(M[i][j] * state[j] for j in range(t)) % Q for i in range(t)]
    return state[0]


# ---- keys / EdDSA: src/lib.rs:284-342, 395-412 --------------------------------------------------
def scalar_key(key32):                                  # src/lib.rs:284-302
    h = bytearray(blake512(key32)[:32])
    h[0] &= 0xF8
    h[31] &= 0x7F
    h[31] |= 0x40
    return int.from_bytes(h, "little") >> 3


def public(key32):                                      # src/lib.rs:304-306
    return mul_scalar(B8, scalar_key(key32))


def sign(key32, msg):                                   # src/lib.rs:308-342
    if msg > Q:
        raise ValueError("msg outside the Finite Field")
    h = blake512(key32)
    msg32 = msg.to_bytes(32, "little") if msg < (1 << 256) else None
    r = int.from_bytes(blake512(h[32:64] + msg32), "little")
    r = modulus(r, SUBORDER)
    r8 = mul_scalar(B8, r)
    a = public(key32)
    hm = poseidon([r8[0], r8[1], a[0], a[1], msg % Q])
    s = (r + hm * (scalar_key(key32) << 3)) % SUBORDER
    return (r8, s)


def verify(pk, sig, msg):                               # src/lib.rs:395-412
    r8, s = sig
    if msg > Q:
        return False
    hm = poseidon([r8[0], r8[1], pk[0], pk[1], msg % Q])   # Fr::from_str reduces mod Q (msg == Q -> 0)
    l = mul_scalar(B8, s)
    k_a = mul_scalar(pk, 8 * hm)
    r = proj_add((r8[0], r8[1], 1), (k_a[0], k_a[1], 1))
    return l == proj_affine(r)


# ---- Schnorr: src/lib.rs:344-385 ---------------------------------------------------------------
def schnorr_hash(pk, msg, c):                           # src/lib.rs:364-373
    if msg > Q:
        raise ValueError("msg outside the Finite Field")
    return poseidon([pk[0], pk[1], c[0], c[1], msg % Q])


def sign_schnorr(key32, m, k):
    """src/lib.rs:345-362 with the 1024-bit nonce `k` supplied by the caller (the reference draws it from
    thread_rng, so there is no known-answer vector); s = k + scalar_key * h is NOT reduced."""
    r = mul_scalar(B8, k)
    pk = public(key32)
    h = schnorr_hash(pk, m, r)
    return r, k + scalar_key(key32) * h


def verify_schnorr(pk, m, r, s):                        # src/lib.rs:375-385
    sg = mul_scalar(B8, s)
    h = schnorr_hash(pk, m, r)
    pk_h = mul_scalar(pk, h)
    right = proj_add((r[0], r[1], 1), (pk_h[0], pk_h[1], 1))
    return sg == proj_affine(right)


def compress_signature(sig):                            # src/lib.rs:245-257
    r8, s = sig
    return compress(r8) + (s & ((1 << 256) - 1)).to_bytes(32, "little")


def decompress_signature(b64):                          # src/lib.rs:260-268
    return (decompress_point(b64[:32]), int.from_bytes(b64[32:], "little"))


# ---- pinning: every known-answer test of the reference for this path ----------------------------
def self_check():
    P = (17777552123799933955779906779655732241715742912184938656739573121738514868268,
         2626589144620713026669568689430873010625803728049924121243784502389097019475)
    P2 = (16540640123574156134436876038791482806971768689494387082833631921987005038935,
          20819045374670962167435360035096875258406992893633759881276124905556507972311)
    pj = lambda p: (p[0], p[1], 1)
    # src/lib.rs:420-459
    assert proj_affine(proj_add(pj(P), pj(P))) == (
        6890855772600357754907169075114257697580319025794532037257385534741338397365,
        4338620300185947561074059802482547481416142213883829469920100239455078257889)
    # src/lib.rs:460-499
    assert proj_affine(proj_add(pj(P), pj(P2))) == (
        7916061937171219682591368294088513039687205273691143098332585753343424131937,
        14035240266687799601661095864649209771790948434046947201833777492504781204499)
    # src/lib.rs:501-552
    m3 = mul_scalar(P, 3)
    assert m3 == proj_affine(proj_add(proj_add(pj(P), pj(P)), pj(P)))
    assert m3 == (19372461775513343691590086534037741906533799473648040012278229434133483800898,
                  9458658722007214007257525444427903161243386465067105737478306991484593958249)
    assert mul_scalar(P, 14035240266687799601661095864649209771790948434046947201833777492504781204499) == (
        17070357974431721403481313912716834497662307308519659060910483826664480189605,
        4014745322800118607127020275658861516666525056516280575712425373174125159339)
    # src/lib.rs:574-594
    c = compress(P)
    assert c.hex() == "53b81ed5bffe9545b54016234682e7b2f699bd42a5e9eae27ff4051bc698ce85"
    assert decompress_point(c) == P
    # src/lib.rs:596-632
    for yb, xb in (("b5328f8791d48f20bec6e481d91c7ada235f1facf22547901c18656b6c3e042f",
                    "b86cc8d9c97daef0afe1a4753c54fb2d8a530dc74c7eee4e72b3fdf2496d2113"),
                   ("70552d3ff548e09266ded29b33ce75139672b062b02aa66bb0d9247ffecf1d0b",
                    "30f1635ba7d56f9cb32c3ffbe6dca508a68c7f43936af11a23c785ce98cb3404")):
        assert decompress_point(bytes.fromhex(yb))[0] == int.from_bytes(bytes.fromhex(xb), "little")
    # src/lib.rs:688-738
    key = bytes.fromhex("0001020304050607080900010203040506070809000102030405060708090001")
    assert blake512(key).hex() == (
        "c992db23d6290c70ffcc02f7abeb00b9d00fa8b43e55d7949c28ba6be7545d32"
        "53882a61bd004a236ef1cdba01b27ba0aedfb08eefdbfb7c19657c880b43ddf1")
    assert scalar_key(key) == 6466070937662820620902051049739362987537906109895538826186780010858059362905
    pk = public(key)
    assert pk == (0x1d5ac1f31407018b7d413a4f52c8f74463b30e6ac2238220ad8b254de4eaa3a2,
                  0x1e1de8a908826c3f9ac2e0ceee929ecd0caf3b99b3ef24523aaab796a6f733c4)
    msg = int.from_bytes(bytes.fromhex("00010203040506070809"), "little")
    sig = sign(key, msg)
    assert sig[0] == (0x192b4e51adf302c8139d356d0e08e2404b5ace440ef41fc78f5c4f2428df0765,
                      0x2202bebcf57b820863e0acc88970b6ca7d987a0d513c2ddeb42e3f5d31b4eddf)
    assert sig[1] == 1672775540645840396591609181675628451599263765380031905495115170613215233181
    assert verify(pk, sig, msg) is True
    # src/utils.rs:229-260
    assert modinv(123456789123456789123456789123456789123456789, 12345678) == 641883
    assert modsqrt(6536923810004159332831702809452452174451353762940761092345538667656658715568,
                   7237005577332262213973186563042994240857116359379907606001950938285454250989) == \
        5464794816676661649783249706827271879994893912039750480019443499440603127256
    # Poseidon vectors recalled from poseidon-rs / go-iden3-crypto test-suites
    assert poseidon([1]) == 0x29176100eaa962bdc1fe6c654d6a3c130e96a4d1168b33848b897dc502820133
    assert poseidon([1, 2]) == 7853200120776062878684798364095072458815029376092732009249414926327459813530
    assert poseidon([1, 2, 0, 0, 0]) == 1018317224307729531995786483840663576608797660851238720571059489595066344487
    assert poseidon([1, 2, 3, 4, 5, 6]) == 20400040500897583745843009878988256314335038853985262692600694741116813247201
    return True


if __name__ == "__main__":
    self_check()
    print("oracle self_check: all reference known-answer tests reproduced")
