"""Shared test plumbing: three interchangeable back ends behind one interface.

    OracleC   oracle/_gen/liboracle.so   C++ restatement of the reference algorithm (the checker)
    HostEmu   tests/hostemu              the DEVICE headers compiled for the host (logic check, no GPU)
    Gpu       libbjj_cuda.so             the product, through the C ABI

All take / return numpy uint8 arrays of shape (n, 32) (little-endian 32-byte elements).
"""
import ctypes
import os
import random
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bjj_oracle as O  # noqa: E402

Q = O.Q
_sz = ctypes.c_size_t


def pack(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).reshape(-1, 32).copy()


def unpack(arr):
    raw = np.ascontiguousarray(arr, dtype=np.uint8).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, w=32):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(-1, w)


# ---------------------------------------------------------------------------------------------------
def build_hostemu():
    d = os.path.join(ROOT, "tests", "hostemu")
    so = os.path.join(d, "libbjj_hostemu.so")
    src = os.path.join(d, "hostemu.cpp")
    csrc = os.path.join(ROOT, "babyjubjub-rs_b200", "csrc")
    sys.path.insert(0, os.path.join(ROOT, "babyjubjub-rs_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bjj_build", os.path.join(ROOT, "babyjubjub-rs_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.generate_constants()
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")] + [b.GEN]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so + ".tmp", src], cwd=d)
        os.replace(so + ".tmp", so)
    return so


def build_oracle_c():
    from oracle import build as ob
    return ob.build()


class _CBackend:
    """common ctypes glue for the two C libraries that share the (n, arrays...) calling convention"""
    prefix = ""
    has_threads = False

    def _call(self, name, n, ins, out_shapes, extra_pre=(), dtype_last=None):
        ins = [_c(a, a.shape[-1] if a.ndim == 2 else 32) for a in ins]
        outs = [np.zeros(s, dtype=np.uint8) for s in out_shapes]
        args = list(extra_pre) + [_sz(n)] + [_p(a) for a in ins] + [_p(o) for o in outs]
        if self.has_threads:
            args.append(ctypes.c_int(self.threads))
        ret = getattr(self.lib, self.prefix + name)(*args)
        return outs, ret

    def fr_op(self, op, a, b):
        outs, _ = self._call_fr(op, a, b)
        return outs

    def add(self, px, py, pz, qx, qy, qz):
        n = len(px)
        outs, _ = self._call(self.names["add"], n, [px, py, pz, qx, qy, qz], [(n, 32)] * 3)
        return outs

    def affine(self, px, py, pz):
        n = len(px)
        outs, _ = self._call(self.names["affine"], n, [px, py, pz], [(n, 32)] * 2)
        return outs

    def mul_scalar(self, px, py, k):
        n = len(px)
        outs, _ = self._call(self.names["mul_scalar"], n, [px, py, k], [(n, 32)] * 2)
        return outs

    def fixed_base(self, k):
        n = len(k)
        outs, _ = self._call(self.names["fixed_base"], n, [k], [(n, 32)] * 2)
        return outs

    def public(self, keys):
        n = len(keys)
        outs, _ = self._call(self.names["public"], n, [keys], [(n, 32)] * 2)
        return outs

    def scalar_key(self, keys):
        n = len(keys)
        outs, _ = self._call(self.names["scalar_key"], n, [keys], [(n, 32)])
        return outs[0]

    def compress(self, px, py):
        n = len(px)
        outs, _ = self._call(self.names["compress"], n, [px, py], [(n, 32)])
        return outs[0]

    def decompress(self, comp):
        n = len(comp)
        outs, _ = self._call(self.names["decompress"], n, [comp], [(n, 32), (n, 32), (n,)])
        return outs

    def verify(self, r8x, r8y, s, ax, ay, msg):
        n = len(r8x)
        outs, _ = self._call(self.names["verify"], n, [r8x, r8y, s, ax, ay, msg], [(n,)])
        return outs[0]

    def verify_compressed(self, sig64, pk32, msg):
        n = len(pk32)
        outs, _ = self._call(self.names["verify_compressed"], n, [_c(sig64, 64), pk32, msg], [(n,), (n,)])
        return outs

    def verify_schnorr(self, pkx, pky, msg, rx, ry, s):
        n = len(pkx)
        outs, _ = self._call(self.names["verify_schnorr"], n, [pkx, pky, msg, rx, ry, s], [(n,), (n,)])
        return outs


class OracleC(_CBackend):
    prefix = "ora_"
    has_threads = True
    names = {k: k + "_batch" for k in ("add", "affine", "mul_scalar", "fixed_base", "public", "scalar_key", "compress",
                                       "decompress", "verify", "verify_compressed", "verify_schnorr")}

    def __init__(self, threads=None):
        self.lib = ctypes.CDLL(build_oracle_c())
        self.threads = threads or min(8, self.lib.ora_hardware_threads() or 1)

    def fr_op(self, op, a, b):
        a, b = _c(a), _c(b)
        out = np.zeros_like(a)
        self.lib.ora_fr_op(op, _sz(len(a)), _p(a), _p(b), _p(out))
        return out

    def poseidon(self, inputs):
        ins = [_c(a) for a in inputs]
        arr = (ctypes.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        out = np.zeros_like(ins[0])
        rc = self.lib.ora_poseidon_batch(len(ins), _sz(len(ins[0])), arr, _p(out), ctypes.c_int(self.threads))
        assert rc == 0
        return out

    def sign(self, keys, msgs):
        keys, msgs = _c(keys), _c(msgs)
        n = len(keys)
        rx, ry, s = (np.zeros((n, 32), dtype=np.uint8) for _ in range(3))
        self.lib.ora_sign_batch(_sz(n), _p(keys), _p(msgs), _p(rx), _p(ry), _p(s), ctypes.c_int(self.threads))
        # status like the product: 4 = "msg outside the Finite Field" (src/lib.rs:310), outputs zero
        st = np.array([4 if m > Q else 0 for m in unpack(msgs)], dtype=np.uint8)
        for a in (rx, ry, s):
            a[st != 0] = 0
        return [rx, ry, s, st]


class HostEmu(_CBackend):
    prefix = "emu_"
    has_threads = False
    names = {k: k for k in ("add", "affine", "mul_scalar", "fixed_base", "public", "scalar_key", "compress",
                            "decompress", "verify", "verify_compressed", "verify_schnorr")}

    def __init__(self):
        self.lib = ctypes.CDLL(build_hostemu())

    def fr_op(self, op, a, b):
        a, b = _c(a), _c(b)
        out = np.zeros_like(a)
        self.lib.emu_fr_op(op, _sz(len(a)), _p(a), _p(b), _p(out))
        return out

    def poseidon(self, inputs):
        ins = [_c(a) for a in inputs]
        arr = (ctypes.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        out = np.zeros_like(ins[0])
        self.lib.emu_poseidon(len(ins), _sz(len(ins[0])), arr, _p(out))
        return out

    def sign(self, keys, msgs):
        n = len(keys)
        outs, _ = self._call("sign", n, [keys, msgs], [(n, 32), (n, 32), (n, 32), (n,)])
        return outs


class Gpu:
    def __init__(self, device=0):
        import babyjubjub_rs_b200 as bjj
        self.bjj = bjj
        self.eng = bjj.Engine(device)

    def fr_op(self, op, a, b):
        return self.eng.fr_op_batch(op, a, b)

    def add(self, *a):
        return list(self.eng.add_batch(*a))

    def affine(self, *a):
        return list(self.eng.affine_batch(*a))

    def mul_scalar(self, *a):
        return list(self.eng.mul_scalar_batch(*a))

    def fixed_base(self, k):
        return list(self.eng.fixed_base_batch(k))

    def public(self, keys):
        return list(self.eng.public_batch(keys))

    def scalar_key(self, keys):
        return self.eng.scalar_key_batch(keys)

    def compress(self, px, py):
        return self.eng.compress_batch(px, py)

    def decompress(self, comp):
        return list(self.eng.decompress_batch(comp))

    def poseidon(self, inputs):
        return self.eng.poseidon_batch(inputs)

    def sign(self, keys, msgs):
        return list(self.eng.sign_batch(keys, msgs))

    def verify(self, *a):
        return self.eng.verify_batch(*a)

    def verify_compressed(self, sig64, pk32, msg):
        return list(self.eng.verify_compressed_batch(sig64, pk32, msg))

    def verify_schnorr(self, pkx, pky, msg, rx, ry, s):
        return list(self.eng.verify_schnorr_batch(pkx, pky, msg, rx, ry, s))


# ---------------------------------------------------------------------------------------------------
# case generators (deterministic)
# ---------------------------------------------------------------------------------------------------
P_KAT = (17777552123799933955779906779655732241715742912184938656739573121738514868268,
         2626589144620713026669568689430873010625803728049924121243784502389097019475)
P2_KAT = (16540640123574156134436876038791482806971768689494387082833631921987005038935,
          20819045374670962167435360035096875258406992893633759881276124905556507972311)
KEY_KAT = bytes.fromhex("0001020304050607080900010203040506070809000102030405060708090001")
MSG_KAT = int.from_bytes(bytes.fromhex("00010203040506070809"), "little")

FIELD_EDGES = [0, 1, 2, Q - 1, Q - 2, (1 << 32) - 1, 1 << 32, (1 << 32) + 1, (1 << 64) - 1, (1 << 64) + 1,
               (1 << 96) - 1, (1 << 128) + 1, (1 << 160) - 1, (1 << 192) + 1, (1 << 224) - 1, (1 << 253) + 5,
               O.Q_HALF, O.Q_HALF + 1, (1 << 256) % Q, Q - ((1 << 256) % Q)]
SCALAR_EDGES = [0, 1, 2, 3, 7, 8, 9, 15, 16, 255, 256, O.SUBORDER - 1, O.SUBORDER, O.SUBORDER + 1, O.ORDER - 1, O.ORDER,
                O.ORDER + 5, Q, (1 << 255), (1 << 256) - 1, int("8" * 64, 16), int("7" * 64, 16), int("f8" * 32, 16),
                int("80" * 32, 16), int("7f" * 32, 16)]


def special_points():
    x4 = O.modsqrt(pow(O.A, Q - 2, Q), Q)           # order-4 points (+-1/sqrt(a), 0)
    on = [P_KAT, P2_KAT, O.B8, (0, 1), (0, Q - 1), (x4, 0), (Q - x4, 0)]
    assert all(O.on_curve(p) for p in on)
    off = [(0, 0), (1, 1), ((P_KAT[0] + 1) % Q, P_KAT[1]), (5, 7), (0, 2), (1, 0), (Q - 1, Q - 1)]
    assert not any(O.on_curve(p) for p in off)
    return on, off


def random_points(rnd, n):
    return [O.mul_scalar(O.B8, rnd.randrange(1 << 251)) for _ in range(n)]


def signature_cases(rnd, n_valid=6):
    """list of [r8x, r8y, S, ax, ay, msg] covering every reference branch; msg/S clipped to 256 bits"""
    def mk(seed):
        r = random.Random(seed)
        key, msg = r.randbytes(32), r.randrange(Q)
        sig, pk = O.sign(key, msg), O.public(key)
        return [sig[0][0], sig[0][1], sig[1], pk[0], pk[1], msg]
    cases = []
    sig, pk = O.sign(KEY_KAT, MSG_KAT), O.public(KEY_KAT)
    cases.append([sig[0][0], sig[0][1], sig[1], pk[0], pk[1], MSG_KAT])          # src/lib.rs:688-738
    for s in range(n_valid):
        cases.append(mk(rnd.randrange(1 << 30)))
    base = mk(100)
    other = mk(101)
    def mod(i, v):
        c = list(base)
        c[i] = v
        return c
    cases.append(mod(2, base[2] ^ 1))                           # S bit flip
    cases.append(mod(2, base[2] ^ (1 << 200)))
    cases.append(mod(5, base[5] ^ 2))                           # msg bit flip
    cases.append(mod(2, base[2] + O.SUBORDER))                  # S + SUBORDER verifies (no range check)
    cases.append(mod(2, base[2] + 7 * O.SUBORDER))
    cases.append(mod(5, Q + 1))                                 # msg > Q -> false
    cases.append(mod(5, (1 << 256) - 1))
    cases.append(mod(0, (base[0] + 1) % Q))                     # off-curve R8
    cases.append(mod(3, (base[3] + 1) % Q))                     # off-curve A
    c = list(base); c[0], c[1] = 0, 0; cases.append(c)          # (0,0)
    c = list(base); c[3], c[4] = 0, 0; cases.append(c)
    c = list(base); c[3], c[4] = 0, 1; cases.append(c)          # identity as A
    c = list(base); c[0], c[1] = 0, 1; cases.append(c)          # identity as R8
    c = list(base); c[0], c[1] = other[0], other[1]; cases.append(c)   # foreign R8
    c = list(base); c[3], c[4] = other[3], other[4]; cases.append(c)   # foreign A
    key2 = rnd.randbytes(32)
    sig0, pk2 = O.sign(key2, 0), O.public(key2)
    cases.append([sig0[0][0], sig0[0][1], sig0[1], pk2[0], pk2[1], Q])   # msg == Q hashes as 0 -> valid
    cases.append([sig0[0][0], sig0[0][1], sig0[1], pk2[0], pk2[1], 0])
    s = rnd.randrange(O.SUBORDER)
    R = O.mul_scalar(O.B8, s)
    cases.append([R[0], R[1], s, 0, 1, 77])                     # A = identity, R8 = S*B8 -> valid
    cases.append([R[0], R[1], s, 0, Q - 1, 77])                 # A of order 2: 8*hm*A = identity -> valid
    x4 = O.modsqrt(pow(O.A, Q - 2, Q), Q)
    cases.append([R[0], R[1], s, x4, 0, 78])                    # A of order 4
    cases.append([0, 1, 0, base[3], base[4], base[5]])          # S = 0, R8 = identity
    # the equation missed by a point of order 2 / 4 only: S*B8 - R8 - 8hm*A is a non-zero torsion point.  (The
    # engine checks v*(...) == O for an ODD v -- an even v would accept these.)  hm changes with R8, so these are
    # built from the relation directly: pick R8' = S*B8 - 8hm'*A + T is not possible in closed form; instead
    # keep A = identity-like points where hm does not matter.
    for tors in ((0, Q - 1), (x4, 0), (Q - x4, 0)):
        Rt = O.proj_affine(O.proj_add(R + (1,), tors + (1,)))
        cases.append([Rt[0], Rt[1], s, 0, 1, 77])               # R8 = S*B8 + T, A = identity -> false
        cases.append([Rt[0], Rt[1], s, 0, Q - 1, 79])           # same with A of order 2
    cases.append(mod(2, base[2] + 37 * O.SUBORDER))             # S close to 2^256: still valid, recoding carry
    cases.append(mod(2, base[2] + 41 * O.SUBORDER))
    cases.append(mod(2, (base[2] + 41 * O.SUBORDER) ^ 1))
    return cases


def cases_to_arrays(cases):
    cols = list(zip(*cases))
    return [pack([v & ((1 << 256) - 1) for v in col]) for col in cols]


def oracle_verify(c):
    return int(O.verify((c[3], c[4]), ((c[0], c[1]), c[2]), c[5]))


SUBORDER_L = O.SUBORDER


def split_scalar_inputs(seed=77, n_random=400):
    """h values for verify's scalar split (csrc/split.cuh): degenerate lattices, large partial quotients, every
    operand length, and random field elements; returns (hs, ss)"""
    L = SUBORDER_L
    rnd = random.Random(seed)
    hs = [0, 1, 2, 3, L - 1, L, L + 1, 2 * L, 7 * L, (L + 1) // 2, (L - 1) // 2, (L + 1) // 2 + 1, Q - 1, 2**256 - 1,
          2**126, 2**127, 2**128 + 1, L // 3, 2 * L // 3, (1 << 200) + 1]
    hs += [pow(2, k, L) for k in (125, 126, 127, 250)]
    hs += [(L + 1) // 2 * k % L for k in (3, 5, 7)]           # 2h = k: short vectors with an even cofactor
    hs += [(L * pn // qn + d) % (1 << 256) for qn in (2, 3, 5, 7, 64, 1 << 20, (1 << 40) + 1) for pn in (1, qn - 1) for d in (0, 1)]
    hs += [L // k for k in (2, 3, 4, 1 << 31, 1 << 32, (1 << 32) + 1, 1 << 64, 1 << 125, 1 << 126, 1 << 127)]
    hs += [rnd.randrange(1 << b) for b in (8, 31, 32, 33, 64, 100, 125, 126, 127, 128, 129, 160, 192, 224, 250, 251, 252, 254, 256)]
    # chosen partial-quotient patterns of h / l (the Lehmer batches of the split run the quotient sequence on the leading
    # words: all-ones = the longest sequence, huge quotients = no step fits a leading word, mixtures = batches that end early)

    def from_cf(qs):
        num, den = 0, 1
        for q in reversed(qs):
            num, den = den, q * den + num
        return (L * num // den) % (1 << 256)
    for pattern in ([1] * 200, [2] * 120, [1, 2] * 90, [1, 1 << 15] * 12, [1 << 16] * 16, [(1 << 31) - 1] * 9, [1 << 32] * 8, [3, 1 << 33] * 6,
                    [1] * 40 + [1 << 20] + [1] * 100, [65535, 1, 65536, 2] * 8, [1] * 73 + [1 << 30, 1 << 30], [255] * 32, [65537] * 15):
        hs += [(from_cf(pattern) + d) % (1 << 256) for d in (-1, 0, 1)]
    hs += [((L >> b) << b) % (1 << 256) for b in range(8, 256, 8)] + [(L >> b) + 1 for b in range(1, 256, 5)]
    hs += [rnd.randrange(Q) for _ in range(n_random)]
    ss = [rnd.randrange(1 << 256) for _ in hs]
    ss[0], ss[1], ss[2] = 0, 2**256 - 1, L
    return hs, ss


def check_split_outputs(hs, ss, us, vs, negs, ws, max_wide=190):
    """the invariants that make the split exact: u = v*h (mod l), v odd and non-zero, w = |v|*s (mod l)"""
    L = SUBORDER_L
    wide = 0
    for h, s, u, v, ng, w in zip(hs, ss, us, vs, negs, ws):
        sv = -v if ng else v
        assert v % 2 == 1 and 0 < v < L, (h, v)
        assert (sv * h - u) % L == 0, (h, u, sv)
        assert (w - v * s) % L == 0 and w < 2 * L, (h, w)
        if max(u, v) >= 0x70000000 << 96:
            wide += 1
    assert wide <= max_wide, wide      # only crafted degenerate inputs (+ a few random ones by a bit) exceed 32 windows
