"""CPU check of the DEVICE code's logic: the .cuh headers compiled for the host (PTX chains replaced by
their C emulation) against the oracle.  This is not a product path -- see tests/hostemu/hostemu.cpp."""
import parity


def test_fr(hostemu, oracle_c):
    parity.check_fr(hostemu, oracle_c, 3000)


def test_add_affine(hostemu, oracle_c):
    parity.check_add(hostemu, oracle_c, 64)


def test_mul_scalar(hostemu, oracle_c):
    parity.check_mul_scalar(hostemu, oracle_c, 64)


def test_fixed_base_public(hostemu, oracle_c):
    parity.check_fixed_base(hostemu, oracle_c, 64)


def test_compress_decompress(hostemu, oracle_c):
    parity.check_compress_decompress(hostemu, oracle_c, 300)


def test_poseidon(hostemu, oracle_c):
    parity.check_poseidon(hostemu, oracle_c, 8)


def test_verify(hostemu, oracle_c):
    parity.check_verify(hostemu, oracle_c, 6)


def test_sign(hostemu, oracle_c):
    parity.check_sign(hostemu, oracle_c, 6)


def test_schnorr(hostemu, oracle_c):
    parity.check_schnorr(hostemu, oracle_c, 3)


def test_verify_without_split(hostemu, oracle_c):
    """full-width scalars (u = hm, v = 1): the 64/65-window path of the same Straus pass"""
    hostemu.lib.emu_set_split(0)
    try:
        parity.check_verify(hostemu, oracle_c, 3)
    finally:
        hostemu.lib.emu_set_split(1)


def test_split_scalars(hostemu):
    """verify's half-size scalars (csrc/split.cuh): u = v*h (mod l) exactly, v odd and non-zero, w = |v|*S (mod l);
    generic inputs give 127-bit scalars, degenerate lattices only longer ones."""
    import ctypes
    import random

    import numpy as np

    from common import Q, pack, unpack
    L = 21888242871839275222246405745257275088614511777268538073601725287587578984328 >> 3
    rnd = random.Random(77)
    hs = [0, 1, 2, 3, L - 1, L, L + 1, 2 * L, 7 * L, (L + 1) // 2, (L - 1) // 2, (L + 1) // 2 + 1, Q - 1, 2**256 - 1,
          2**126, 2**127, 2**128 + 1, L // 3, 2 * L // 3, (1 << 200) + 1]
    hs += [pow(2, k, L) for k in (125, 126, 127, 250)]
    hs += [(L + 1) // 2 * k % L for k in (3, 5, 7)]           # 2h = k: short vectors with an even cofactor
    hs += [rnd.randrange(Q) for _ in range(400)]
    ss = [rnd.randrange(1 << 256) for _ in hs]
    ss[0], ss[1], ss[2] = 0, 2**256 - 1, L
    n = len(hs)
    H, S = pack(hs), pack(ss)
    U, V, W = np.zeros_like(H), np.zeros_like(H), np.zeros_like(H)
    neg = np.zeros(n, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    hostemu.lib.emu_split(ctypes.c_size_t(n), p(H), p(S), p(U), p(V), p(neg), p(W))
    wide = 0
    for h, s, u, v, w, ng in zip(hs, ss, unpack(U), unpack(V), unpack(W), neg):
        sv = -v if ng else v
        assert v % 2 == 1 and 0 < v < L, (h, v)
        assert (sv * h - u) % L == 0, (h, u, sv)
        assert (w - v * s) % L == 0 and w < 2 * L, (h, w)
        if max(u, v) >= 0x70000000 << 96:
            wide += 1
    # only the crafted degenerate inputs may need more than 32 windows (+ a few random ones by a bit)
    assert wide <= 40, wide
