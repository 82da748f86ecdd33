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

    import numpy as np

    from common import check_split_outputs, pack, split_scalar_inputs, unpack
    hs, ss = split_scalar_inputs()
    n = len(hs)
    H, S = pack(hs), pack(ss)
    U, V, W = np.zeros_like(H), np.zeros_like(H), np.zeros_like(H)
    neg = np.zeros(n, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    hostemu.lib.emu_split(ctypes.c_size_t(n), p(H), p(S), p(U), p(V), p(neg), p(W))
    check_split_outputs(hs, ss, unpack(U), unpack(V), neg, unpack(W))
