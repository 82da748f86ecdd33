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
