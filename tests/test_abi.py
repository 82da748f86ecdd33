"""The C-ABI library loads without a GPU and exports every symbol include/bjj_cuda.h declares."""
import ctypes
import os
import re

import pytest

from common import ROOT

HEADER = os.path.join(ROOT, "include", "bjj_cuda.h")
LIB = os.path.join(ROOT, "babyjubjub-rs_b200", "libbjj_cuda.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bjj_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_four_north_star_entry_points():
    syms = declared_symbols()
    for name in ("bjj_mul_scalar_batch", "bjj_public_batch", "bjj_decompress_batch", "bjj_verify_batch",
                 "bjj_add_batch", "bjj_verify_compressed_batch", "bjj_poseidon_batch", "bjj_fixed_base_batch"):
        assert name in syms and name + "_dev" in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        pytest.fail("libbjj_cuda.so not built: run __graft_entry__.build()")
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    import babyjubjub_rs_b200 as bjj
    from babyjubjub_rs_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()
    assert bjj.Q % 2 == 1


def test_no_cpu_fallback_without_device():
    """no compute without a GPU: creating an Engine must fail loudly on a box with no CUDA device"""
    import babyjubjub_rs_b200 as bjj
    lib = bjj._lib.load()
    if lib.bjj_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        bjj.Engine(0)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "babyjubjub-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
                assert "hostemu" not in text or f in ("fr.cuh",), f


def test_rust_ffi_declares_every_host_flavour_symbol():
    """rust/src/ffi.rs cannot be compiled here (no toolchain), but it can be read: every host-pointer entry point of
    the header must have an `extern "C"` declaration there, and nothing that the header lacks."""
    text = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (bjj_[a-z0-9_]+)\s*\(", text))
    header = set(declared_symbols())
    host = {s for s in header if not s.endswith("_dev")}
    assert not (host - declared), sorted(host - declared)
    assert not (declared - header), sorted(declared - header)
    lib_rs = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    assert "unsafe impl Sync for Engine" not in lib_rs           # the C context is one-thread-at-a-time
    assert "Mutex<Engine>" in lib_rs
