"""host/bjj.hpp (the C++ mirror of the reference API) compiles against include/bjj_cuda.h and links with
libbjj_cuda.so without a GPU; on the GPU box the program replays the reference's unit tests."""
import os
import subprocess

import pytest

from common import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_host.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host")
PKG = os.path.join(ROOT, "babyjubjub-rs_b200")


def build():
    deps = [SRC, os.path.join(ROOT, "host", "bjj.hpp"), os.path.join(ROOT, "include", "bjj_cuda.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", EXE, SRC, "-L" + PKG, "-lbjj_cuda",
                               "-Wl,-rpath," + PKG])
    return EXE


def test_cpp_host_mirror_builds_and_links():
    assert os.path.exists(build())


@pytest.mark.gpu
def test_cpp_host_mirror_replays_reference_tests():
    out = subprocess.run([build()], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror ok" in out.stdout
