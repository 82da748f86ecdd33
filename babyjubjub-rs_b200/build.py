"""Builds libbjj_cuda.so for sm_100a with nvcc (in-tree, so the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "generated", "bjj_consts.inc")
LIB = os.path.join(HERE, "libbjj_cuda.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "20091"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(HERE, "..", "include", "bjj_cuda.h"))
    return out


def generate_constants(force=False):
    gen_py = os.path.join(HERE, "tools", "gen_constants.py")
    if force or not _newer(GEN, [gen_py]):
        subprocess.check_call([sys.executable, gen_py, GEN])
    return GEN


def build(force=False, verbose=False):
    generate_constants()
    srcs = sources() + [GEN]
    if not force and _newer(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB + ".tmp", os.path.join(CSRC, "bjj_cuda.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
