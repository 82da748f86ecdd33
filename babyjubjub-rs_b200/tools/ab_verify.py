#!/usr/bin/env python3
"""A/B harness for the verify kernels: several builds of csrc/k_verify.cu, timed on the SAME GPU.

The GPUs of the pool differ by up to 9 % on k_verify_ec, so two variants can only be
compared inside one gpurun call.  This tool

  build   compiles k_verify.cu once per variant (a variant = a set of -D macros) and links each against the
          objects of the normal build into  ab_variants/<name>/libbjj_cuda.so  (in-tree, so it travels to the
          GPU box; the directory is git-ignored through *.so / *.o),
  run     (on the GPU box) swaps each variant in turn for babyjubjub-rs_b200/libbjj_cuda.so, runs the verify
          parity tests and prints BJJ_PHASE_TIMING lines of bench.py for it, then restores the original.

    python babyjubjub-rs_b200/tools/ab_verify.py build base: sqr:BJJ_DEDICATED_SQR=1
    gpurun --timeout 600 -- 'python babyjubjub-rs_b200/tools/ab_verify.py run base sqr'

AB_SKIP_TESTS=1 skips the parity tests, AB_STEPS=n sets the timed steps.  Nothing here is on the product path.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
OUT = os.path.join(ROOT, "ab_variants")
LIB = os.path.join(PKG, "libbjj_cuda.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "20091", "-Xcompiler", "-fPIC"]
OTHER_UNITS = ["bjj_cuda", "bjj_multi", "k_mulscalar", "k_sign", "k_poseidon"]


def build(specs):
    sys.path.insert(0, PKG)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bjj_build", os.path.join(PKG, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build(tools=False)                                 # the normal objects (and the generated constants)
    nvcc = os.environ.get("NVCC", "nvcc")
    for s in specs:
        name, _, macros = s.partition(":")
        d = os.path.join(OUT, name)
        os.makedirs(d, exist_ok=True)
        defs = ["-D" + m for m in macros.split(",") if m]
        obj = os.path.join(d, "k_verify.o")
        if name.startswith("all_"):                      # a variant named all_* rebuilds EVERY unit with the macros
            from concurrent.futures import ThreadPoolExecutor
            units = ["k_verify"] + OTHER_UNITS
            with ThreadPoolExecutor(max_workers=len(units)) as ex:
                list(ex.map(lambda u: subprocess.check_call([nvcc] + FLAGS + ["-diag-suppress", "177"] + defs + ["-c", os.path.join(CSRC, u + ".cu"), "-o", os.path.join(d, u + ".o")], cwd=CSRC), units))
            subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", os.path.join(d, "libbjj_cuda.so")] + [os.path.join(d, u + ".o") for u in units])
        else:
            subprocess.check_call([nvcc] + FLAGS + defs + ["-c", os.path.join(CSRC, "k_verify.cu"), "-o", obj], cwd=CSRC)
            subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", os.path.join(d, "libbjj_cuda.so"), obj] +
                                  [os.path.join(OBJ, u + ".o") for u in OTHER_UNITS])
        res = subprocess.run(["cuobjdump", "-res-usage", obj], stdout=subprocess.PIPE, text=True).stdout.splitlines()
        regs = [ln.strip().split()[0] for prev, ln in zip(res, res[1:]) if "k_verify_ec" in prev]
        print("built %-12s %-40s k_verify_ec %s" % (name, " ".join(defs) or "(no macros)", regs[0] if regs else "?"))


def run(names, steps=int(os.environ.get("AB_STEPS", "3"))):
    keep = LIB + ".ab_keep"
    shutil.copy2(LIB, keep)
    try:
        for name in names:
            shutil.copy2(os.path.join(OUT, name, "libbjj_cuda.so"), LIB)
            print("== variant %s" % name, flush=True)
            if not os.environ.get("AB_SKIP_TESTS"):
                t = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q", "-k", "verify or schnorr"],
                                   cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                print(t.stdout.strip().splitlines()[-1], flush=True)
            env = dict(os.environ, BJJ_PHASE_TIMING="1")
            p = subprocess.run([sys.executable, "bench.py", "--steps", str(steps), "--warmup", "3", "--no-secondary", "--cpu-seconds", "1"],
                               cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            lines = [ln for ln in p.stderr.splitlines() if "lanes=2097152" in ln]
            for ln in lines[-2:]:
                print(ln, flush=True)
    finally:
        shutil.copy2(keep, LIB)
        os.remove(keep)


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("build", "run"):
        sys.exit(__doc__)
    (build if sys.argv[1] == "build" else run)(sys.argv[2:])
