#!/usr/bin/env python3
"""Launches every kernel of the library ONCE at bench size (after one warm-up pass), for `ncu` to capture.

    ncu --set full --clock-control none -k regex:^k_ -s <warm-up launches> -o gpurun_out/kernels \
        python babyjubjub-rs_b200/tools/profile_kernels.py
    python babyjubjub-rs_b200/tools/ncu_summarize.py gpurun_out/kernels.ncu-rep profiles/r2

Prints the number of kernel launches of the warm-up pass on stderr ("warmup_launches=N") when run with --count.
Nothing here is on the product path; it only calls the C ABI (device-pointer flavour) through the Python mirror.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import babyjubjub_rs_b200 as bjj  # noqa: E402

LOG2 = int(os.environ.get("PROFILE_LOG2_LANES", "20"))


def main():
    dev = torch.device("cuda", 0)
    eng = bjj.Engine(0)
    lib, ctx = eng.lib, eng.ctx
    n = 1 << LOG2
    g = torch.Generator(device=dev)
    g.manual_seed(1)

    def d(t):
        return ctypes.c_void_p(t.data_ptr())

    def u8(*shape):
        return torch.empty(shape, dtype=torch.uint8, device=dev)

    keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msgs = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msgs[:, 31] &= 0x1F
    k = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    k[:, 31] &= 0x3F
    r8x, r8y, s, ax, ay, cr, ca, ox, oy, dx, dy, ho = (u8(n, 32) for _ in range(12))
    st, ok = u8(n), u8(n)

    def one_pass():
        lib.bjj_sign_batch_dev(ctx, n, d(keys), d(msgs), d(r8x), d(r8y), d(s), d(st), None)
        lib.bjj_public_batch_dev(ctx, n, d(keys), d(ax), d(ay), None)
        lib.bjj_fixed_base_batch_dev(ctx, n, d(k), d(ox), d(oy), None)
        lib.bjj_compress_batch_dev(ctx, n, d(r8x), d(r8y), d(cr), None)
        lib.bjj_compress_batch_dev(ctx, n, d(ax), d(ay), d(ca), None)
        lib.bjj_decompress_batch_dev(ctx, n, d(cr), d(dx), d(dy), d(st), None)
        lib.bjj_mul_scalar_batch_dev(ctx, n, d(ax), d(ay), d(k), d(ox), d(oy), None)
        arr = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in (r8x, r8y, ax, ay, msgs)])
        lib.bjj_poseidon_batch_dev(ctx, 5, n, arr, d(ho), None)
        # verify on a 10 % corrupted copy with ~1.7 % off-curve lanes (the bench mix, approximately)
        bad = ax.clone()
        bad[::59, 0] ^= 1
        m2 = msgs.clone()
        m2[::12, 1] ^= 4
        lib.bjj_verify_batch_dev(ctx, n, d(r8x), d(r8y), d(s), d(bad), d(ay), d(m2), d(ok), None)
        sig64 = torch.cat([cr, s], dim=1).contiguous()
        lib.bjj_verify_compressed_batch_dev(ctx, n, d(sig64), d(ca), d(m2), d(ok), d(st), None)
        lib.bjj_add_batch_dev(ctx, n, d(ax), d(ay), d(r8x), d(r8x), d(r8y), d(ay), d(ox), d(oy), d(dx), None)
        eng.sync()

    l0 = eng.kernel_launches
    one_pass()
    warm = eng.kernel_launches - l0
    print("warmup_launches=%d" % warm, file=sys.stderr, flush=True)
    if "--count" in sys.argv:
        return
    one_pass()


if __name__ == "__main__":
    main()
