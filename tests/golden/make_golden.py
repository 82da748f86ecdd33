#!/usr/bin/env python3
"""Regenerates tests/golden/edge_vectors.json from the pure-Python oracle (oracle/bjj_oracle.py), whose
own correctness is pinned on the reference's known-answer tests (reference_kats.json).  The reference is
a Rust crate and cannot be run in this image, so these vectors cover what the reference's tests do NOT
pin ("gaps"): negative verify cases, every Err branch of decompress_point,
S >= SUBORDER, msg == Q, off-curve / low-order / (0,0) points, zero scalars.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from common import O, Q, SCALAR_EDGES, signature_cases, special_points, oracle_verify  # noqa: E402

STATUS = {O.ERR_Y_RANGE: 1, O.ERR_NO_INV: 2, O.ERR_NOT_SQUARE: 3}


def main():
    rnd = random.Random(0xB200)
    on, off = special_points()
    out = {"_made_by": "tests/golden/make_golden.py (pure-Python oracle, seed 0xB200)"}
    ms = []
    for p in on + off:
        for k in SCALAR_EDGES[:6] + [rnd.randrange(1 << 256)]:
            r = O.mul_scalar(p, k)
            ms.append({"px": str(p[0]), "py": str(p[1]), "k": str(k), "rx": str(r[0]), "ry": str(r[1]), "on_curve": O.on_curve(p)})
    out["mul_scalar"] = ms
    dc = []
    blobs = [O.compress(p) for p in on] + [int(v).to_bytes(32, "little") for v in
             (0, 1, Q - 1, Q, Q + 1, 2, (1 << 255) | 1, (1 << 255) | (Q - 1), (1 << 255), (1 << 256) - 1, 5, (1 << 255) | 5)]
    blobs += [rnd.randbytes(32) for _ in range(24)]
    for b in blobs:
        try:
            p, st = O.decompress_point(b), 0
        except ValueError as e:
            p, st = (0, 0), STATUS[str(e)]
        dc.append({"in": b.hex(), "x": str(p[0]), "y": str(p[1]), "status": st})
    out["decompress"] = dc
    vf = []
    for c in signature_cases(random.Random(7), 3):
        vf.append({"r8x": str(c[0]), "r8y": str(c[1]), "s": str(c[2] & ((1 << 256) - 1)), "ax": str(c[3]), "ay": str(c[4]),
                   "msg": str(c[5] & ((1 << 256) - 1)), "ok": oracle_verify(c)})
    out["verify"] = vf
    ps = []
    for nin in range(1, 7):
        for ins in (list(range(1, nin + 1)), [0] * nin, [Q - 1] * nin, [rnd.randrange(Q) for _ in range(nin)]):
            ps.append({"in": [str(v) for v in ins], "out": str(O.poseidon(ins))})
    out["poseidon"] = ps
    with open(os.path.join(HERE, "edge_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
        f.write("\n")
    print("wrote edge_vectors.json: %d mul_scalar, %d decompress, %d verify, %d poseidon" % (len(ms), len(dc), len(vf), len(ps)))


if __name__ == "__main__":
    main()
