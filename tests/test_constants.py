"""The product's build-time constant generator (babyjubjub-rs_b200/tools/gen_constants.py) and the
oracle's (oracle/poseidon_constants.py) are independent implementations of the Grain LFSR; they must
agree, and the emitted Montgomery constants must decode to the reference's values (src/lib.rs:28-58)."""
import importlib.util
import os
import re

from common import O, Q, ROOT


def _gen():
    spec = importlib.util.spec_from_file_location("gen_constants", os.path.join(ROOT, "babyjubjub-rs_b200", "tools", "gen_constants.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_poseidon_tables_agree():
    from oracle import poseidon_constants as pc
    g = _gen()
    for t in (2, 3, 6, 7):
        rc, mds = g.poseidon_tables(t)
        C, M = pc.constants(t)
        assert rc == C and mds == M


def test_emitted_constants_decode():
    g = _gen()
    inc = os.path.join(ROOT, "babyjubjub-rs_b200", "csrc", "generated", "bjj_consts.inc")
    if not os.path.exists(inc):
        g.main()
    text = open(inc).read()
    rinv = pow(1 << 256, -1, Q)

    def const(name):
        m = re.search(r"BJJ_%s\[8\] = \{([^}]*)\}" % name, text)
        w = [int(x.strip().rstrip("u"), 16) for x in m.group(1).split(",")]
        return sum(v << (32 * i) for i, v in enumerate(w))
    assert const("Q") == Q and const("TWOQ") == 2 * Q and const("QHALF") == Q >> 1
    assert const("ONE_M") * rinv % Q == 1
    assert const("A_M") * rinv % Q == O.A and const("D_M") * rinv % Q == O.D
    assert (const("B8X_M") * rinv % Q, const("B8Y_M") * rinv % Q) == O.B8
    s = const("SQRT_NEG_A_M") * rinv % Q
    assert (s * s + O.A) % Q == 0
    assert const("INV_SQRT_NEG_A_M") * rinv % Q * s % Q == 1
    assert const("SUBORDER") == O.SUBORDER and const("ORDER") == O.ORDER
    assert const("TWO_DP_M") * rinv % Q == (-2 * O.D * pow(O.A, -1, Q)) % Q


def test_optimized_poseidon_schedule_equals_dense():
    """the sparse partial-round schedule baked into the GPU tables must hash like the dense oracle"""
    import random
    g = _gen()
    rnd = random.Random(9)
    for t in range(2, 8):
        for ins in ([0] * (t - 1), list(range(1, t)), [rnd.randrange(Q) for _ in range(t - 1)]):
            assert g.poseidon_optimized_eval(ins) == O.poseidon(ins), t


def test_grouped_poseidon_schedule_equals_dense():
    """the grouped partial rounds (three rounds share the reductions of lanes 1..t-1; gen_constants.py::poseidon_groups),
    evaluated from the table the device reads, must hash like the dense oracle -- for every group size"""
    import random
    g = _gen()
    rnd = random.Random(10)
    for t in range(2, 8):
        for ins in ([0] * (t - 1), [Q - 1] * (t - 1), [rnd.randrange(Q) for _ in range(t - 1)]):
            want = O.poseidon(ins)
            for group in (1, 2, 3, 4):
                assert g.poseidon_grouped_eval(ins, group) == want, (t, group)
    # table geometry the device code assumes
    for t in range(2, 8):
        for K, row in g.poseidon_groups(t):
            assert len(row) == 2 * K * t + K * (K - 1) // 2


def test_fixed_exponent_schedules():
    """the sliding-window schedules of fr_pow_sched (Fermat inversion, square-root exponent) compute the plain powers, and
    bench.py counts their multiplications and squarings as generated"""
    import random
    import bench
    g = _gen()
    rnd = random.Random(11)
    t = (Q - 1) >> 28
    for name, e, counted in (("QM2", Q - 2, bench.FERMAT), ("TM1H", (t - 1) // 2, bench.SQRT_POW)):
        for a in (0, 1, 2, Q - 1, rnd.randrange(Q), rnd.randrange(Q)):
            assert g.pow_schedule_eval(a, e) == pow(a, e, Q), name
        steps, tail = g.pow_schedule(e)
        assert all(0 <= idx < 8 and 0 < n < 256 for n, idx in steps)
        assert counted == (len(steps) - 1 + 7, sum(n for n, _ in steps[1:]) + tail + 1), name
