"""GPU parity through the C ABI (libbjj_cuda.so) against the oracle.  Bit-exact: this is integer work."""
import os
import numpy as np
import pytest

import parity
from common import O, Q, pack, unpack

pytestmark = pytest.mark.gpu


def test_fr(gpu, oracle_c):
    parity.check_fr(gpu, oracle_c, 1 << 16)


def test_add_affine(gpu, oracle_c):
    parity.check_add(gpu, oracle_c, 1024)


def test_mul_scalar(gpu, oracle_c):
    parity.check_mul_scalar(gpu, oracle_c, 4096)


def test_fixed_base_public(gpu, oracle_c):
    parity.check_fixed_base(gpu, oracle_c, 4096)


def test_compress_decompress(gpu, oracle_c):
    parity.check_compress_decompress(gpu, oracle_c, 8192)


def test_poseidon(gpu, oracle_c):
    parity.check_poseidon(gpu, oracle_c, 512)


def test_verify(gpu, oracle_c):
    parity.check_verify(gpu, oracle_c, 64)


def test_sign(gpu, oracle_c):
    parity.check_sign(gpu, oracle_c, 256)


def test_reference_api_mirror(gpu):
    """the reference's own unit tests, re-expressed on the host mirror (src/lib.rs:420-738)"""
    bjj = gpu.bjj
    from common import KEY_KAT, MSG_KAT, P2_KAT, P_KAT
    p = bjj.Point(*P_KAT)
    r = p.projective().add(p.projective()).affine()                                   # test_add_same_point
    assert (r.x, r.y) == (6890855772600357754907169075114257697580319025794532037257385534741338397365,
                          4338620300185947561074059802482547481416142213883829469920100239455078257889)
    r = p.projective().add(bjj.Point(*P2_KAT).projective()).affine()                  # test_add_different_points
    assert (r.x, r.y) == (7916061937171219682591368294088513039687205273691143098332585753343424131937,
                          14035240266687799601661095864649209771790948434046947201833777492504781204499)
    m3 = p.mul_scalar(3)                                                              # test_mul_scalar
    a3 = p.projective().add(p.projective()).add(p.projective()).affine()
    assert m3.equals(a3)
    assert m3.x == 19372461775513343691590086534037741906533799473648040012278229434133483800898
    c = p.compress()                                                                  # test_point_compress_decompress
    assert c.hex() == "53b81ed5bffe9545b54016234682e7b2f699bd42a5e9eae27ff4051bc698ce85"
    assert bjj.decompress_point(c).equals(p)
    with pytest.raises(ValueError, match="not a mod p square"):
        bjj.decompress_point((1).to_bytes(32, "little"))
    with pytest.raises(ValueError, match="y outside the Finite Field over R"):
        bjj.decompress_point(Q.to_bytes(32, "little"))
    sk = bjj.PrivateKey.import_(KEY_KAT)                                              # test_circomlib_testvector
    assert sk.scalar_key() == 6466070937662820620902051049739362987537906109895538826186780010858059362905
    pk = sk.public()
    assert pk.x == 0x1d5ac1f31407018b7d413a4f52c8f74463b30e6ac2238220ad8b254de4eaa3a2
    osig = O.sign(KEY_KAT, MSG_KAT)
    sig = sk.sign(MSG_KAT)
    assert (sig.r_b8.x, sig.r_b8.y, sig.s) == (osig[0][0], osig[0][1], osig[1])
    with pytest.raises(ValueError, match="msg outside the Finite Field"):
        sk.sign(Q + 1)
    assert bjj.verify(pk, sig, MSG_KAT) is True
    assert bjj.verify(pk, sig, MSG_KAT + 1) is False
    sig2 = bjj.decompress_signature(sig.compress())                                   # test_signature_compress_decompress
    assert sig2.r_b8.equals(sig.r_b8) and sig2.s == sig.s
    assert bjj.verify_batch([pk, pk], [sig, sig2], [MSG_KAT, MSG_KAT + 2]) == [True, False]
    pts = bjj.mul_scalar_batch([p, p], [3, 0])
    assert pts[0].equals(m3) and (pts[1].x, pts[1].y) == (0, 1)
    assert bjj.public_batch([sk])[0].equals(pk)
    d = bjj.decompress_batch([c, (1).to_bytes(32, "little")])
    assert d[0].equals(p) and isinstance(d[1], ValueError)
    k = (1 << 1023) + 0xDEADBEEF                                                      # test_schnorr_signature
    r_s, s_s = sk.sign_schnorr(MSG_KAT, k=k)
    o_r, o_s = O.sign_schnorr(KEY_KAT, MSG_KAT, k)
    assert (r_s.x, r_s.y, s_s) == (o_r[0], o_r[1], o_s)
    assert bjj.verify_schnorr(pk, MSG_KAT, r_s, s_s) is True
    assert bjj.verify_schnorr(pk, MSG_KAT + 1, r_s, s_s) is False
    sk2 = bjj.new_key()
    r2, s2 = sk2.sign_schnorr(12345)
    assert bjj.verify_schnorr(sk2.public(), 12345, r2, s2) is True


def test_noncanonical_inputs_are_flagged(gpu):
    bjj = gpu.bjj
    a = pack([Q + 5, 3])
    with pytest.raises(bjj.BjjError) as ei:
        gpu.eng.fr_op_batch(bjj.FR_MUL, a, a)
    assert ei.value.code == bjj.ERR_NONCANONICAL
    # the flag is cleared by the failing call; the next call is clean
    out = gpu.eng.fr_op_batch(bjj.FR_MUL, pack([2, 3]), pack([5, 7]))
    assert unpack(out) == [10, 21]


def test_empty_and_ragged_batches(gpu, oracle_c):
    e = np.zeros((0, 32), dtype=np.uint8)
    assert gpu.eng.verify_batch(e, e, e, e, e, e).shape == (0,)
    assert gpu.eng.fixed_base_batch(e)[0].shape == (0, 32)
    for n in (1, 31, 33, 127, 129, 1000):
        k = pack([(i * 0x9E3779B97F4A7C15 + 12345) % (1 << 256) for i in range(n)])
        got, exp = gpu.fixed_base(k), oracle_c.fixed_base(k)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])


def test_large_batch_properties(gpu, oracle_c):
    """size-independent checks at a chunk-crossing size: linearity of fixed-base, compress/decompress
    round trip, and a sampled oracle comparison"""
    n = (1 << 20) + 777            # crosses the 2^20-lane pipeline chunk
    rng = np.random.default_rng(42)
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F               # < 2^253
    rx, ry = gpu.eng.fixed_base_batch(k)
    comp = gpu.eng.compress_batch(rx, ry)
    dx, dy, st = gpu.eng.decompress_batch(comp)
    assert not st.any() and np.array_equal(dx, rx) and np.array_equal(dy, ry)
    idx = rng.choice(n, size=512, replace=False)
    ex, ey = oracle_c.fixed_base(k[idx])
    assert np.array_equal(rx[idx], ex) and np.array_equal(ry[idx], ey)
    # (k * B8) via the variable-base path must agree with the comb
    b8x, b8y = pack([O.B8[0]] * 4096), pack([O.B8[1]] * 4096)
    vx, vy = gpu.eng.mul_scalar_batch(b8x, b8y, k[:4096])
    assert np.array_equal(vx, rx[:4096]) and np.array_equal(vy, ry[:4096])


def test_multi_device_sharding(gpu, oracle_c):
    """verify_batch sharded over every visible device equals the oracle; on a multi-GPU box every device takes part"""
    import random
    from common import cases_to_arrays, signature_cases
    bjj = gpu.bjj
    ndev = bjj._lib.load().bjj_device_count()
    mg = bjj.multi_gpu()
    assert len(mg.engines) == ndev
    if ndev > 1:
        assert len(mg.engines) > 1
    arrs = cases_to_arrays(signature_cases(random.Random(21), 5))
    ok = bjj.verify_batch_multi(mg, *arrs)
    assert np.array_equal(ok, oracle_c.verify(*arrs))


def test_multi_engine_single_caller(gpu, oracle_c):
    """bjj_multi_*: ONE host batch cut into contiguous shards over every device by the library's own threads
    (config 4 as written).  Results must equal the single-device calls and the oracle, with pageable inputs, with
    page-locking on, and with lane counts that do not divide by the device count."""
    import random
    from common import cases_to_arrays, signature_cases
    bjj = gpu.bjj
    ndev = bjj._lib.load().bjj_device_count()
    me = bjj.MultiEngine()
    assert me.devices == ndev
    arrs = cases_to_arrays(signature_cases(random.Random(33), 7))
    exp = oracle_c.verify(*arrs)
    assert np.array_equal(me.verify_batch(*arrs), exp)
    me.set_host_register(True)
    assert np.array_equal(me.verify_batch(*arrs), exp)
    me.set_host_register(False)
    # ragged sizes: fewer lanes than devices, and a prime lane count
    for n in (1, 3, 1009):
        k = pack([(i * 0x9E3779B97F4A7C15 + 777) % (1 << 253) for i in range(n)])
        gx, gy = me.fixed_base_batch(k)
        ex, ey = oracle_c.fixed_base(k)
        assert np.array_equal(gx, ex) and np.array_equal(gy, ey)
        px, py = me.public_batch(k)
        sx, sy = gpu.eng.public_batch(k)
        assert np.array_equal(px, sx) and np.array_equal(py, sy)
        mx, my = me.mul_scalar_batch(gx, gy, k)
        qx, qy = oracle_c.mul_scalar(gx, gy, k)
        assert np.array_equal(mx, qx) and np.array_equal(my, qy)
        comp = gpu.eng.compress_batch(gx, gy)
        dx, dy, st = me.decompress_batch(comp)
        assert not st.any() and np.array_equal(dx, gx) and np.array_equal(dy, gy)
    launches = me.kernel_launches
    assert launches > 0
    me.close()


def test_device_pointer_flavour_across_subbatches(gpu, oracle_c):
    """the _dev entry points with more lanes than one internal sub-batch (2^21): sign -> public -> verify on
    device-resident data, using only the library's own allocation / copy helpers (no torch)"""
    import ctypes
    eng = gpu.eng
    lib, ctx = eng.lib, eng.ctx
    n = (1 << 21) + 4321
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    msgs[:, 31] &= 0x1F
    msgs[5] = 0xFF                                  # msg > Q: sign status 4, verify false

    def dev(nbytes):
        p = lib.bjj_dev_alloc(ctx, nbytes)
        assert p
        return ctypes.c_void_p(p)
    d = {k: dev(32 * n) for k in ("keys", "msgs", "r8x", "r8y", "s", "ax", "ay")}
    d_st, d_ok = dev(n), dev(n)
    try:
        for name, arr in (("keys", keys), ("msgs", msgs)):
            assert lib.bjj_memcpy_h2d(ctx, d[name], arr.ctypes.data_as(ctypes.c_void_p), 32 * n) == 0
        assert lib.bjj_sign_batch_dev(ctx, n, d["keys"], d["msgs"], d["r8x"], d["r8y"], d["s"], d_st, None) == 0
        assert lib.bjj_public_batch_dev(ctx, n, d["keys"], d["ax"], d["ay"], None) == 0
        assert lib.bjj_verify_batch_dev(ctx, n, d["r8x"], d["r8y"], d["s"], d["ax"], d["ay"], d["msgs"], d_ok, None) == 0
        ok = np.empty(n, dtype=np.uint8)
        st = np.empty(n, dtype=np.uint8)
        ax = np.empty((n, 32), dtype=np.uint8)
        for dst, src, nb in ((ok, d_ok, n), (st, d_st, n), (ax, d["ax"], 32 * n)):
            assert lib.bjj_memcpy_d2h(ctx, dst.ctypes.data_as(ctypes.c_void_p), src, nb) == 0
        eng.sync()
        assert st[5] == 4 and ok[5] == 0
        mask = np.ones(n, dtype=bool)
        mask[5] = False
        assert st[mask].max() == 0 and ok[mask].all()          # every signature made on the device verifies
        idx = np.concatenate([np.arange(8), np.arange((1 << 21) - 4, (1 << 21) + 4), np.arange(n - 8, n)])
        ex, _ = oracle_c.public(keys[idx])
        assert np.array_equal(ax[idx], ex)
        # the same signatures through the HOST flavour: ramped chunks (2^18, 2^19, 2^20, ...) on two copy streams,
        # one compute stream, exact lanes (off-curve points below) joined at each chunk's copy-out
        n2 = (1 << 20) + (1 << 18) + 333
        host = {}
        for name in ("r8x", "r8y", "s", "ay"):
            host[name] = np.empty((n2, 32), dtype=np.uint8)
            assert lib.bjj_memcpy_d2h(ctx, host[name].ctypes.data_as(ctypes.c_void_p), d[name], 32 * n2) == 0
        eng.sync()
        hax, hmsg = ax[:n2].copy(), msgs[:n2].copy()
        bad = np.arange(100, n2, 4099)                   # spread over every chunk
        for j, i in enumerate(bad):
            if j % 3 == 0:
                hax[i, 0] ^= 1                           # A off the curve (or another point): exact lane, rejects
            elif j % 3 == 1:
                host["r8x"][i, 0] ^= 1                   # R8 off the curve
            else:
                host["s"][i, 0] ^= 1                     # wrong S
        ok2 = eng.verify_batch(host["r8x"], host["r8y"], host["s"], hax, host["ay"], hmsg)
        exp2 = np.ones(n2, dtype=np.uint8)
        exp2[5] = 0
        exp2[bad] = 0
        assert np.array_equal(ok2, exp2)
        chk = np.concatenate([bad[:64], bad[-64:], np.arange(16)])
        assert np.array_equal(ok2[chk], oracle_c.verify(host["r8x"][chk], host["r8y"][chk], host["s"][chk], hax[chk],
                                                        host["ay"][chk], hmsg[chk]))
        # The call above took PAGEABLE numpy arrays: the library staged them through its page-locked mirror.  The same
        # batch from page-locked arrays (bjj_host_alloc: copied by the DMA engine directly) and from a mix of both must
        # give the same bytes.
        cols = [host["r8x"], host["r8y"], host["s"], hax, host["ay"], hmsg]
        pinned_ptrs = []

        def pinned_like(a):
            ptr = lib.bjj_host_alloc(a.nbytes)
            assert ptr
            pinned_ptrs.append(ptr)
            v = np.frombuffer((ctypes.c_uint8 * a.nbytes).from_address(ptr), dtype=np.uint8).reshape(a.shape)
            v[...] = a
            return v
        try:
            pcols = [pinned_like(c) for c in cols]
            pok = pinned_like(np.zeros(n2, dtype=np.uint8))

            def call(cs, okbuf):
                okbuf[...] = 7
                rc = lib.bjj_verify_batch(ctx, n2, *[c.ctypes.data_as(ctypes.c_void_p) for c in cs], okbuf.ctypes.data_as(ctypes.c_void_p))
                assert rc == 0
                return okbuf.copy()
            assert np.array_equal(call(pcols, pok), exp2), "all arrays page-locked"
            mixed = [pcols[0], cols[1], pcols[2], cols[3], cols[4], pcols[5]]
            assert np.array_equal(call(mixed, np.empty(n2, dtype=np.uint8)), exp2), "mixed, pageable result"
            assert np.array_equal(call(cols, pok), exp2), "pageable inputs, page-locked result"
        finally:
            for ptr in pinned_ptrs:
                lib.bjj_host_free(ctypes.c_void_p(ptr))
    finally:
        for p in list(d.values()) + [d_st, d_ok]:
            lib.bjj_dev_free(ctx, p)


def test_host_flavour_chunk_boundaries(gpu, oracle_c):
    """bjj_verify_batch / bjj_public_batch from PAGEABLE arrays at sizes around every chunk boundary of the host pipeline
    (first chunk 2^18, then up to 2^21 for verify and the doubling ramp for public; short remainders are folded into the
    last chunk): every lane answered, none twice, results as constructed"""
    import ctypes
    eng = gpu.eng
    lib, ctx = eng.lib, eng.ctx
    nmax = (1 << 21) + (1 << 18) + 5
    rng = np.random.default_rng(11)
    keys = rng.integers(0, 256, size=(nmax, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, size=(nmax, 32), dtype=np.uint8)
    msgs[:, 31] &= 0x1F

    def dev(nbytes):
        p = lib.bjj_dev_alloc(ctx, nbytes)
        assert p
        return ctypes.c_void_p(p)
    d = {k: dev(32 * nmax) for k in ("keys", "msgs", "r8x", "r8y", "s", "ax", "ay")}
    d_st = dev(nmax)
    try:
        for name, arr in (("keys", keys), ("msgs", msgs)):
            assert lib.bjj_memcpy_h2d(ctx, d[name], arr.ctypes.data_as(ctypes.c_void_p), 32 * nmax) == 0
        assert lib.bjj_sign_batch_dev(ctx, nmax, d["keys"], d["msgs"], d["r8x"], d["r8y"], d["s"], d_st, None) == 0
        assert lib.bjj_public_batch_dev(ctx, nmax, d["keys"], d["ax"], d["ay"], None) == 0
        host = {}
        for name in ("r8x", "r8y", "s", "ax", "ay"):
            host[name] = np.empty((nmax, 32), dtype=np.uint8)
            assert lib.bjj_memcpy_d2h(ctx, host[name].ctypes.data_as(ctypes.c_void_p), d[name], 32 * nmax) == 0
        eng.sync()
    finally:
        for p in list(d.values()) + [d_st]:
            lib.bjj_dev_free(ctx, p)
    bad = np.arange(7, nmax, 65521)
    host["s"][bad, 1] ^= 2
    sizes = [(1 << 18) - 1, (1 << 18) + 1, (1 << 19) + 3, 3 * (1 << 18), (1 << 20) + (1 << 17) + 1, (1 << 21) - 1, (1 << 21) + 1,
             (1 << 21) + (1 << 17), nmax]
    for n in sizes:
        ok = eng.verify_batch(host["r8x"][:n], host["r8y"][:n], host["s"][:n], host["ax"][:n], host["ay"][:n], msgs[:n])
        exp = np.ones(n, dtype=np.uint8)
        exp[bad[bad < n]] = 0
        assert ok.shape == (n,) and np.array_equal(ok, exp), n
    for n in ((1 << 18) + 1, 3 * (1 << 18) + 7, (1 << 20) + (1 << 18) + 1):
        px, py = eng.public_batch(keys[:n])
        assert np.array_equal(px, host["ax"][:n]) and np.array_equal(py, host["ay"][:n]), n
    idx = np.concatenate([np.arange(4), bad[:8], np.arange(nmax - 4, nmax)])
    assert np.array_equal(exp[idx], oracle_c.verify(host["r8x"][idx], host["r8y"][idx], host["s"][idx], host["ax"][idx], host["ay"][idx], msgs[idx]))


def test_config1_bench_workloads_1024(gpu, oracle_c):
    """BASELINE config 1: the criterion workloads of benches/bench_babyjubjub.rs (add, mul_scalar_small,
    mul_scalar, compress, decompress, sign, verify) on a 1,024-element synthetic batch, lane for lane
    against the oracle.  Lane 0 carries the bench's own inputs (benches/bench_babyjubjub.rs:15-53)."""
    import random
    from common import P_KAT
    n = 1024
    rnd = random.Random(0xB200)
    pts = [P_KAT] + [O.mul_scalar(O.B8, rnd.randrange(1 << 251)) for _ in range(63)]
    pts = [pts[i % 64] for i in range(n)]
    px, py = pack([p[0] for p in pts]), pack([p[1] for p in pts])
    one = pack([1] * n)
    # add (projective, bench: p + p)
    qx, qy = np.roll(px, 1, axis=0), np.roll(py, 1, axis=0)
    qx[0], qy[0] = px[0], py[0]
    got, exp = gpu.add(px, py, one, qx, qy, one), oracle_c.add(px, py, one, qx, qy, one)
    assert all(np.array_equal(g, e) for g, e in zip(got, exp))
    # mul_scalar_small (3) and mul_scalar (the bench's 251-bit scalar on lane 0, uniform 254-bit elsewhere)
    k3 = pack([3] * n)
    kb = pack([2626589144620713026669568689430873010625803728049924121243784502389097019475] +
              [rnd.randrange(1 << 254) for _ in range(n - 1)])
    for k in (k3, kb):
        got, exp = gpu.mul_scalar(px, py, k), oracle_c.mul_scalar(px, py, k)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
    # compress / decompress
    comp = gpu.compress(px, py)
    assert np.array_equal(comp, oracle_c.compress(px, py))
    dx, dy, st = gpu.decompress(comp)
    assert not st.any() and np.array_equal(dx, px) and np.array_equal(dy, py)
    # sign (msg = 5 on lane 0) and verify
    keys = np.frombuffer(b"".join(rnd.randbytes(32) for _ in range(n)), dtype=np.uint8).reshape(n, 32)
    msgs = pack([5] + [rnd.randrange(Q) for _ in range(n - 1)])
    g = gpu.sign(keys, msgs)
    e = oracle_c.sign(keys, msgs)
    assert all(np.array_equal(a, b) for a, b in zip(g, e))
    ax, ay = gpu.public(keys)
    ok = gpu.verify(g[0], g[1], g[2], ax, ay, msgs)
    assert ok.all() and np.array_equal(ok, oracle_c.verify(g[0], g[1], g[2], ax, ay, msgs))


def test_verify_random_corruption_mix_vs_oracle(gpu, oracle_c):
    """16,384 signatures with the benchmark's 10 % corruption mix (8 classes), every lane against the oracle"""
    import importlib.util
    import os
    from common import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n = 1 << 14
    rng = np.random.default_rng(99)
    keys = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    msgs[:, 31] &= 0x1F
    rx, ry, s, st = gpu.sign(keys, msgs)
    assert not st.any()
    ax, ay = gpu.public(keys)
    cols = [rx.copy(), ry.copy(), s.copy(), ax.copy(), ay.copy(), msgs.copy()]
    expected, cls = bench.corrupt(cols, 1234, 4)          # 25 % corrupted: more lanes per class
    assert all((cls == c).sum() > 100 for c in range(len(bench.CORRUPTIONS)))
    ok = gpu.verify(*cols)
    ref = oracle_c.verify(*cols)
    assert np.array_equal(ok, ref)
    assert np.array_equal(ok, expected)
    # and the compressed pipeline on the same (decodable) signatures
    sig64 = np.concatenate([gpu.compress(cols[0], cols[1]), cols[2]], axis=1)
    pk32 = gpu.compress(cols[3], cols[4])
    gok, gst = gpu.verify_compressed(sig64, pk32, cols[5])
    eok, est = oracle_c.verify_compressed(sig64, pk32, cols[5])
    assert np.array_equal(gst, est) and np.array_equal(gok, eok)


def test_abi_argument_errors_and_context_lifecycle(gpu):
    """C-ABI error behaviour: null pointers / bad counts are BJJ_ERR_ARG, contexts are independent and can be
    created, used from their own host threads, and destroyed repeatedly"""
    import ctypes
    import threading
    bjj = gpu.bjj
    lib = gpu.eng.lib
    ctx = gpu.eng.ctx
    buf = pack([1, 2, 3])
    p = buf.ctypes.data_as(ctypes.c_void_p)
    assert lib.bjj_fixed_base_batch(ctx, 3, None, p, p) == 2                       # BJJ_ERR_ARG
    assert lib.bjj_fr_op_batch(ctx, 9, 3, p, p, p) == 2
    assert lib.bjj_poseidon_batch(ctx, 0, 3, (ctypes.c_void_p * 1)(p), p) == 2
    assert lib.bjj_poseidon_batch(ctx, 9, 3, (ctypes.c_void_p * 1)(p), p) == 2
    assert lib.bjj_verify_batch(None, 3, p, p, p, p, p, p, p) == 2
    assert lib.bjj_fixed_base_batch(ctx, 0, p, p, p) == 0                          # empty batch is fine
    assert lib.bjj_status_string(3) == b"not a mod p square"
    assert lib.bjj_status_string(1) == b"y outside the Finite Field over R"
    assert lib.bjj_status_string(4) == b"msg outside the Finite Field"
    assert lib.bjj_error_string(3).startswith(b"field element input >= Q")
    assert lib.bjj_device(ctx) == 0 and lib.bjj_stream(ctx)
    before = gpu.eng.kernel_launches
    gpu.eng.fixed_base_batch(pack([5]))
    assert gpu.eng.kernel_launches == before + 2                                   # comb kernel + batched affine
    # two more contexts on the same device, each driven by its own host thread
    k = pack([(i * 0x9E3779B97F4A7C15 + 7) % (1 << 256) for i in range(3000)])
    ref = gpu.eng.fixed_base_batch(k)
    out = {}

    def work(tag):
        e = bjj.Engine(0)
        for _ in range(3):
            out[tag] = e.fixed_base_batch(k)
        e.close()
    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for tag in range(2):
        assert np.array_equal(out[tag][0], ref[0]) and np.array_equal(out[tag][1], ref[1])
    for _ in range(3):                                                              # create / destroy repeatedly
        e = bjj.Engine(0)
        assert unpack(e.fr_op_batch(bjj.FR_ADD, pack([Q - 1]), pack([2])))[0] == 1
        e.close()


def test_schnorr(gpu, oracle_c):
    parity.check_schnorr(gpu, oracle_c, 64)


def test_split_scalars_on_device(gpu, hostemu):
    """the device code of verify's scalar split on crafted h (degenerate lattices, whole-limb quotients, every
    length) -- inputs verify itself never produces: the relation holds exactly and the device agrees with the
    host build of the same header bit for bit"""
    import ctypes
    from common import check_split_outputs, split_scalar_inputs
    hs, ss = split_scalar_inputs(n_random=4000)
    H, S = pack(hs), pack(ss)
    u, v, neg, w = gpu.eng.split_scalars_batch(H, S)
    check_split_outputs(hs, ss, unpack(u), unpack(v), neg, unpack(w), max_wide=400)
    n = len(hs)
    U, V, W = np.zeros_like(H), np.zeros_like(H), np.zeros_like(H)
    neg_h = np.zeros(n, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    hostemu.lib.emu_split(ctypes.c_size_t(n), p(H), p(S), p(U), p(V), p(neg_h), p(W))
    assert np.array_equal(u, U) and np.array_equal(v, V) and np.array_equal(neg, neg_h) and np.array_equal(w, W)


def test_wide_scalars_match_the_reference_bigint_semantics(gpu):
    """The reference takes BigInt scalars of any size (src/lib.rs:149-164 loops over n.bits(); verify hands S to
    B8.mul_scalar unreduced, :405).  On-curve points: a wide scalar acts mod the group order; off-curve points: every
    bit of the wide scalar is replayed.  Checked against the integer oracle, which runs the literal loop."""
    import random
    bjj = gpu.bjj
    rnd = random.Random(2024)
    b8 = bjj.Point(*O.B8)
    off = bjj.Point((O.B8[0] + 1) % O.Q, O.B8[1])            # x + 1: not on the curve
    assert O.on_curve(O.B8) and not O.on_curve((off.x, off.y))
    pts, ks = [], []
    for bits in (257, 300, 511, 512, 1024, 2048):
        for p in (b8, off, bjj.Point(0, 1), bjj.Point(0, 0)):
            pts.append(p)
            ks.append(rnd.getrandbits(bits) | (1 << (bits - 1)))
    # mixed widths in one batch, a negative scalar (sign dropped, src/lib.rs:156) and a narrow one beside wide ones
    pts += [b8, off, b8]
    ks += [-(rnd.getrandbits(700)), 5, O.SUBORDER * 8 + 3]
    got = bjj.mul_scalar_batch(pts, ks)
    for p, k, g in zip(pts, ks, got):
        ex, ey = O.mul_scalar((p.x, p.y), abs(k))
        assert (g.x, g.y) == (ex, ey), (p, k.bit_length())
    # single-point API with a wide scalar
    k = rnd.getrandbits(900)
    assert (lambda r: (r.x, r.y))(off.mul_scalar(k)) == O.mul_scalar((off.x, off.y), k)
    # verify with S >= 2^256: same group element as S mod SUBORDER (B8 has order SUBORDER)
    key = bytes(rnd.getrandbits(8) for _ in range(32))
    msg = rnd.getrandbits(250)
    (r8, s) = O.sign(key, msg)
    pk = O.public(key)
    for s_wide in (s + (1 << 256) * 7 * O.SUBORDER, s + O.SUBORDER * (1 << 300), -(s + O.SUBORDER * (1 << 260)), s + (1 << 256)):
        exp = O.verify(pk, (r8, abs(s_wide)), msg)
        got = bjj.verify(bjj.Point(*pk), bjj.Signature(bjj.Point(*r8), s_wide), msg)
        assert got == exp, s_wide.bit_length()
    oks = bjj.verify_batch([bjj.Point(*pk)] * 2, [bjj.Signature(bjj.Point(*r8), s + (O.SUBORDER << 270)),
                                                    bjj.Signature(bjj.Point(*r8), s + (1 << 256))], [msg, msg])
    assert oks == [True, O.verify(pk, (r8, s + (1 << 256)), msg)]


def test_alternative_paths_selected_by_environment(oracle_c):
    """the paths a context selects at bjj_init from the environment stay bit-exact: verify WITHOUT the half-size scalars
    (BJJ_VERIFY_SPLIT=0: EdDSA through the 64-window Straus pass) and sign as ONE fused kernel (BJJ_SIGN_FUSED=1)
    instead of the pipeline of kernels, public keys through the fused k_public (BJJ_PUBLIC_FUSED=1)"""
    from common import Gpu
    old = {k: os.environ.get(k) for k in ("BJJ_VERIFY_SPLIT", "BJJ_SIGN_FUSED", "BJJ_PUBLIC_FUSED")}
    os.environ["BJJ_VERIFY_SPLIT"] = "0"
    os.environ["BJJ_SIGN_FUSED"] = "1"
    os.environ["BJJ_PUBLIC_FUSED"] = "1"
    try:
        alt = Gpu(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        parity.check_verify(alt, oracle_c, 32)
        parity.check_sign(alt, oracle_c, 64)
        parity.check_fixed_base(alt, oracle_c, 64)
    finally:
        alt.eng.close()
