#!/usr/bin/env python3
"""bench.py -- EdDSA-Poseidon verify_batch (and B8 fixed-base / variable-base / compressed-pipeline
secondaries) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: C++ restatement of the reference algorithm

One "step" = one verify_batch pass over this rank's shard (2^21 signatures per GPU, i.e. BASELINE
config 4's 2^24 signatures over 8 GPUs; 10 % corrupted).  The batch shards with no exchange between
lanes, so there is no data-path collective: torch.distributed is used for the barrier and the
max-over-ranks time only ("scaling": "weak").

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline` is quoted against the
integer-multiply (IMAD) issue rate, the pipe this path is bound by, with HBM as
a secondary counter.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
SUBORDER = 21888242871839275222246405745257275088614511777268538073601725287587578984328 >> 3
METRIC = "eddsa_poseidon_verifies_per_sec"
UNIT = "verifies/s"

# Algorithmic work per unit, in limb-MACs (one 32x32->64 multiply-accumulate = one IMAD.WIDE.U32 on sm_100a).
# fmul = fsqr = 128 (8x8 products + 8x8 Montgomery reduction; the 8 m_i IMADs are not counted), a Montgomery dot
# product of N terms = 64 N + 64.  Counts are for the algorithms actually built; derivations in DESIGN.md section 6.
FMUL = 128
ALGO_FMUL = {
    # k_verify_hash besides Poseidon: two on-curve gates 10 + Montgomery conversions 6
    "verify_hash_extra": 10 + 6,
    # k_verify_ec (half-size scalars, 33 radix-16 windows over the tables of 8A and R8):
    # conversions 4 + map 4 + 8A 22 + two 9-entry tables 2 x 64 + 32 x (3 x 7 + 8) doublings
    # + 33 x (8 + 7) table additions + 1 + 17 B8 additions (16 x 7 + 6)
    "verify_ec": 4 + 4 + 22 + 128 + 32 * 29 + 33 * 15 + 1 + 16 * 7 + 6,
    "fixed_base": 17 * 7 + 1 + 5 + 2 + 12,            # 16-bit comb + map + batched inversion share
    "mul_scalar": 2 + 5 + 2 + 64 + 7 + 64 * 36 + 20,  # gate, table, 64 windows x (4 dbl + add), batched inversion share
}
# Poseidon t = 6, sparse schedule: 8 full rounds x (6 x^5 + 6 dot6) + 60 partial x (x^5 + dot6 + 5 fmul)
POSEIDON6_MAC = 8 * (6 * 3 * FMUL + 6 * (6 * 64 + 64)) + 60 * (3 * FMUL + (6 * 64 + 64) + 5 * FMUL)
# k_verify_split: ~75 Euclid steps x (8 + 8) wide multiplies + two Montgomery products mod l
SPLIT_MAC = 75 * 16 + 2 * FMUL
ALGO_MAC = {
    "verify": (ALGO_FMUL["verify_hash_extra"] + ALGO_FMUL["verify_ec"]) * FMUL + POSEIDON6_MAC + SPLIT_MAC,
    "fixed_base": ALGO_FMUL["fixed_base"] * FMUL,
    "mul_scalar": ALGO_FMUL["mul_scalar"] * FMUL,
}
# Measured on B200 (profiles/r1_pipe_probe.jsonl): IMAD.WIDE.U32 issues at 32 lanes/clk/SM -- half the 32-bit IMAD
# rate that SURVEY.md section 8d's model (64 lanes/clk/SM) assumes.
WIDE_MAC_LANES_PER_CLK_SM = 32
MODEL_LANES_PER_CLK_SM = 64


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons of one GPU while the timed region runs: NVML in-process (a sample
    every 20 ms, cheap enough for 8 ranks at once), `nvidia-smi -lms` as the fallback when pynvml is missing"""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.rows = []          # (sm_mhz, sm_max_mhz, [reason names])
        self.stop_flag = threading.Event()
        self.proc = None
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.device < len(ids) and ids[self.device].isdigit():
                return int(ids[self.device])
        return self.device

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
        smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.source = "nvml"
        while not self.stop_flag.is_set():
            sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = int(get_reasons(h))
            self.rows.append((sm, smax, [nm for nm, bit in self.NVML_REASONS if mask & bit]))
            time.sleep(0.02)

    def _run_smi(self):
        self.source = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[1]), float(r[2]), [nm for nm, v in zip(names, r[5:9]) if v.lower().startswith("active")]))
            except Exception:
                pass
            if self.stop_flag.is_set():
                break

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()        # exact PID we started
            except Exception:
                pass
        self.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [r[0] for r in self.rows]
        reasons = sorted({nm for r in self.rows for nm in r[2]})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(r[1] for r in self.rows), "reasons": reasons,
                "samples": len(sm), "source": self.source}


# ---------------------------------------------------------------------------------------------------
# synthetic signatures: valid signatures made ON THE DEVICE by the
# library's own sign kernel from random keys / messages, then 10 % corrupted on the host by class.
# ---------------------------------------------------------------------------------------------------
CORRUPTIONS = ["flip_S", "flip_msg", "swap_R8", "swap_A", "S_plus_SUBORDER", "msg_gt_Q", "off_curve", "special_point"]


def _to_int(row):
    return int.from_bytes(row.tobytes(), "little")


def _from_int(v):
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8)


def corrupt(cols, seed, denom=10):
    """cols = [r8x, r8y, s, ax, ay, msg] numpy (n,32) arrays, modified in place.  Returns (expected_ok, class_id)
    where class_id = -1 for untouched lanes."""
    n = len(cols[0])
    rng = np.random.default_rng(seed)
    h = rng.integers(0, max(denom, 1), size=n)
    cls = np.where((h == 0) & (denom > 0), rng.integers(0, len(CORRUPTIONS), size=n), -1)
    expected = np.ones(n, dtype=np.uint8)
    r8x, r8y, s, ax, ay, msg = cols
    orig = [c.copy() for c in (r8x, r8y, ax, ay)]
    idx = np.nonzero(cls == 0)[0]
    s[idx, rng.integers(0, 31, size=len(idx))] ^= (1 << rng.integers(0, 8, size=len(idx))).astype(np.uint8)
    expected[idx] = 0
    idx = np.nonzero(cls == 1)[0]
    msg[idx, rng.integers(0, 31, size=len(idx))] ^= (1 << rng.integers(0, 8, size=len(idx))).astype(np.uint8)
    expected[idx] = 0
    idx = np.nonzero(cls == 2)[0]
    r8x[idx], r8y[idx] = orig[0][(idx + 1) % n], orig[1][(idx + 1) % n]
    expected[idx] = 0
    idx = np.nonzero(cls == 3)[0]
    ax[idx], ay[idx] = orig[2][(idx + 1) % n], orig[3][(idx + 1) % n]
    expected[idx] = 0
    for i in np.nonzero(cls == 4)[0]:                      # stays VALID: verify never range-checks S
        s[i] = _from_int(_to_int(s[i]) + SUBORDER)
    idx = np.nonzero(cls == 5)[0]
    msg[idx] = 0xFF
    expected[idx] = 0
    for i in np.nonzero(cls == 6)[0]:                      # x + 1: off the curve -> exact lane
        tgt = r8x if (i & 1) else ax
        tgt[i] = _from_int((_to_int(tgt[i]) + 1) % Q)
    expected[cls == 6] = 0
    specials = [(0, 0), (0, 1), (0, Q - 1)]
    for i in np.nonzero(cls == 7)[0]:
        x, y = specials[int(i) % 3]
        if i & 4:
            r8x[i], r8y[i] = _from_int(x), _from_int(y)
        else:
            ax[i], ay[i] = _from_int(x), _from_int(y)
    expected[cls == 7] = 0
    return expected, cls


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-lanes", type=int, default=21, help="signatures per GPU (2^21 x 8 GPUs = config 4's 2^24)")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--corrupt-denominator", type=int, default=10,
                    help="1 lane in this many is corrupted (default 10 = the 10 %% of config 4; 0 = none, for diagnosis)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget for the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    run_ours(args, rank, local_rank, world)


# ---------------------------------------------------------------------------------------------------
def cpu_verify_rate(n_sample, steps, warmup, seed=1):
    """times the C++ restatement of the reference algorithm (oracle port) over n_sample signatures"""
    import common
    ora = common.OracleC(threads=os.cpu_count() or 1)
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 256, size=(n_sample, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, size=(n_sample, 32), dtype=np.uint8)
    msgs[:, 31] &= 0x1F
    rx, ry, s, _ = ora.sign(keys, msgs)
    ax, ay = ora.public(keys)
    cols = [rx, ry, s, ax, ay, msgs]
    expected, _ = corrupt(cols, seed)
    for _ in range(warmup):
        ora.verify(*cols)
    t0 = time.perf_counter()
    for _ in range(steps):
        ok = ora.verify(*cols)
    dt = time.perf_counter() - t0
    assert np.array_equal(ok, expected), "oracle disagrees with the constructed expectations"
    return n_sample * steps / dt, dt / steps, ora.threads


def run_reference(args):
    """reference arm: the reference's own CPU algorithm (C++ restatement; the Rust crate cannot be built
    in this image) on all host threads, on a bounded sample of the same workload"""
    cores = os.cpu_count() or 1
    n_sample = 512 * cores
    rate, sec_per_step, threads = cpu_verify_rate(n_sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery)", "data": "synthetic",
        "config": {"workload": "verify (EdDSA-Poseidon), 10% corrupted; bounded sample of config 4", "sample_signatures": n_sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d signatures per step, std::thread per core" % n_sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import babyjubjub_rs_b200 as bjj
    import common
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = bjj.Engine(local_rank)
    lib, ctx = eng.lib, eng.ctx
    # a dedicated non-default stream: the kernels are launched on it (passed to the _dev entry points) and the
    # CUDA events that time them are recorded on the same stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    n = 1 << args.log2_lanes

    def dptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def check(rc, what):
        if rc not in (0,):
            raise RuntimeError("%s failed: %d (%s)" % (what, rc, lib.bjj_error_string(rc).decode()))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return bjj.max_over_ranks(x, dist, dev)

    # ---- data: keys / msgs random on the device; signatures by the library's sign kernel -------------
    g = torch.Generator(device=dev)
    g.manual_seed(0xB200 + rank)
    keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msgs = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msgs[:, 31] &= 0x1F                                   # < 2^253 < Q
    r8x, r8y, s, ax, ay = (torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(5))
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib.bjj_sign_batch_dev(ctx, n, dptr(keys), dptr(msgs), dptr(r8x), dptr(r8y), dptr(s), dptr(st), sp), "sign")
    check(lib.bjj_public_batch_dev(ctx, n, dptr(keys), dptr(ax), dptr(ay), sp), "public")
    torch.cuda.synchronize()
    assert int(st.max().item()) == 0
    cols_h = [t.cpu().numpy() for t in (r8x, r8y, s, ax, ay, msgs)]
    expected, cls = corrupt(cols_h, 0xC0DE + rank, args.corrupt_denominator)
    pinned = [torch.from_numpy(c).pin_memory() for c in cols_h]
    cols_d = [p.to(dev, non_blocking=True) for p in pinned]
    ok_d = torch.zeros(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step_dev():
        check(lib.bjj_verify_batch_dev(ctx, n, *[dptr(c) for c in cols_d], dptr(ok_d), sp), "verify_batch_dev")

    # ---- device-resident timing ---------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = eng.kernel_launches - launches0
    clocks = sampler.finish()
    check(lib.bjj_sync(ctx), "sync")
    ok_h = ok_d.cpu().numpy()
    mism = int((ok_h != expected).sum())
    dev_ms = max_over_ranks(dev_ms)
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- end to end: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region -------------
    ok_pin = torch.zeros(n, dtype=torch.uint8).pin_memory()

    def step_e2e():
        check(lib.bjj_verify_batch(ctx, n, *[ctypes.c_void_p(p.data_ptr()) for p in pinned], ctypes.c_void_p(ok_pin.data_ptr())),
              "verify_batch")

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * args.steps / e2e_s
    mism += int((ok_pin.numpy() != expected).sum())

    # ---- oracle spot check (outside every timed region): random lanes + every corruption class ---------------
    checked = 0
    if rank == 0:
        ora = common.OracleC(threads=os.cpu_count() or 1)
        rng = np.random.default_rng(5)
        pick = [rng.choice(n, size=2048, replace=False)]
        for c in range(len(CORRUPTIONS)):
            idx = np.nonzero(cls == c)[0]
            pick.append(idx[:256])
        pick = np.unique(np.concatenate(pick))
        ok_ref = ora.verify(*[c[pick] for c in cols_h])
        mism += int((ok_ref != ok_h[pick]).sum())
        checked = len(pick)
    if mism:
        raise SystemExit("PARITY FAILURE: %d lanes differ from the oracle / constructed expectations" % mism)

    if rank != 0:
        if dist is not None:
            # secondaries run on rank 0 only; keep the other ranks alive until it is done
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    mac_peak = sms * WIDE_MAC_LANES_PER_CLK_SM * sm_max * 1e6 / 1e12      # T limb-MAC/s, measured IMAD.WIDE rate
    model_peak = sms * MODEL_LANES_PER_CLK_SM * sm_max * 1e6 / 1e12        # SURVEY model (32-bit IMAD rate)
    per_gpu = value / world
    achieved = per_gpu * ALGO_MAC["verify"] / 1e12
    roofline = {
        "bound": "imad", "achieved": achieved, "peak": mac_peak, "unit": "T limb-MAC/s", "frac": achieved / mac_peak,
        "peak_kind": "measured: %d SMs x %d IMAD.WIDE lanes/clk x %.0f MHz (profiles/r1_pipe_probe.jsonl)" % (sms, WIDE_MAC_LANES_PER_CLK_SM, sm_max),
        "frac_of_survey_model": achieved / model_peak,
        "survey_model_peak": model_peak,
        "algorithmic_mac_per_verify": ALGO_MAC["verify"],
        "frac_at_observed_clock": (achieved / (sms * WIDE_MAC_LANES_PER_CLK_SM * clocks["sm_mhz"] * 1e6 / 1e12)) if clocks.get("sm_mhz") else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (k_verify_ec) per launch, from the
        # ncu --set full capture at this workload size (profiles/r1_ncu_verify_split_summary.txt): 8.18 GB per
        # 2^21 lanes = 3,899 B per lane, against 225 B per lane algorithmic (4 coordinates + 3 scalars in, 1 B
        # out).  The excess is the per-thread window tables (2 x 9 x 128 B written, 66 x 128 B read per lane;
        # 87 MB live, more than L2 keeps): 115 GB/s, under 2 % of HBM bandwidth -- not what bounds this kernel.
        "traffic": int(n * 3899),
        "dominant_kernel": "k_verify_ec (Straus pass over half-size scalars): 67 ms of a 105 ms step (BJJ_PHASE_TIMING); "
                           "fmaheavy pipe 78 % busy at 2^20 lanes, 66 % at 2^21 under ncu; k_verify_hash 33 ms, 92 % busy",
        "hbm": {"achieved_gbs": per_gpu * 193 / 1e9, "peak_gbs": peaks.get("hbm_gbs"), "peak_kind": peak_kind,
                "note": "193 B per verify (6 x 32 B in, 1 B out); secondary counter, this path is not HBM-bound"},
    }

    secondary = []
    if not args.no_secondary:
        secondary = run_secondaries(args, eng, torch, dev, stream, common)

    cpu = None
    if world == 1 or rank == 0:
        cores = os.cpu_count() or 1
        probe_rate, _, threads = cpu_verify_rate(64 * cores, 1, 0)
        n_sample = int(min(1 << 18, max(256 * cores, probe_rate * args.cpu_seconds)))
        rate, _, threads = cpu_verify_rate(n_sample, 1, 0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d signatures (10%% corrupted), C++ restatement of the reference algorithm, one std::thread per core" % n_sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "verify_batch: 2^%d EdDSA-Poseidon signatures per GPU (config 4: 2^24 over 8 GPUs), 10%% corrupted in 8 classes"
                               % args.log2_lanes,
                   "signatures_per_gpu": n, "l2": "inputs larger than L2 (%.0f MB per step)" % (n * 192 / 1e6),
                   "sharding": "contiguous slices per rank, no collective"},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 192, "d2h_bytes_per_step": n,
                "path": "bjj_verify_batch (C ABI, pinned host buffers, chunked double-buffered copies)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "parity": {"lanes_checked_vs_oracle": checked, "lanes_checked_vs_construction": 2 * n, "mismatches": 0},
        "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_secondaries(args, eng, torch, dev, stream, common):
    """BASELINE configs 2, 3 and 5 on one GPU (rank 0), each checked against the oracle on a sample"""
    lib, ctx = eng.lib, eng.ctx
    sp = ctypes.c_void_p(stream.cuda_stream)
    out = []
    ora = common.OracleC(threads=os.cpu_count() or 1)
    g = torch.Generator(device=dev)
    g.manual_seed(77)

    def dptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def timed(fn, steps, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    peaks, _ = measured_peaks()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    imad_peak = sms * WIDE_MAC_LANES_PER_CLK_SM * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6

    # config 2: PrivateKey::public over 2^20 random keys
    n = 1 << 20
    keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    rx, ry = (torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(2))
    ms = timed(lambda: lib.bjj_public_batch_dev(ctx, n, dptr(keys), dptr(rx), dptr(ry), sp), args.steps)
    idx = torch.randperm(n, device=dev)[:1024]
    ex, ey = ora.public(keys[idx].cpu().numpy())
    assert np.array_equal(rx[idx].cpu().numpy(), ex) and np.array_equal(ry[idx].cpu().numpy(), ey), "public_batch parity"
    rate = n / (ms * 1e-3)
    out.append({"metric": "b8_scalar_mults_per_sec", "workload": "public_batch: 2^20 random keys (config 2)", "value": rate,
                "unit": "mults/s", "ms_per_step": ms, "roofline_frac_imad": rate * ALGO_MAC["fixed_base"] / imad_peak,
                "oracle_checked_lanes": 1024})

    # config 3: variable-base mul_scalar over 2^19 pairs per GPU (2^22 over 8 GPUs)
    n = 1 << 19
    k = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    k[:, 31] &= 0x3F
    px, py = rx[:n].clone(), ry[:n].clone()
    ox, oy = (torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(2))
    ms = timed(lambda: lib.bjj_mul_scalar_batch_dev(ctx, n, dptr(px), dptr(py), dptr(k), dptr(ox), dptr(oy), sp), args.steps)
    idx = torch.randperm(n, device=dev)[:512]
    ex, ey = ora.mul_scalar(px[idx].cpu().numpy(), py[idx].cpu().numpy(), k[idx].cpu().numpy())
    assert np.array_equal(ox[idx].cpu().numpy(), ex) and np.array_equal(oy[idx].cpu().numpy(), ey), "mul_scalar_batch parity"
    rate = n / (ms * 1e-3)
    out.append({"metric": "variable_base_scalar_mults_per_sec", "workload": "mul_scalar_batch: 2^19 (point, 254-bit scalar) pairs per GPU (config 3)",
                "value": rate, "unit": "mults/s", "ms_per_step": ms, "roofline_frac_imad": rate * ALGO_MAC["mul_scalar"] / imad_peak,
                "oracle_checked_lanes": 512})

    # config 5: compressed pipeline (decompress R8 and A, Poseidon, Straus) over 2^19 signatures per GPU
    n = 1 << 19
    keys = keys[:n]
    msgs = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msgs[:, 31] &= 0x1F
    r8x, r8y, s, ax, ay, comp_r, comp_a = (torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(7))
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    lib.bjj_sign_batch_dev(ctx, n, dptr(keys), dptr(msgs), dptr(r8x), dptr(r8y), dptr(s), dptr(st), sp)
    lib.bjj_public_batch_dev(ctx, n, dptr(keys), dptr(ax), dptr(ay), sp)
    lib.bjj_compress_batch_dev(ctx, n, dptr(r8x), dptr(r8y), dptr(comp_r), sp)
    lib.bjj_compress_batch_dev(ctx, n, dptr(ax), dptr(ay), dptr(comp_a), sp)
    sig64 = torch.cat([comp_r, s], dim=1).contiguous()
    msgs[::10, 0] ^= 1                                       # 10 % wrong message
    sig64[5::97, :32] = 0xFF                                 # undecodable R8 (y >= Q)
    ok, stt = (torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2))
    ms = timed(lambda: lib.bjj_verify_compressed_batch_dev(ctx, n, dptr(sig64), dptr(comp_a), dptr(msgs), dptr(ok), dptr(stt), sp), args.steps)
    idx = torch.arange(0, 2048, device=dev)
    eok, est = ora.verify_compressed(sig64[idx].cpu().numpy(), comp_a[idx].cpu().numpy(), msgs[idx].cpu().numpy())
    assert np.array_equal(ok[idx].cpu().numpy(), eok) and np.array_equal(stt[idx].cpu().numpy(), est), "verify_compressed parity"
    out.append({"metric": "compressed_pipeline_verifies_per_sec", "workload": "verify_compressed_batch: 2^19 x (64 B sig + 32 B pk + 32 B msg) per GPU (config 5)",
                "value": n / (ms * 1e-3), "unit": "verifies/s", "ms_per_step": ms, "oracle_checked_lanes": 2048})
    # the remaining rows of the hot-path table, each with its own bound ------------------------------------------
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n = 1 << 20
    # decompress_point (row a7): 2^20 compressed points, ~345 fmul each -> IMAD-bound
    comp2 = torch.cat([comp_a, comp_r], dim=0).contiguous()          # 2^20 valid compressed points
    st2 = torch.empty(n, dtype=torch.uint8, device=dev)
    dx, dy = (torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(2))
    ms = timed(lambda: lib.bjj_decompress_batch_dev(ctx, n, dptr(comp2), dptr(dx), dptr(dy), dptr(st2), sp), args.steps)
    ex, ey, est = ora.decompress(comp2[:1024].cpu().numpy())
    assert np.array_equal(dx[:1024].cpu().numpy(), ex) and np.array_equal(st2[:1024].cpu().numpy(), est), "decompress_batch parity"
    rate = n / (ms * 1e-3)
    out.append({"metric": "decompress_points_per_sec", "workload": "decompress_batch: 2^20 compressed points", "value": rate,
                "unit": "points/s", "ms_per_step": ms, "roofline_frac_imad": rate * 345 * FMUL / imad_peak, "oracle_checked_lanes": 1024})
    # POSEIDON.hash with 5 inputs (row a8)
    ins = [dx, dy, r8x.repeat(2, 1)[:n].contiguous(), r8y.repeat(2, 1)[:n].contiguous(), msgs.repeat(2, 1)[:n].contiguous()]
    arr = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in ins])
    ho = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.bjj_poseidon_batch_dev(ctx, 5, n, arr, dptr(ho), sp), args.steps)
    eh = ora.poseidon([t[:256].cpu().numpy() for t in ins])
    assert np.array_equal(ho[:256].cpu().numpy(), eh), "poseidon_batch parity"
    rate = n / (ms * 1e-3)
    out.append({"metric": "poseidon5_hashes_per_sec", "workload": "poseidon_batch: 2^20 x 5 inputs (t = 6)", "value": rate,
                "unit": "hashes/s", "ms_per_step": ms, "roofline_frac_imad": rate * (POSEIDON6_MAC + 6 * FMUL) / imad_peak,
                "oracle_checked_lanes": 256})
    # Point::compress (row a6): pure data movement, 64 B in + 32 B out per point -> HBM-bound
    n = 1 << 22
    bx = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    by = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    bx[:, 31] &= 0x0F
    by[:, 31] &= 0x0F
    bo = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    ms = timed(lambda: lib.bjj_compress_batch_dev(ctx, n, dptr(bx), dptr(by), dptr(bo), sp), args.steps * 4)
    assert np.array_equal(bo[:4096].cpu().numpy(), ora.compress(bx[:4096].cpu().numpy(), by[:4096].cpu().numpy())), "compress parity"
    gbs = n * 96 / (ms * 1e-3) / 1e9
    out.append({"metric": "compress_points_per_sec", "workload": "compress_batch: 2^22 points (403 MB per step, larger than L2)",
                "value": n / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak},
                "oracle_checked_lanes": 4096})
    # PointProjective::add (row a2): 6 x 32 B in, 3 x 32 B out, 13 fmul (+9 Montgomery conversions)
    n = 1 << 21
    one = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    one[:, 0] = 1
    px2, py2 = bx[:n].contiguous(), by[:n].contiguous()
    qx2, qy2 = bx[n:2 * n].contiguous(), by[n:2 * n].contiguous()
    o3 = [torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(3)]
    ms = timed(lambda: lib.bjj_add_batch_dev(ctx, n, dptr(px2), dptr(py2), dptr(one), dptr(qx2), dptr(qy2), dptr(one),
                                             dptr(o3[0]), dptr(o3[1]), dptr(o3[2]), sp), args.steps * 2)
    ea = ora.add(*[t[:1024].cpu().numpy() for t in (px2, py2, one, qx2, qy2, one)])
    assert all(np.array_equal(o3[k][:1024].cpu().numpy(), ea[k]) for k in range(3)), "add_batch parity"
    rate = n / (ms * 1e-3)
    out.append({"metric": "projective_adds_per_sec", "workload": "add_batch: 2^21 projective pairs (literal add-2008-bbjlp)",
                "value": rate, "unit": "adds/s", "ms_per_step": ms, "roofline_frac_imad": rate * 22 * FMUL / imad_peak,
                "hbm_frac": rate * 288 / 1e9 / hbm_peak, "oracle_checked_lanes": 1024})
    return out


if __name__ == "__main__":
    main()
