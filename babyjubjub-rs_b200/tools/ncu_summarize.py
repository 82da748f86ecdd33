#!/usr/bin/env python3
"""ncu report -> the per-kernel counters the roofline claims rest on.

    python babyjubjub-rs_b200/tools/ncu_summarize.py <report.ncu-rep> <out-prefix> [lanes]

Writes <out-prefix>_ncu_kernels_summary.txt (human readable) and <out-prefix>_kernel_metrics.json (read by bench.py).
For every kernel in the report: duration, registers, multiplier-pipe (fmaheavy) and ALU activity, issue slots, the
main stall reasons, DRAM bytes, and -- from the per-instruction executed counts of the source page -- how many
IMAD.WIDE warp-instructions ran, how many other instructions rode the same (fma) pipe, and the ratio of executed
IMAD.WIDE to the algorithmic count of the algorithm where one is known.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.per_cycle_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
FMA_LIGHT = ("IMAD.MOV", "IMAD.IADD", "IMAD.X", "IMAD.SHL", "IMAD", "HFMA2", "IMAD.HI", "FFMA", "FMUL", "FADD")
# algorithmic IMAD.WIDE per lane by kernel-name prefix: bench.py::KERNEL_MAC (fmul = 128, fsqr = 100)
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import KERNEL_MAC as ALGO_WIDE_PER_LANE  # noqa: E402
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_MS = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def opcode_class(src):
    toks = src.split()
    if not toks:
        return "?"
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    base = op.split(".")[0]
    if base == "IMAD":
        for suf in ("WIDE", "MOV", "IADD", "SHL", "HI"):
            if "." + suf in op:
                return "IMAD." + suf
        if op.endswith(".X") or ".X." in op:
            return "IMAD.X"
    return base


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        kernels.append(r)
    metrics = {}
    txt = ["ncu --set full --clock-control none; one launch of every kernel at 2^%d lanes (tools/profile_kernels.py)." % (lanes.bit_length() - 1),
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES and counters, not absolute times.", ""]
    seen = collections.Counter()
    for idx, r in enumerate(kernels):
        name = r[col["Kernel Name"]]
        short = name.split("(")[0]
        seen[short] += 1
        key = short if seen[short] == 1 else "%s#%d" % (short, seen[short])
        m = {}
        txt.append("kernel: %s" % name[:110])
        for h in RAW:
            if h not in col:
                continue
            v, u = r[col[h]], units[col[h]]
            txt.append("  %-88s %-16s %s" % (h, u, v))
            try:
                m[h] = float(v.replace(",", ""))
            except ValueError:
                m[h] = None
            if h == "gpu__time_duration.sum" and m[h] is not None:
                m["ms"] = m[h] * TO_MS.get(u, 1.0)
            if h.startswith("dram__bytes") and m[h] is not None:
                m[h] = m[h] * TO_BYTES.get(u, 1)
        # per-instruction executed counts of this launch
        src = ncu_csv(rep, "source", ["--launch-skip", str(idx), "--launch-count", "1"])
        ex = collections.Counter()
        if len(src) > 2:
            h2 = {h: i for i, h in enumerate(src[1])}
            if "Instructions Executed" in h2:
                for row in src[2:]:
                    if len(row) <= h2["Instructions Executed"]:
                        continue
                    try:
                        ex[opcode_class(row[1])] += int(row[h2["Instructions Executed"]])
                    except ValueError:
                        pass
        total = sum(ex.values())
        # ncu lists the instructions of some kernels once per inlining context: normalise the per-opcode counts to the
        # kernel's own smsp__inst_executed.sum
        raw_total = m.get("smsp__inst_executed.sum")
        if total and raw_total and abs(total / raw_total - 1.0) > 0.01:
            scale = raw_total / total
            ex = collections.Counter({k: int(round(v * scale)) for k, v in ex.items()})
            total = sum(ex.values())
        wide = ex.get("IMAD.WIDE", 0)
        light = sum(v for k, v in ex.items() if k in FMA_LIGHT)
        if total:
            txt.append("  executed warp-instructions: total %d, IMAD.WIDE %d (%.1f %%), other fma-pipe %d (%.1f %% of fma-pipe cycles at 2 vs 4 cycles)"
                       % (total, wide, 100.0 * wide / total, light, 100.0 * 2 * light / max(1, 4 * wide + 2 * light)))
            txt.append("  top opcodes: " + ", ".join("%s %d" % kv for kv in ex.most_common(10)))
        algo = None
        for pfx, v in ALGO_WIDE_PER_LANE.items():
            if pfx in name:
                algo = v
        out = {
            "ms_per_2p20_lanes": m.get("ms"),
            "registers": m.get("launch__registers_per_thread"),
            "fmaheavy_pct": m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "alu_pct": m.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_per_sm": m.get("sm__warps_active.avg.per_cycle_active"),
            "no_instruction_per_issue": m.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
            "math_pipe_throttle_per_issue": m.get("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
            "long_scoreboard_per_issue": m.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            "dram_bytes": (m.get("dram__bytes_read.sum") or 0) + (m.get("dram__bytes_write.sum") or 0),
            "l2_hit_pct": m.get("lts__t_sector_hit_rate.pct"),
            "executed_warp_inst": total, "executed_imad_wide": wide, "executed_fma_pipe_other": light,
            "fma_pipe_passenger_frac": (2.0 * light / (4.0 * wide + 2.0 * light)) if wide else None,
        }
        out["dram_bytes_per_lane"] = out["dram_bytes"] / lanes if out["dram_bytes"] else None
        if algo and wide:
            out["executed_wide_over_algorithmic"] = wide * 32.0 / (lanes * algo)
            txt.append("  executed IMAD.WIDE x 32 lanes / (2^%d lanes x %d algorithmic) = %.3f" % (lanes.bit_length() - 1, algo, out["executed_wide_over_algorithmic"]))
        metrics[key] = out
        txt.append("")
    with open(prefix + "_ncu_kernels_summary.txt", "w") as f:
        f.write("\n".join(txt) + "\n")
    with open(prefix + "_kernel_metrics.json", "w") as f:
        json.dump(metrics, f, indent=1, sort_keys=True)
    print("wrote %d kernels" % len(metrics))


if __name__ == "__main__":
    main()
