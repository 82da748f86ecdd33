#!/bin/bash
# A/B of whole-library variants on ONE GPU:  ab_bench.sh <outdir> <variant> [<variant> ...]   ("main" = the library as built)
# Swaps ab_variants/<variant>/libbjj_cuda.so in, runs a short bench.py with every secondary row, restores the library.
out=$1; shift
mkdir -p "$out"
cp babyjubjub-rs_b200/libbjj_cuda.so /tmp/keep.so
for v in "$@"; do
  if [ "$v" != main ]; then cp ab_variants/$v/libbjj_cuda.so babyjubjub-rs_b200/libbjj_cuda.so; else cp /tmp/keep.so babyjubjub-rs_b200/libbjj_cuda.so; fi
  python bench.py --steps 3 --warmup 3 --cpu-seconds 1 --log2-total 21 > "$out/bench_$v.json" 2> "$out/bench_$v.err"
  python - "$out/bench_$v.json" "$v" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
rows = [("verify", d["value"], d["ms_per_step"]), ("verify e2e", d["e2e"]["value"], None)]
s = d.get("second_metric") or {}
rows.append(("public", s.get("value"), s.get("ms_per_step")))
rows.append(("public e2e", (s.get("e2e") or {}).get("value"), None))
for c in d.get("configs") or []:
    rows.append(("config %d" % c["config"], c["value"], c["ms_per_step"]))
for c in d.get("secondary") or []:
    rows.append((c["metric"], c["value"], c["ms_per_step"]))
print("== %s  (%s)" % (sys.argv[2], d["gpu"]["uuid"]))
for name, v, ms in rows:
    print("  %-36s %14.4g %s" % (name, v or 0, ("%8.3f ms" % ms) if ms else ""))
P
done
cp /tmp/keep.so babyjubjub-rs_b200/libbjj_cuda.so
