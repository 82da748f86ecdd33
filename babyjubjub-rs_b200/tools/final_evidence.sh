#!/bin/bash
# Everything the round's claims rest on, from ONE build on ONE GPU box:  bash babyjubjub-rs_b200/tools/final_evidence.sh <out-dir>
#   1. the GPU parity suite            2. the default bench line (+ its stderr)      3. the reference (CPU) arm
#   4. ncu --set full of one launch of every kernel (summarised)   5. the ncu launch list of a short bench run
#   6. compute-sanitizer over the parity tests of the queue / table / shared-memory kernels
out=${1:-gpurun_out/final}
mkdir -p "$out"
echo "== pytest -m gpu" | tee "$out/log.txt"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a "$out/log.txt"
echo "== bench.py" | tee -a "$out/log.txt"
timeout 1500 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "rc=$?" | tee -a "$out/log.txt"
echo "== bench.py --impl reference" | tee -a "$out/log.txt"
timeout 900 python bench.py --impl reference > "$out/bench_reference.json" 2> "$out/bench_reference.err"; echo "rc=$?" | tee -a "$out/log.txt"
echo "== ncu --set full, every kernel" | tee -a "$out/log.txt"
warm=$(python babyjubjub-rs_b200/tools/profile_kernels.py --count 2>&1 | sed -n 's/.*warmup_launches=\([0-9]*\).*/\1/p' | tail -1)
echo "warm-up launches: $warm" | tee -a "$out/log.txt"
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:^k_ -s "$warm" -f -o "$out/kernels" \
    python babyjubjub-rs_b200/tools/profile_kernels.py > "$out/ncu_kernels.log" 2>&1; echo "rc=$?" | tee -a "$out/log.txt"
python babyjubjub-rs_b200/tools/ncu_summarize.py "$out/kernels.ncu-rep" "$out/r2" >> "$out/log.txt" 2>&1
rm -f "$out/kernels.ncu-rep"      # hundreds of MB; the summaries are what is kept
echo "== ncu launch list" | tee -a "$out/log.txt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-secondary --cpu-seconds 1 > "$out/launches_bench.log" 2>&1; echo "rc=$?" | tee -a "$out/log.txt"
echo "== compute-sanitizer" | tee -a "$out/log.txt"
bash babyjubjub-rs_b200/tools/sanitize.sh "$out/sanitize" 2>&1 | tee -a "$out/log.txt"
