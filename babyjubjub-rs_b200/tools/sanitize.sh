#!/bin/bash
# compute-sanitizer over the GPU parity tests of the queue / table / shared-memory kernels.
# Usage (on a GPU box):  bash babyjubjub-rs_b200/tools/sanitize.sh <out-dir>
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
sel="test_verify or test_mul_scalar or test_split or test_schnorr or test_sign or test_fixed_base or test_wide or test_empty or test_compress_decompress or test_poseidon or test_alternative"
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool" | tee "$out/$tool.txt"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -m gpu -x -q -k "$sel" >> "$out/$tool.txt" 2>&1
  echo "exit=$?" >> "$out/$tool.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit=" "$out/$tool.txt" | tail -4
done
