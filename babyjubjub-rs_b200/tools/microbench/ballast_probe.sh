#!/bin/bash
# Compile-only (no GPU): passengers on the fma pipe inside the mul2 / mul subroutines for three callers.  See ballast_probe.cu.
here=$(cd "$(dirname "$0")" && pwd)
tmp=$(mktemp -d)
for v in 0 1 2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 20091 -diag-suppress 177 -DBJJ_BALLAST=0 -DVARIANT=$v \
       -I "$here/../../csrc" -cubin -o "$tmp/p$v.cubin" "$here/ballast_probe.cu" || exit 1
  echo "== VARIANT $v"
  cuobjdump -sass "$tmp/p$v.cubin" | awk '
    /\/\*[0-9a-f]+\*\// {
      op = ""
      for (i = 1; i <= NF; i++) if ($i ~ /^[A-Z][A-Z0-9_.]+$/ && $i !~ /^U?P[0-9]$/) { op = $i; break }
      if (op ~ /^IMAD.WIDE/) w++; else if (op ~ /^(IMAD|HFMA2)/) { p++; pk[op]++ }
      tot++
      if (op ~ /^RET/ || op ~ /^EXIT/) {
        if (w >= 128) { printf "  subroutine with %d IMAD.WIDE: %d instructions, %d other fma-pipe instructions:", w, tot, p; for (o in pk) printf " %s=%d", o, pk[o]; printf "\n" }
        tot = w = p = 0; delete pk
      }
    }'
done
rm -rf "$tmp"
