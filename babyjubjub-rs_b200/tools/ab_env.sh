#!/bin/bash
# A/B over environment knobs of ONE build, on the same GPU:  ab_env.sh <variant> "<ENV=.. ENV=..>" ["<ENV..>" ...]
# Swaps ab_variants/<variant>/libbjj_cuda.so in, runs bench.py once per environment string, restores the library.
variant=$1; shift
cp babyjubjub-rs_b200/libbjj_cuda.so /tmp/keep.so
cp ab_variants/$variant/libbjj_cuda.so babyjubjub-rs_b200/libbjj_cuda.so
for envs in "$@"; do
  echo "== $variant $envs"
  env $envs BJJ_PHASE_TIMING=1 python bench.py --steps 3 --warmup 3 --no-secondary --cpu-seconds 1 2>&1 >/tmp/out.json | grep "lanes=2097152" | tail -2
  python -c "import json;d=json.loads(open('/tmp/out.json').read().strip().splitlines()[-1]);print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
cp /tmp/keep.so babyjubjub-rs_b200/libbjj_cuda.so
