#!/bin/bash
# A/B over environment knobs of the library AS BUILT, on one GPU:  ab_knobs.sh "<ENV=.. ENV=..>" ["<ENV..>" ...]   ("-" = no knob)
# Per environment string: per-phase times (a synchronous run, BJJ_PHASE_TIMING), then the bench values and the per-chunk
# timeline of the host flavour from a second, undisturbed run (BJJ_PIPE_TIMING only reads events after the call).
for envs in "$@"; do
  [ "$envs" = "-" ] && envs="BJJ_NOP=1"
  echo "== $envs"
  env $envs BJJ_PHASE_TIMING=1 python bench.py --steps 2 --warmup 2 --no-secondary --cpu-seconds 0.5 2>/tmp/err.txt >/dev/null
  grep "lanes=2097152 grid" /tmp/err.txt | tail -2
  env $envs BJJ_PIPE_TIMING=1 python bench.py --steps 4 --warmup 3 --no-secondary --cpu-seconds 0.5 2>/tmp/err.txt >/tmp/out.json
  grep "bjj pipe" /tmp/err.txt | tail -3
  python -c "import json;d=json.loads(open('/tmp/out.json').read().strip().splitlines()[-1]);print('value %.4g ms %.3f e2e %.4g  %s' % (d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu']['uuid']))"
done
