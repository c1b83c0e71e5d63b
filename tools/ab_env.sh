#!/bin/bash
# A/B bench under environment settings: tools/ab_env.sh <workload> "VAR=val" "VAR2=val" ...   ("-" = none)
wl=$1; shift
for kv in "$@"; do
  if [ "$kv" = "-" ]; then envs=""; else envs="$kv"; fi
  env $envs python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[$kv]', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'], {k: round(v,4) for k,v in d['stages_ms_per_step'].items()}, 'parity', d['parity'].get('ok'), d['parity'].get('error'))
"
done
