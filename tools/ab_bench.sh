#!/bin/bash
# A/B bench of library builds: tools/ab_bench.sh <workload> <lib1> <lib2> ...   ("default" = the in-tree library)
wl=$1; shift
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset SGPR_B200_LIB; else export SGPR_B200_LIB=$PWD/$lib; fi
  python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'], {k: round(v,4) for k,v in d['stages_ms_per_step'].items()}, 'parity', d['parity'].get('ok'), d['parity'].get('error'))
"
done
