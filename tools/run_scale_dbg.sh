#!/bin/bash
n=$1; wl=$2; shift; shift
for ex in "$@"; do
SGPR_BENCH_ALLRANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --workload $wl --steps 20 --warmup 5 --exchange $ex 2> gpurun_out/dbg_${ex}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$ex', 'gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f ms' % d['e2e']['ms_per_step'], {k: round(v,4) for k,v in d['stages_ms_per_step'].items()}, 'parity', d['parity'].get('ok'))
"
grep "^rank" gpurun_out/dbg_${ex}.err | sort
done
