#!/bin/bash
# tools/run_scale.sh <n_gpus> [workload ...]: bench.py under torchrun, one JSON per workload into gpurun_out/
n=$1; shift
for wl in "$@"; do
  if [ "$n" = "1" ]; then
    python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_${wl}_${n}gpu.json 2> gpurun_out/r02_scale_${wl}_${n}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload $wl --steps 20 --warmup 5 > gpurun_out/r02_scale_${wl}_${n}gpu.json 2> gpurun_out/r02_scale_${wl}_${n}gpu.err
  fi
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_scale_${wl}_${n}gpu.json').read().strip().splitlines()[-1])
print('$wl', 'gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4f ms' % d['e2e']['ms_per_step'], {k: round(v,4) for k,v in d['stages_ms_per_step'].items()}, 'parity', d['parity'].get('ok'), d['parity'].get('error'), d['parity'].get('sharded_vs_unsharded_full_size'))
" || tail -5 gpurun_out/r02_scale_${wl}_${n}gpu.err
done
