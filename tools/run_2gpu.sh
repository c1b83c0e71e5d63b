T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/p2p_check.py c2 2>&1 | grep -v Warn | tail -6
timeout 300 $T tools/p2p_check.py c3 2>&1 | grep "step 7\|FAIL\|Error" | tail -4
timeout 300 $T tools/p2p_check.py c3 nccl 2>&1 | grep "step 7\|FAIL\|Error" | tail -4
for ex in p2p p2p-nccl; do
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$ex', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'], {k: round(v,4) for k,v in d['stages_ms_per_step'].items()}, 'parity', d['parity'].get('ok'), d['parity'].get('error'), d['parity'].get('sharded_vs_unsharded_full_size'))
"
done
