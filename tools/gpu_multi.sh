#!/bin/bash
# multi-GPU bench lines on one box: weak scaling of the headline (config 2) and strong scaling of the sharded config 3.
# usage: tools/gpu_multi.sh <tag> "<N list, e.g. 1 2 4 8>" [configs, default "2 3"]
TAG=${1:-multi}
mkdir -p gpurun_out
for cfg in ${3:-2 3}; do
for n in $2; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --config $cfg --no-cpu-baseline > gpurun_out/${TAG}_cfg${cfg}_n$n.json 2> gpurun_out/${TAG}_cfg${cfg}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 --config $cfg --no-cpu-baseline > gpurun_out/${TAG}_cfg${cfg}_n$n.json 2> gpurun_out/${TAG}_cfg${cfg}_n$n.err
  fi
  tail -1 gpurun_out/${TAG}_cfg${cfg}_n$n.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('cfg $cfg N=$n', d['scaling'], 'value %.1f G tris/s' % (d['value']/1e9), 'ms/step %.4f' % d['ms_per_step'], 'median %.4f' % d['frame_ms']['median'], 'p99 %.4f' % d['frame_ms']['p99'], 'max %.4f' % d['frame_ms']['max'], 'e2e ms %.4f' % d['e2e']['ms_per_step'], 'per rank', d['frame_ms']['sum_per_rank'], d['config'].get('shards'))
except Exception as e:
    print('cfg $cfg N=$n FAILED', e)
"
  tail -3 gpurun_out/${TAG}_cfg${cfg}_n$n.err
done
done
