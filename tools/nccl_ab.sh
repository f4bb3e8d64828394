#!/bin/bash
# Same-box A/B of NCCL settings for the data-parallel step (VERDICT r1 "weak" 9: a fixed ~5 ms appears at N=2):
#   tools/nccl_ab.sh <n_gpus> "<VAR=val ...>" "<VAR=val ...>" ...      -> gpurun_out/nccl_ab_<n>.log
N=${1:-2}; shift
mkdir -p gpurun_out
: > gpurun_out/nccl_ab_$N.log
for cfg in "default" "$@"; do
  for rep in 1 2; do
    echo "== $cfg (rep $rep)" >> gpurun_out/nccl_ab_$N.log
    ( [ "$cfg" != "default" ] && export $cfg
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
        bench.py --gpus $N --steps 8 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'clocks': d['clocks']}))" ) >> gpurun_out/nccl_ab_$N.log
  done
done
cat gpurun_out/nccl_ab_$N.log
