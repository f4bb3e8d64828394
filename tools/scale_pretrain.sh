#!/bin/bash
# pretrain (configs[1]/[2]) at N = 1 and N on ONE box: tools/scale_pretrain.sh <N> -> gpurun_out/scale_pretrain_n{1,N}.json
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_pretrain_n1.json 2> gpurun_out/scale_pretrain_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_pretrain_n$N.json 2> gpurun_out/scale_pretrain_n$N.err
echo "pretrain N=$N rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_pretrain_n*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms', 'e2e', round(d['e2e']['value'], 1), d['clocks'])
    except Exception as e:
        print(f, 'no line', e)
PY
