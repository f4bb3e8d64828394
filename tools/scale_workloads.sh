#!/bin/bash
# Multi-GPU numbers of the secondary workloads on ONE box (BASELINE.json configs[3] / configs[4]):
#   tools/scale_workloads.sh <N>   -> gpurun_out/scale_<workload>_n{1,N}.json
# N = 1 references run concurrently on GPUs 0 / 1 (same box, same build), then the N-rank runs back to back.
N=${1:-8}
mkdir -p gpurun_out
run1() { CUDA_VISIBLE_DEVICES=$2 timeout 900 python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$1_n1.json 2> gpurun_out/scale_$1_n1.err; }
run1 moco 0 &
run1 finetune1024 1 &
wait
for wl in moco finetune1024; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
    bench.py --workload $wl --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${wl}_n$N.json 2> gpurun_out/scale_${wl}_n$N.err
  echo "$wl N=$N rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_*_n*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms', 'e2e', round(d['e2e']['value'], 1), d['clocks'])
    except Exception as e:
        print(f, 'no line', e)
PY
