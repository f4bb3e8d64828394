#!/bin/bash
# Same-box, interleaved A/B of whole-step time for CMU_DEBUG_KNOBS settings (N GPUs):
#   tools/ab_knobs.sh "<knobs A>" "<knobs B>" [reps] [n_gpus] [extra bench args]   -> gpurun_out/ab_knobs_<n>.log
A="$1"; B="$2"; REPS=${3:-2}; N=${4:-1}; shift 4
mkdir -p gpurun_out
LOG=gpurun_out/ab_knobs_$N.log
: > $LOG
for rep in $(seq 1 $REPS); do
  for cfg in "$A" "$B"; do
    echo "== knobs='$cfg' rep $rep" >> $LOG
    if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200))"; fi
    CMU_DEBUG_KNOBS="$cfg" CMU_BENCH_VERBOSE=1 timeout 600 $LAUNCH bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/ab_knobs_last.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(json.dumps({'ms_per_step': round(d['ms_per_step'], 2), 'e2e_ms': round(d['e2e']['ms_per_step'], 2), 'k1_tflops': round(r['achieved'], 1), 'sm_mhz': d['clocks']['sm_mhz'],
                          'kern': {k: v['tflops'] for k, v in r['all_tensor_core_kernels'].items()}}))" >> $LOG
    echo "rc=${PIPESTATUS[0]}" >> $LOG
    grep -E "cmu:|launch failure|illegal|Traceback" gpurun_out/ab_knobs_last.err | head -3 >> $LOG; grep "^\[bench" gpurun_out/ab_knobs_last.err | tail -2 >> $LOG
  done
done
cat $LOG
