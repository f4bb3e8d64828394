for rep in 1 2; do
  for cfg in "CMU_X=0" "CMU_NO_MASK_PREFETCH=1"; do
    echo "== $cfg rep $rep"
    env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"
  done
done
