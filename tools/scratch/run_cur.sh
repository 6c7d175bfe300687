for c in ${CONFIGS:-2 3}; do
  python bench.py --config $c --frames ${FRAMES:-5920} --steps 2 --warmup 2 --cpu-sample 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($c, round(d['value']), d['config']['kernel_ms_per_step'], d['config']['status_max'], d['config']['mean_pol_iters'])"
done
