#!/bin/bash
python -m pytest tests -m gpu -x -q -k "pol or golden or cluster or clash or dipind or forces or energy or full or large" 2>&1 | tail -2
run() {
  cfg=$1; shift
  env "$@" python bench.py --config $cfg --only --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', $cfg, round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['detail']['kernel_ms_per_step'].items()}, d['detail']['status_max'], round(d['roofline']['frac'],4))"
}
run 3 SLVGPU_TAILFILL=0
run 2 A=1
