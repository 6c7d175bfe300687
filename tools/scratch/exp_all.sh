#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() {
  cfg=$1; shift
  env "$@" python bench.py --config $cfg --only --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', $cfg, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['detail']['kernel_ms_per_step'].items()}, d['detail']['status_max'])"
}
run 3 SLVGPU_TAILFILL=0
run 3 A=1
run 5 A=1
run 4 A=1
