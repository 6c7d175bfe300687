#!/bin/bash
python -m pytest tests -m gpu -x -q -k "rep or repulsion or full or batch" 2>&1 | tail -3
run() {
  env "$@" python bench.py --config 3 --only --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']), {k: round(v,2) for k,v in d['detail']['kernel_ms_per_step'].items()}, d['detail']['status_max'])"
}
run SLVGPU_REP_PPB=4
run SLVGPU_REP_PPB=8
run SLVGPU_REP_PPB=16
run SLVGPU_REP_PPB=32
run SLVGPU_REP_BT=64 SLVGPU_REP_PPB=8
