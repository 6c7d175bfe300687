#!/bin/bash
run() {
  lib=$1; cfg=$2
  SLVGPU_TAILFILL=0 SLVGPU_LIB=$lib python bench.py --config $cfg --only --no-cpu --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', $cfg, round(d['value']), {k: round(v,2) for k,v in d['detail']['kernel_ms_per_step'].items()}, d['detail']['status_max'])"
}
for lib in build_variants/*.so; do
  run $PWD/$lib 3
done
