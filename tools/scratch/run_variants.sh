for v in ${VARIANTS:-t512_s0 rows2}; do
  for c in ${CONFIGS:-2 3}; do
    SLVGPU_LIB=$PWD/slv_b200/csrc/variants/lib_$v.so python bench.py --config $c --frames 5920 --steps 2 --warmup 2 --cpu-sample 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', $c, round(d['value']), d['config']['kernel_ms_per_step'], d['config']['status_max'])"
  done
done
