#!/bin/bash
for ppb in 8 32; do
SLVGPU_REP_PPB=$ppb ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_exp_ppb$ppb.csv \
   python bench.py --config 3 --only --no-cpu --frames 2000 --steps 2 --warmup 3 > gpurun_out/ncu_launch_exp.log 2>&1
python tools/launch_summary.py gpurun_out/launches_exp_ppb$ppb.csv > gpurun_out/launches_exp_ppb${ppb}_summary.txt
done
