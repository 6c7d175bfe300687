#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:k_pol -s 2 -c 1 -f -o gpurun_out/k_pol_r02b \
    python bench.py --config 3 --only --no-cpu --frames 2220 --steps 1 --warmup 3 > gpurun_out/ncu_pol.log 2>&1
ls -la gpurun_out/k_pol_r02b.ncu-rep
