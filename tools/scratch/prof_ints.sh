#!/bin/bash
SLVGPU_REP_PPB=8 ncu --set full --import-source on --clock-control none -k regex:k_rep_ints -s 23 -c 5 -f -o gpurun_out/k_rep_ints_r02h \
    python bench.py --config 3 --only --no-cpu --frames 1000 --steps 1 --warmup 3 > gpurun_out/ncu_ints.log 2>&1
ls -la gpurun_out/k_rep_ints_r02h.ncu-rep
