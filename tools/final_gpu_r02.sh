#!/bin/bash
# Round-2 closing measurements on one B200 (run through gpurun from the repo root; outputs under gpurun_out/, at most
# 64 MiB per call come back: the three parts run as separate gpurun calls).   usage: tools/final_gpu_r02.sh prof|lists|bench
T=${TAG:-r02d}
set -x
case "$1" in
prof)   # ncu --set full of the dominant kernel on the headline workload and on config 2, and of the other per-frame kernels
  ncu --set full --import-source on --clock-control none -k regex:k_pol -s 3 -c 1 -f -o gpurun_out/k_pol_${T}_c3 \
      python bench.py --config 3 --only --no-cpu --frames 2220 --steps 1 --warmup 3 > gpurun_out/ncu_pol3.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:k_pol -s 3 -c 1 -f -o gpurun_out/k_pol_${T}_c2 \
      python bench.py --config 2 --only --no-cpu --frames 2220 --steps 1 --warmup 3 > gpurun_out/ncu_pol2.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:"k_rep_pair|k_frame_elect|k_disp" -s 8 -c 4 -f -o gpurun_out/k_others_${T} \
      python bench.py --config 3 --only --no-cpu --frames 1000 --steps 1 --warmup 3 > gpurun_out/ncu_others.log 2>&1
  ls -la gpurun_out/*${T}*.ncu-rep ;;
lists)  # the integral kernels of one batch and fragment type (no source, 23 launches: the reports of a call must stay under 64 MiB), k_frame_elect / k_disp, then the launch lists of the bench command
  ncu --set full --clock-control none -k regex:k_rep_ints -s 23 -c 23 -f -o gpurun_out/k_rep_ints_${T} \
      python bench.py --config 3 --only --no-cpu --frames 1000 --steps 1 --warmup 3 > gpurun_out/ncu_ints.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:"k_frame_elect|k_disp" -s 4 -c 2 -f -o gpurun_out/k_elect_disp_${T} \
      python bench.py --config 3 --only --no-cpu --frames 1000 --steps 1 --warmup 3 > gpurun_out/ncu_ed.log 2>&1
  ls -la gpurun_out/*${T}*.ncu-rep
  for c in 3 2; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c $([ $c = 3 ] && echo 700 || echo 40) --csv --log-file gpurun_out/launches_${T}_config$c.csv \
        python bench.py --config $c --only --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_launch_$c.log 2>&1
    python tools/launch_summary.py gpurun_out/launches_${T}_config$c.csv > gpurun_out/launches_${T}_config${c}_summary.txt
    head -8 gpurun_out/launches_${T}_config${c}_summary.txt
  done ;;
bench)
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err
  tail -c 600 gpurun_out/bench_${T}.json
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_reference.json 2>/dev/null
  cut -c1-300 gpurun_out/bench_${T}_reference.json
  python bench.py --config 4 --only --steps 5 --warmup 3 > gpurun_out/bench_${T}_config4.json 2>/dev/null
  cut -c1-300 gpurun_out/bench_${T}_config4.json ;;
esac
