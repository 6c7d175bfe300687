#!/bin/bash
# Round-2 closing measurements on one B200 (run through gpurun from the repo root; outputs under gpurun_out/).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
tail -c 600 gpurun_out/bench_r02_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_reference.json 2>/dev/null
cat gpurun_out/bench_r02_final_reference.json | cut -c1-300
python bench.py --config 4 --only --steps 5 --warmup 3 > gpurun_out/bench_r02_final_config4.json 2>/dev/null
for c in 3 2 5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02_config$c.csv \
      python bench.py --config $c --only --no-cpu --frames 2000 --steps 2 --warmup 3 > gpurun_out/ncu_launch_$c.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_r02_config$c.csv > gpurun_out/launches_r02_config${c}_summary.txt
  head -8 gpurun_out/launches_r02_config${c}_summary.txt
done
ncu --set full --import-source on --clock-control none -k regex:k_pol -s 3 -c 1 -f -o gpurun_out/k_pol_r02 \
    python bench.py --config 2 --only --no-cpu --frames 1184 --steps 1 --warmup 3 > gpurun_out/ncu_pol.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_rep_pair|k_frame_elect|k_disp" -s 12 -c 6 -f -o gpurun_out/k_rep_pair_r02 \
    python bench.py --config 3 --only --no-cpu --frames 1000 --steps 1 --warmup 3 > gpurun_out/ncu_rep.log 2>&1
ls -la gpurun_out/*.ncu-rep
