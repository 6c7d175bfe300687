set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01h_config2.csv python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_pol" --launch-skip 3 -c 1 -o gpurun_out/prof_pol_r01h python bench.py --frames 1184 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_pol_full.log 2>&1
python bench.py --config 3 --frames 2000 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01h_config3.csv python bench.py --config 3 --frames 2000 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_c3.log 2>&1
