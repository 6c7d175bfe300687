set -x
for c in 2 3 5; do python bench.py --config $c > gpurun_out/bench_r01g_config$c.json 2> gpurun_out/bench_r01g_config$c.err; done
python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01g_config2.csv python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_pol" --launch-skip 3 -c 1 -o gpurun_out/prof_pol_r01g python bench.py --frames 1184 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_pol_full.log 2>&1
