set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in 3 5; do python bench.py --config $c > gpurun_out/bench_r01f_config$c.json 2> gpurun_out/bench_r01f_config$c.err; done
python bench.py --config 3 --frames 2000 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01f_config3.csv python bench.py --config 3 --frames 2000 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_c3.log 2>&1
