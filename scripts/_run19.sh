set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 236000 -c 9000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2_ncu_bench.log | cut -c1-300
wc -l gpurun_out/r2_launches_bench.csv
