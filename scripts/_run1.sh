set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
TNUTS_PUMP_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
python scripts/tc_cut_sweep.py > gpurun_out/r2_cut_sweep.log 2>&1
tail -3 gpurun_out/r2_pytest1.log; tail -c 1500 gpurun_out/r2_bench1.json; cat gpurun_out/r2_cut_sweep.log
