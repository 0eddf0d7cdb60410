set -x
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py tests/test_gpu_sampling.py tests/test_gpu_multi.py tests/test_gpu_tc.py -x -q 2>&1 | tail -5
timeout 300 python scripts/_run4.py 2>&1 | grep "device s"
timeout 600 python scripts/straggler_stats.py > gpurun_out/r2_straggler.log 2>&1; cat gpurun_out/r2_straggler.log
for occ in 1 2; do echo "occ $occ"; TNUTS_BLOCK_OCC=$occ timeout 300 python scripts/gpu_prof_hier.py; done
timeout 600 python scripts/run_config.py hier > gpurun_out/r2_config4_v2.json 2> gpurun_out/r2_config4_v2.err; cat gpurun_out/r2_config4_v2.json
( time timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_driver.json 2> gpurun_out/r2_bench_driver.err ) 2>&1 | tail -4
tail -c 1500 gpurun_out/r2_bench_driver.err
