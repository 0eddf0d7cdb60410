set -x
for m in 0 1 2 3; do echo "bar mode $m"; TNUTS_BAR_MODE=$m TNUTS_PUMP_TRACE=1 timeout 300 python scripts/_run4.py 2>&1 | grep "device s"; done
timeout 600 python scripts/straggler_stats.py > gpurun_out/r2_straggler.log 2>&1; cat gpurun_out/r2_straggler.log
nproc; free -g | head -2
( time timeout 2400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err ) 2>&1 | tail -4
( time timeout 2400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driver.json 2> gpurun_out/r2_bench_driver.err ) 2>&1 | tail -4
tail -c 3000 gpurun_out/r2_ref_arm.json; tail -c 600 gpurun_out/r2_bench_driver.err
