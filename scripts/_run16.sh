set -x
nproc
( time timeout 2400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err ) 2>&1 | tail -4
( time timeout 2400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driver.json 2> gpurun_out/r2_bench_driver.err ) 2>&1 | tail -4
tail -c 400 gpurun_out/r2_bench_driver.err
python scripts/gpu_prof_tc.py > gpurun_out/r2_prof_tc_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_logistic_tc -s 3 -c 2 -o gpurun_out/r2_prof_tc -f python scripts/gpu_prof_tc.py > gpurun_out/r2_ncu_tc.log 2>&1
tail -3 gpurun_out/r2_ncu_tc.log
