set -x
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_shapes.py -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2b.log
tail -5 gpurun_out/r2_pytest2b.log
TNUTS_PUMP_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
grep -v "^\[tnuts pump\] [0-9]" gpurun_out/r2_bench2.err | tail -20
cat gpurun_out/r2_bench2.json
