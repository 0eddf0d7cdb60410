set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "unit_metric" 2>&1 | tail -60
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampling.py tests/test_gpu_multi.py tests/test_gpu_tc.py tests/test_gpu_cabi.py -q 2>&1 | tail -15
TNUTS_PUMP_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_trace.json 2> gpurun_out/r2_bench_trace.err
grep "persist\]" gpurun_out/r2_bench_trace.err | tail -12
