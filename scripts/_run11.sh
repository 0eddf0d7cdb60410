set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampling.py tests/test_gpu_multi.py tests/test_gpu_cabi.py tests/test_gpu_shapes.py -q 2>&1 | tail -8
TNUTS_PUMP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_trace2.json 2> gpurun_out/r2_bench_trace2.err
grep -v "^\[tnuts pump\] [0-9]* done" gpurun_out/r2_bench_trace2.err | head -60
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_trace2.json')); print(d['value'], d['ms_per_step'], d['full_job']['device_seconds'], d['e2e']['value'], d['roofline']['k_tc_persist'])"
