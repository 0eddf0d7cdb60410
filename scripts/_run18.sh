set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
( time timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err ) 2>&1 | tail -4
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_q.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_timing_ms'], d['full_job']['device_seconds'], d['full_job']['e2e_seconds'], d['full_job']['host_timing_ms'])"
