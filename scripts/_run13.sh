TNUTS_PUMP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_trace3.json 2> gpurun_out/r2_bench_trace3.err
grep -v "^\[tnuts pump\] [0-9]* done" gpurun_out/r2_bench_trace3.err | grep -v "by CTA" | head -44 | cut -c1-330
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_trace3.json')); print(d['value'], d['ms_per_step'], d['full_job']['device_seconds'], d['e2e']['value'], d['roofline']['k_tc_persist'])"
