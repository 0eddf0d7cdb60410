set -x
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err ) 2>&1 | tail -4
tail -c 400 gpurun_out/r2_bench_4gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_4gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['full_job']['device_seconds'], {k:v.get('ok') for k,v in d['parity_checks'].items()}, d['rhat_max'], d['secondary']['value'], d['secondary']['e2e']['value'])"
