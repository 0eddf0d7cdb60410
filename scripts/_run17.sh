set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sampling.py tests/test_gpu_cabi.py -q 2>&1 | tail -6
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/r2_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['full_job']['device_seconds'], d.get('parity_checks'), d['rhat_max'], d['secondary']['value'], d['secondary']['e2e']['value'])"
