timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_sampling.py -x -q 2>&1 | tail -5
TNUTS_PUMP_TRACE=1 timeout 300 python scripts/_run4.py 2>&1 | grep -v "^\[tnuts pump\] [0-9]* done" | grep "at CTA 0\|width\|inside the state\|device s" | cut -c1-330
