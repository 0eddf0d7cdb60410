TNUTS_PUMP_TRACE=1 timeout 300 python scripts/_run4.py 2>&1 | grep "persist\]\|device s"
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_shapes.py -x -q 2>&1 | tail -3
