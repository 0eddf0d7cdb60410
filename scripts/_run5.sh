set -x
TNUTS_PUMP_TRACE=1 timeout 600 python scripts/_run4.py 2>&1 | grep "persist\]"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -25 gpurun_out/r2_pytest5.log
