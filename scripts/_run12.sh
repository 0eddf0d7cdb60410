TNUTS_PUMP_TRACE=1 timeout 300 python scripts/_run4.py 2>&1 | grep -v "^\[tnuts pump\] [0-9]* done" | grep "slowest\|at CTA 0\|width\|state machine phase" | tail -12 | cut -c1-600
