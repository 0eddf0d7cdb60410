set -x
timeout 120 python scripts/gpu_prof_persist.py > gpurun_out/r2_prof_persist_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_persist -c 1 -o gpurun_out/r2_prof_persist -f python scripts/gpu_prof_persist.py > gpurun_out/r2_ncu_persist.log 2>&1
cat gpurun_out/r2_prof_persist_plain.log; tail -4 gpurun_out/r2_ncu_persist.log
