set -x
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py tests/test_gpu_sampling.py -x -q 2>&1 | tail -5
for occ in 1 2; do echo "occ $occ"; TNUTS_BLOCK_OCC=$occ timeout 300 python scripts/gpu_prof_hier.py; done
timeout 600 python scripts/run_config.py hier > gpurun_out/r2_config4_v2.json 2> gpurun_out/r2_config4_v2.err; cat gpurun_out/r2_config4_v2.json
