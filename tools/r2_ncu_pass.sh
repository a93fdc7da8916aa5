python -m pytest tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"sparse_(pass|sg)_kernel|gw_kernel" -c 4 -o gpurun_out/r2_pass_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_ncu_full.log 2>&1; tail -3 gpurun_out/r2_ncu_full.log
