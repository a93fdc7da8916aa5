# launch list (kernel shares of one C4 step) and the --set full capture of the four pass kernels; only after the plain run exited 0
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu.log 2>&1; tail -2 gpurun_out/r2_ncu.log
ncu --set full --clock-control none --import-source on -k regex:"sparse_(pass|sg)_kernel|gw_kernel|gt_kernel" -c 4 -o gpurun_out/r2_pass_full -f python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_ncu_full.log 2>&1; tail -3 gpurun_out/r2_ncu_full.log
