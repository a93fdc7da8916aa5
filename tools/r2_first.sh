# round-2 first GPU pass: tests, C4 bench, other configs, launch list
set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -5 gpurun_out/r2_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -3 gpurun_out/r2_bench_c4.err; cat gpurun_out/r2_bench_c4.json
bash tools/run_cfg.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu.log 2>&1; tail -2 gpurun_out/r2_ncu.log
