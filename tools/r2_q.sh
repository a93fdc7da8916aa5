python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py tests/test_gpu_linalg.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_b.json 2> gpurun_out/r2_b.err; tail -3 gpurun_out/r2_b.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b.json'));print(d['ms_per_step'],d['kernel_ms'], d['checksum'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu.log 2>&1
grep -E "partition_kernel|hist_local|cta_base" gpurun_out/r2_launches.csv | tail -5 | cut -c1-60,200-
python bench.py --config C2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c2.json 2> gpurun_out/r2_c2.err || tail -5 gpurun_out/r2_c2.err
python -c "
import json;d=json.load(open('gpurun_out/r2_c2.json'));print('C2', d['ms_per_step'], d['value'], d['unit'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
