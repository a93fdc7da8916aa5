python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -15 gpurun_out/r2_tests.log
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_e2e.json 2> gpurun_out/r2_e2e.err || tail -3 gpurun_out/r2_e2e.err
python -c "
import json;d=json.load(open('gpurun_out/r2_e2e.json'));e=d['e2e'];print(d['ms_per_step'], d['kernel_ms'], 'e2e', e['ms_per_step'], 'span', e['device_span_ms_last_step'], 'pass', e['pass_kernels_ms_last_step'], e['checksum_matches_device_arm'])"
python bench.py --config C1 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_c1.json 2> gpurun_out/r2_c1.err || tail -5 gpurun_out/r2_c1.err
python -c "
import json;d=json.load(open('gpurun_out/r2_c1.json'));print('C1', d['ms_per_step'], d['value'], d['unit'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_c1_launches.csv python bench.py --config C1 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_c1ncu.log 2>&1
