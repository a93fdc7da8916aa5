python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -15 gpurun_out/r2_tests.log
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_b.json 2> gpurun_out/r2_b.err; tail -3 gpurun_out/r2_b.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b.json'));print(d['ms_per_step'],d['kernel_ms'], d['checksum'])"
python bench.py --config C1 --steps 10 --warmup 3 > gpurun_out/r2_c1.json 2> gpurun_out/r2_c1.err || tail -5 gpurun_out/r2_c1.err
python -c "
import json;d=json.load(open('gpurun_out/r2_c1.json'));print('C1', d['ms_per_step'], d['value'], d['unit'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'cpu', d.get('cpu_baseline',{}).get('ms'))"
