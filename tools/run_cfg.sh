python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in C1 C2 C3; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_cfg_$c.json 2> gpurun_out/r2_cfg_$c.err || tail -5 gpurun_out/r2_cfg_$c.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_cfg_$c.json'));print('$c', d['ms_per_step'], d['value'], d['unit'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'cpu', d.get('cpu_baseline',{}).get('ms'))"
done
python bench.py --config C5 --nnz 2e7 --steps 2 --warmup 1 > gpurun_out/r2_cfg_C5s.json 2> gpurun_out/r2_cfg_C5s.err || tail -5 gpurun_out/r2_cfg_C5s.err
python -c "
import json;d=json.load(open('gpurun_out/r2_cfg_C5s.json'));print('C5 2e7', d['ms_per_step'], d['value'], d['unit'], 'e2e', d['e2e']['ms_per_step'], 'cpu', d.get('cpu_baseline',{}).get('value'))"
