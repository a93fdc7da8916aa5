python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -2 gpurun_out/r2_tests.log
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -2 gpurun_out/r2_bench_c4.err
timeout 600 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_c5_1.json 2> gpurun_out/r2_c5_1.err
python -c "
import json
for f in ['r2_bench_c4','r2_c5_1']:
    d=json.load(open('gpurun_out/'+f+'.json')); e=d.get('e2e') or {}
    print(f, d.get('ms_per_step'), d.get('value'), 'e2e', e.get('ms_per_step'), e.get('value'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'), d.get('gpu_launches'))
"
