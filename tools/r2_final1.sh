python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -2 gpurun_out/r2_bench_c4.err
python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -2 gpurun_out/r2_bench_ref.err
for c in C1 C2 C3; do python bench.py --config $c > gpurun_out/r2_cfg_$c.json 2> gpurun_out/r2_cfg_$c.err; tail -1 gpurun_out/r2_cfg_$c.err; done
python -c "
import json
for f in ['r2_bench_c4','r2_bench_ref','r2_cfg_C1','r2_cfg_C2','r2_cfg_C3']:
    d=json.load(open('gpurun_out/'+f+'.json')); e=d.get('e2e') or {}
    print(f, d.get('ms_per_step'), d.get('value'), 'e2e', e.get('ms_per_step'), e.get('value'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'), d.get('gpu_launches'))
"
