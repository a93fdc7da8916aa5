python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for c in C1 C2 C3; do python bench.py --config $c > gpurun_out/r2_cfg_$c.json 2> gpurun_out/r2_cfg_$c.err; tail -2 gpurun_out/r2_cfg_$c.err; done
python -c "
import json
for f in ['r2_cfg_C1','r2_cfg_C2','r2_cfg_C3']:
    d=json.load(open('gpurun_out/'+f+'.json')); e=d.get('e2e') or {}
    print(f, d.get('ms_per_step'), d.get('value'), 'e2e', e.get('ms_per_step'), d.get('gpu_launches'), (d.get('roofline') or {}).get('frac'))
"
