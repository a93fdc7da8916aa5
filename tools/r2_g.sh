timeout 600 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_c5_1.json 2> gpurun_out/r2_c5_1.err; python -c "
import json
d=json.load(open('gpurun_out/r2_c5_1.json')); print('C5 graphs', d['ms_per_step'], d['gpu_launches'])"
TTSK_GRAPHS=0 timeout 600 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_c5_ng.json 2> gpurun_out/r2_c5_ng.err; python -c "
import json
d=json.load(open('gpurun_out/r2_c5_ng.json')); print('C5 no graphs', d['ms_per_step'], d['gpu_launches'])"
