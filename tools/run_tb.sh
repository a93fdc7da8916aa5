python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; tail -3 gpurun_out/r2_t2.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b2.json'));print(d['ms_per_step'],d['kernel_ms'])"; tail -3 gpurun_out/r2_b2.err
