python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_bx.json 2> gpurun_out/r2_bx.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bx.json'));print(d['ms_per_step'],d['kernel_ms'], d['checksum'])"; tail -3 gpurun_out/r2_bx.err
