python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for st in 8 6 4 3; do TTSK_GT_STAGES=$st python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_g2.json 2> gpurun_out/r2_g2.err; python -c "
import json;d=json.load(open('gpurun_out/r2_g2.json'));print('stages $st', d['ms_per_step'], d['kernel_ms'], d['checksum'])"; done
