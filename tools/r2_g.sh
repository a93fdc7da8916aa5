python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
for n in 1e8 1.25e7; do python bench.py --nnz $n --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_g2.json 2> gpurun_out/r2_g2.err; python -c "
import json;d=json.load(open('gpurun_out/r2_g2.json'));print('nnz $n', d['ms_per_step'], d['kernel_ms'], d['checksum'])"; done
python tools/e2e_passes.py 2>&1 | grep "host call\|device call" | cut -c1-200
