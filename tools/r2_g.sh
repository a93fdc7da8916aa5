python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_g2.json 2> gpurun_out/r2_g2.err; python -c "
import json;d=json.load(open('gpurun_out/r2_g2.json'));print(d['ms_per_step'], d['kernel_ms'], d['checksum'])"
