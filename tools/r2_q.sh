python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
for f in 1 0; do
export TTSK_NO_GA=$f
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_b.json 2> gpurun_out/r2_b.err; tail -3 gpurun_out/r2_b.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b.json'));print('no_ga=$f', d['ms_per_step'],d['kernel_ms'], d['checksum'])"
done
