python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -15 gpurun_out/r2_tests.log
TTSK_DEBUG=1 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2_b.json 2> gpurun_out/r2_b.err
grep "ttsk\]" gpurun_out/r2_b.err | sort | uniq -c | head; tail -3 gpurun_out/r2_b.err
python -c "
import json;d=json.load(open('gpurun_out/r2_b.json'));print(d['ms_per_step'],d['kernel_ms'], d['checksum'])"
