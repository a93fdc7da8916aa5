python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -4 gpurun_out/r2_tests.log
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -3 gpurun_out/r2_bench_c4.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_c4.json'));e=d['e2e'];print(d['ms_per_step'], d['kernel_ms'], 'e2e', e['ms_per_step'], 'span', e['device_span_ms_last_step'], 'pass', e['pass_kernels_ms_last_step'], e['checksum_matches_device_arm'], d['roofline']['frac'], d['fp64_pipe'], d['cpu_baseline']['value'])"
python tools/e2e_passes.py > gpurun_out/r2_e2e_chunks.txt 2>&1; tail -12 gpurun_out/r2_e2e_chunks.txt
