python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_e2e.json 2> gpurun_out/r2_e2e.err || tail -3 gpurun_out/r2_e2e.err
python -c "
import json;d=json.load(open('gpurun_out/r2_e2e.json'));e=d['e2e'];print(d['ms_per_step'], d['kernel_ms']['per_pass_last_step'], 'e2e', e['ms_per_step'], 'span', e['device_span_ms_last_step'], 'pass', e['pass_kernels_ms_last_step'], e['checksum_matches_device_arm'])"
ncu --set full --clock-control none --import-source on -k regex:"dense_first_pass" -c 1 -o gpurun_out/r2_dense_fp -f python bench.py --config C1 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_dense.log 2>&1; tail -2 gpurun_out/r2_ncu_dense.log
