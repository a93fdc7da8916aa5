# L2 fetch granularity vs per-pass times (cache-policy knob; results unchanged)
for g in 0 32 64 128; do
  if [ "$g" != "0" ]; then export TTSK_L2_FETCH=$g; else unset TTSK_L2_FETCH; fi
  python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/exp_g.json 2> gpurun_out/exp_g.err
  echo "l2fetch=$g $(python -c "
import json;d=json.load(open('gpurun_out/exp_g.json'));print(d['ms_per_step'], d['kernel_ms']['per_pass_last_step'], d['checksum'])")"
done
