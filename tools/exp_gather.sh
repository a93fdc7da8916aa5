# pass-1 (table gather) experiments on a profiling build: record prefetch on/off x L2 fetch granularity
for cfg in "0 0" "16 0" "0 32" "16 32" "0 128"; do
  set -- $cfg
  export TTSK_ABLATE=$1
  if [ "$2" != "0" ]; then export TTSK_L2_FETCH=$2; else unset TTSK_L2_FETCH; fi
  python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/exp_g.json 2> gpurun_out/exp_g.err
  echo "ablate=$1 l2fetch=$2 $(python -c "
import json;d=json.load(open('gpurun_out/exp_g.json'));print(d['kernel_ms']['per_pass_last_step'], d['checksum'])")"
done
