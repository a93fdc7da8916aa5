# usage: bash tools/r2_c5.sh N   (C5: blocked_stream_sketch of TensorSum(100 TT + sparse), 1.25e8 nonzeros per GPU)
N=$1
if [ "$N" = "1" ]; then
  timeout 500 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_c5_$N.json 2> gpurun_out/r2_c5_$N.err
else
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --config C5 --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_c5_$N.json 2> gpurun_out/r2_c5_$N.err
fi
echo "rc=$?"; tail -c 400 gpurun_out/r2_c5_$N.err
python -c "
import json
for ln in open('gpurun_out/r2_c5_$N.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('C5 N=$N', d['ms_per_step'], d['value'], d['unit'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'launches', d['gpu_launches'])"
