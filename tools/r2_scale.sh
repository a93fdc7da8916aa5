# usage: bash tools/r2_scale.sh N   (C4 strong scaling line at N GPUs, device-resident + e2e)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
echo "rc=$?"; tail -c 600 gpurun_out/r2_scale_$N.err; ls -la gpurun_out/r2_scale_$N.json
python -c "
import json;d=json.load(open('gpurun_out/r2_scale_$N.json'));e=d['e2e'];print('N=$N', d['ms_per_step'], d['kernel_ms']['per_pass_last_step'], 'e2e', e['ms_per_step'], 'span', e['device_span_ms_last_step'], e['checksum_matches_device_arm'], d['checksum'])"
