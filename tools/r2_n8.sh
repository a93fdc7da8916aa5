N=$1
bash tools/r2_scale.sh $N
bash tools/r2_c5.sh $N
python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -2
