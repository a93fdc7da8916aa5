python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k "host_streaming" 2>&1 | tail -40
