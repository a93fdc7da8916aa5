for s in 16777216 25000000 33554432 50000000; do echo "stage $s"; TTSK_STAGE_NNZ=$s python tools/e2e_passes.py 2>&1 | grep "host call" | cut -c1-200; done
