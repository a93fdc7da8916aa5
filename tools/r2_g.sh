python tools/e2e_passes.py > gpurun_out/r2_e2e_chunks.txt 2>&1; tail -12 gpurun_out/r2_e2e_chunks.txt | cut -c1-300
