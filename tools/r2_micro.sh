timeout 120 tools/fp64_mix_microbench > gpurun_out/r2_fp64_mix.txt 2>&1
timeout 120 tools/ndtri_mix_microbench > gpurun_out/r2_ndtri_mix.txt 2>&1
timeout 120 tools/pipe_overlap_microbench > gpurun_out/r2_pipe_overlap.txt 2>&1
cat gpurun_out/r2_fp64_mix.txt gpurun_out/r2_ndtri_mix.txt gpurun_out/r2_pipe_overlap.txt
