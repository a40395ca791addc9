mkdir -p gpurun_out
./tools/microbench 2>&1 | tee gpurun_out/micro.log
