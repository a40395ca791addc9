mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fwd_fast_kernel -s 1 -c 1 -o gpurun_out/fwdf_r1 -f python tools/profile_c2.py 296 2 > gpurun_out/ncu_fwdf.log 2>&1; tail -3 gpurun_out/ncu_fwdf.log
ls -la gpurun_out/*.ncu-rep
