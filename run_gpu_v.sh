mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_fast_kernel -s 1 -c 1 -o gpurun_out/prof_fwdfast_final -f python tools/profile_c2.py 148 2 > gpurun_out/ncu_fwd.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fwd.log; ls -la gpurun_out/prof_fwdfast_final.ncu-rep
