set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; tail -3 gpurun_out/bench_r1_e.err; cat gpurun_out/bench_r1_e.json
ncu --set full --clock-control none --import-source on -k regex:bwd_fast -s 1 -c 1 -o gpurun_out/prof_bwd_r1d python tools/profile_c2.py 296 2 > gpurun_out/prof_bwd.log 2>&1
