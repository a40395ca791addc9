set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --cpu-sample 2048 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -3 gpurun_out/bench_r1_b.err; cat gpurun_out/bench_r1_b.json
