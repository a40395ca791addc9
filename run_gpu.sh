python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --cpu-sample 2048 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -3 gpurun_out/bench_r1_g.err; cat gpurun_out/bench_r1_g.json
