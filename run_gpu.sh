python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/bench_configs.py 2>&1 | tee gpurun_out/configs_r1.jsonl | tail -6
