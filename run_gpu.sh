python -m pytest tests -x -q -m gpu 2>&1 | tail -5
BENCH_E2E_BREAKDOWN=1 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.err; grep e2e gpurun_out/bench_r1_i.err | tail -4; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_i.json')); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e'])"
