python bench.py --steps 5 --warmup 3 --cpu-sample 2048 > gpurun_out/bench_r1_k.json 2> gpurun_out/bench_r1_k.err; grep "bench\]" gpurun_out/bench_r1_k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_k.json')); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 2 --warmup 1 --cpu-sample 2048 > gpurun_out/bench_r1_ref.json 2>/dev/null; cat gpurun_out/bench_r1_ref.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/bench_under_ncu.log 2>&1
grep -c fwd_kernel gpurun_out/launches_r1.csv
