mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
timeout 900 python bench.py --config C4 --steps 2 --warmup 1 > gpurun_out/bench_C4.json 2>gpurun_out/bench_C4.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_C4.json').read());print('C4',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver'],d['kernel_paths'],d.get('cpu_baseline',{}).get('value'))"
timeout 600 python bench.py --config C2SOC --steps 2 --warmup 1 --cpu-sample 64 > gpurun_out/bench_C2SOC.json 2>gpurun_out/bench_C2SOC.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_C2SOC.json').read());print('C2SOC',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver'],d['kernel_paths'],d.get('cpu_baseline',{}).get('value'))"
timeout 300 python tools/bench_shapes.py > gpurun_out/shapes.jsonl 2>gpurun_out/shapes.err; cat gpurun_out/shapes.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2>gpurun_out/bench_c2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_c2.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms'],d['solver'],d['cpu_baseline'])"
