mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
for c in C5 EXP; do timeout 200 python tools/phase_generic.py $c > gpurun_out/phase_$c.log 2>&1; cat gpurun_out/phase_$c.log; done
timeout 300 python tools/train_loop.py > gpurun_out/train_loop.json 2>gpurun_out/train_loop.err; cat gpurun_out/train_loop.json
timeout 600 python bench.py --config C2SOC --steps 2 --warmup 1 --cpu-sample 64 > gpurun_out/bench_C2SOC.json 2>gpurun_out/bench_C2SOC.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_C2SOC.json').read());print('C2SOC',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver'],d.get('cpu_baseline',{}).get('value'))"
for c in C5 EXP; do
timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_$c.json 2>gpurun_out/bench_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_$c.json').read());print('$c',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver'],d.get('cpu_baseline',{}).get('value'))"
done
