mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/tests.log | tail -30
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_aa.json 2>gpurun_out/bench_aa.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_aa.json').read());print('AA on ',d['value'],d['e2e']['value'],d['kernel_ms'],d['solver'])"
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --set acceleration_lookback=0 > gpurun_out/bench_noaa.json 2>gpurun_out/bench_noaa.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_noaa.json').read());print('AA off',d['value'],d['e2e']['value'],d['kernel_ms'],d['solver'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_ref.json').read());print('REF',d['value'],d['cpu_baseline'])"
for c in C5 EXP C3; do
timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_$c.json 2>gpurun_out/bench_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_$c.json').read());print('$c',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver'],d.get('cpu_baseline',{}).get('value'))"
done
timeout 300 python tools/phase_profile.py > gpurun_out/phase.log 2>&1; tail -30 gpurun_out/phase.log
