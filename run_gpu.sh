mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
for c in C5 EXP C3 C1; do timeout 200 python tools/phase_generic.py $c > gpurun_out/phase_$c.log 2>&1; cat gpurun_out/phase_$c.log; done
timeout 600 python bench.py --config C4 --steps 2 --warmup 1 > gpurun_out/bench_C4.json 2>gpurun_out/bench_C4.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_C4.json').read());print('C4',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver'],d.get('cpu_baseline',{}))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -c 1 -o gpurun_out/c5_fwd_r2 python tools/phase_generic.py C5 > gpurun_out/ncu_c5.log 2>&1; tail -3 gpurun_out/ncu_c5.log
