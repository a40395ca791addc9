mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2>gpurun_out/bench_r2_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_final.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms'],d['e2e']['pageable_inputs']['value'],d['cpu_baseline']['value']);f=d['e2e']['fused_params'];print({k:round(f[k]['value']) for k in ('setup_cached','setup_every_call')})"
