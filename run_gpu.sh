mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2>gpurun_out/bench_r2_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_final.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms']);f=d['e2e']['fused_params'];print({k:(round(f[k]['value']),f[k]['one_step_wall_ms']) for k in ('setup_cached','setup_every_call')})"
timeout 200 python tools/bench_shapes.py > gpurun_out/shapes.jsonl 2>gpurun_out/shapes.err; python - <<'PY'
import json
for l in open('gpurun_out/shapes.jsonl'):
    d=json.loads(l); print('  ',d['n'],d['m'],d['geometry'],d['fwd_ms_per_4096'],d['us_per_instance_iteration'])
PY
timeout 300 python tools/train_loop.py 2048 > gpurun_out/train_loop_r2.json 2>gpurun_out/train_loop_r2.err; python -c "
import json;d=json.loads(open('gpurun_out/train_loop_r2.json').read());print({k:round(d[k]['ms_per_step_mean'],2) for k in ('cold','warm','warm_cached')})"
