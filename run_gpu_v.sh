mkdir -p gpurun_out
timeout 200 python tools/bench_shapes.py > gpurun_out/shapes.jsonl 2>gpurun_out/shapes.err; python - <<'PY'
import json
for l in open('gpurun_out/shapes.jsonl'):
    d=json.loads(l); print('  ',d['n'],d['m'],d['geometry'],d['fwd_ms_per_4096'],d['us_per_instance_iteration'])
PY
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/v_default.json 2>gpurun_out/v_default.err
python -c "
import json;d=json.loads(open('gpurun_out/v_default.json').read().strip().splitlines()[-1]);print('default', round(d['value']), d['kernel_ms'], round(d['e2e']['value']))"
timeout 300 python -m pytest tests/test_gpu_cached.py tests/test_gpu_parity.py -q -x 2>&1 | tail -2
