mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 600 python tools/bench_configs.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], d['fwd_ms'], d['bwd_ms'], d['problems_per_s'], d['solved'])"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_last.json 2>gpurun_out/bench_last.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_last.json').read());print(d['value'],d['e2e']['value'],d['e2e']['diagnostic_wall_ms_per_step'],d['kernel_ms'])"
