mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_default.json)"
python -c "
import json;d=json.loads(open('gpurun_out/bench_default.json').read());print(d['value'],d['e2e']['value'],d['e2e']['diagnostic_wall_ms_per_step'],d['e2e']['diagnostic_wall_ms_median'],d['kernel_paths'],d['gpu_launches'],d['clocks'])"
