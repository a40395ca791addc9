mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2>gpurun_out/bench_r2_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_final.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms'])"
timeout 300 python tools/train_loop.py 2048 > gpurun_out/train_loop_r2.json 2>gpurun_out/train_loop_r2.err; cut -c1-900 gpurun_out/train_loop_r2.json
timeout 200 python bench.py --config C3 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/c3.json 2>gpurun_out/c3.err; echo "C3 rc=$?"
