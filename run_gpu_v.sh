mkdir -p gpurun_out
P=cvxpylayers_b200
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/tests.log | cut -c1-300
timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/v_default.json 2>gpurun_out/v_default.err
python -c "
import json;d=json.loads(open('gpurun_out/v_default.json').read().strip().splitlines()[-1]);print('default', round(d['value']), d['kernel_ms'], round(d['e2e']['value']))"
BCONE_LIB=$PWD/$P/libbcone_SUBPROF.so timeout 200 python tools/phase_profile.py > gpurun_out/phase_sub.txt 2>&1; cat gpurun_out/phase_sub.txt | head -40
timeout 200 python tools/bench_shapes.py > gpurun_out/shapes.jsonl 2>gpurun_out/shapes.err; tail -5 gpurun_out/shapes.jsonl | cut -c1-300
