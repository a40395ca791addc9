mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -q -x 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_v.json 2>gpurun_out/bench_v.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms']);print(d['e2e']['pageable_inputs'])"
tail -2 gpurun_out/bench_v.err
