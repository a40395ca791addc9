mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_aa.json 2>gpurun_out/bench_aa.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_aa.json').read());print('AA on ',d['value'],d['e2e']['value'],d['kernel_ms'],d['solver'])"
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --set acceleration_lookback=0 > gpurun_out/bench_noaa.json 2>gpurun_out/bench_noaa.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_noaa.json').read());print('AA off',d['value'],d['e2e']['value'],d['kernel_ms'],d['solver'])"
