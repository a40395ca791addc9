mkdir -p gpurun_out
P=cvxpylayers_b200
timeout 600 python -m pytest tests/test_gpu_cached.py -q > gpurun_out/tests_cached.log 2>&1; echo "cached tests rc=$?"; tail -15 gpurun_out/tests_cached.log | cut -c1-300
timeout 300 python tools/train_loop.py 2048 > gpurun_out/train_loop_r2b.json 2>gpurun_out/train_loop_r2b.err; echo "train rc=$?"; cat gpurun_out/train_loop_r2b.json | cut -c1-1200
for v in default BAR7 INV BOTH; do
  lib=$P/libbcone_$v.so; [ $v = default ] && lib=$P/libbcone.so
  BCONE_LIB=$PWD/$lib timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/v_$v.json 2>gpurun_out/v_$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/v_$v.json').read().strip().splitlines()[-1]);print('$v', round(d['value']), d['kernel_ms'], round(d['e2e']['value']))"
done
BCONE_LIB=$PWD/$P/libbcone_BOTH.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -q -x 2>&1 | tail -3
BCONE_LIB=$PWD/$P/libbcone_SUBPROF.so timeout 200 python tools/phase_profile.py > gpurun_out/phase_sub.txt 2>&1; cat gpurun_out/phase_sub.txt | head -40
