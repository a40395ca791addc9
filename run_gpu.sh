mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for sc in 2 1; do for rep in 1 2; do
BCONE_SMALL_CTA=$sc timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_C3_sc${sc}_$rep.json 2>gpurun_out/bench_C3_sc${sc}_$rep.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_C3_sc${sc}_$rep.json').read());print('C3 small=$sc rep $rep',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['solver']['solved'])" || tail -2 gpurun_out/bench_C3_sc${sc}_$rep.err
done; done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2>gpurun_out/bench_r2_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_final.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms'])"
