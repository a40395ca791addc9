mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
for sc in 1 0; do for c in C3 C5 EXP C1; do
BCONE_SMALL_CTA=$sc timeout 300 python bench.py --config $c --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/bench_${c}_sc$sc.json 2>gpurun_out/bench_${c}_sc$sc.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_${c}_sc$sc.json').read());print('$c small=$sc',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['kernel_geometry'])"
done; done
timeout 300 python tools/active_frac_sweep.py > gpurun_out/active_frac.jsonl 2>gpurun_out/active_frac.err; cat gpurun_out/active_frac.jsonl
