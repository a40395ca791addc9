mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for rep in 1 2 3 4 5 6; do timeout 200 python bench.py --config C3 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/c3rep_$rep.json 2>gpurun_out/c3rep_$rep.err; echo "rep $rep rc=$? $(tail -1 gpurun_out/c3rep_$rep.err | cut -c1-120)"; done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2>gpurun_out/bench_r2_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_final.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms'])"
timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_sweep.py > gpurun_out/san2_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "^C[0-9]|^EXP|^sparse|RACECHECK SUMMARY|Race reported" gpurun_out/san2_racecheck.log | cut -c1-250 | head -30
