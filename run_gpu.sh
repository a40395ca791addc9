mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/tests.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2>gpurun_out/bench_r2_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_final.json').read());print('C2',d['value'],d['e2e']['value'],d['kernel_ms'],d['solver'],d['cpu_baseline']['value'],d['roofline']['frac'])"
for c in C3 C5 EXP C1 C4; do
timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_r2_$c.json 2>gpurun_out/bench_r2_$c.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r2_$c.json').read());print('$c',round(d['value']),round(d['e2e']['value']),d['kernel_ms'],d['kernel_geometry']['fwd_ctas_per_sm'],d.get('cpu_baseline',{}).get('value'))" || tail -3 gpurun_out/bench_r2_$c.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches_r2.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_fast_kernel -s 1 -c 1 -o gpurun_out/fwdfast_r2 python tools/profile_c2.py 296 2 > gpurun_out/ncu_fwdf_r2.log 2>&1; tail -2 gpurun_out/ncu_fwdf_r2.log
