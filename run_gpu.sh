mkdir -p gpurun_out
B200_TRACE=1 BENCH_E2E_BREAKDOWN=1 timeout 600 python bench.py --steps 20 --warmup 3 --cpu-sample 0 2>gpurun_out/bench_trace.err > gpurun_out/bench_trace.json
grep -c "b200\] bwd" gpurun_out/bench_trace.err
grep "b200\] bwd\|\[e2e\]" gpurun_out/bench_trace.err | paste - - | awk '{print}' | cut -c1-230 | tail -24
tail -1 gpurun_out/bench_trace.err
