BENCH_E2E_BREAKDOWN=1 B200_TRACE=1 python bench.py --steps 3 --warmup 3 --cpu-sample 0 2>&1 | grep -v "^{" | tail -20
