mkdir -p gpurun_out
./tools/microbench 2>&1 | tail -14 | tee gpurun_out/micro.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests.log
timeout 300 python tools/phase_profile.py 592 2>&1 | grep -v "^  \[" | tee gpurun_out/phase.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
