mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_sweep.py C2 > gpurun_out/san3_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "^C2|RACECHECK SUMMARY|Race reported" gpurun_out/san3_racecheck.log | cut -c1-250 | head
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_sweep.py C2 > gpurun_out/san3_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "^C2|ERROR SUMMARY|Invalid" gpurun_out/san3_memcheck.log | cut -c1-250 | head
timeout 300 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_sweep.py C2 > gpurun_out/san3_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "^C2|ERROR SUMMARY|Uninitialized" gpurun_out/san3_initcheck.log | cut -c1-250 | head
