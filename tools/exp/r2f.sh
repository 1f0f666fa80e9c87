mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests.log; tail -12 gpurun_out/r2f_tests.log
echo "== new scan kernel (default)"; timeout 300 python tools/power_trace.py 100000 3.0 1 2 > gpurun_out/r2f_power_scan_new.log 2>&1; cat gpurun_out/r2f_power_scan_new.log
echo "== old query-major kernel (EVDR_TC_MODE=1cta)"; EVDR_TC_MODE=1cta timeout 300 python tools/power_trace.py 100000 3.0 1 2 > gpurun_out/r2f_power_scan_old.log 2>&1; cat gpurun_out/r2f_power_scan_old.log
