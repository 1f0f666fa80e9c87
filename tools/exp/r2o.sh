mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_0_rows.py tests/test_gpu_b_tc.py tests/test_gpu_g_fuzz.py tests/test_gpu_f_fullsize.py -m gpu -q -x --timeout 300 > gpurun_out/r2o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_tests.log; tail -4 gpurun_out/r2o_tests.log
export TC_POWER_REPEAT=1
export TC_POWER_VARIANTS='[["default",{"EVDR_TC_MODE":""}],["ew8",{"EVDR_TC_MODE":"ew8"}]]'
for nq in 32 64 128 256 1024; do echo "== nq=$nq"; timeout 200 python tools/tc_power.py 100000 $nq 2.5; done > gpurun_out/r2o_prefetch.log 2>&1
cat gpurun_out/r2o_prefetch.log
