mkdir -p gpurun_out
timeout 600 python tools/power_trace.py 100000 3.0 16 32 64 1024 > gpurun_out/r2h_power_trace_batched.log 2>&1; cat gpurun_out/r2h_power_trace_batched.log
export TC_POWER_REPEAT=1
export TC_POWER_VARIANTS='[["default",{"EVDR_TC_MODE":""}],["ew8",{"EVDR_TC_MODE":"ew8"}]]'
for nq in 12 16 24 32 48; do echo "== nq=$nq"; timeout 200 python tools/tc_power.py 100000 $nq 2.0; done > gpurun_out/r2h_ew8_small_q.log 2>&1
cat gpurun_out/r2h_ew8_small_q.log
