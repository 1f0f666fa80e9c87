mkdir -p gpurun_out
export TC_POWER_REPEAT=1
export TC_POWER_VARIANTS='[["default",{"EVDR_TC_MODE":""}],["pair",{"EVDR_TC_MODE":"pair"}]]'
for nq in 20 24 28 40; do echo "== nq=$nq"; timeout 200 python tools/tc_power.py 100000 $nq 1.5; done > gpurun_out/r2ac_pair_even.log 2>&1
cat gpurun_out/r2ac_pair_even.log
