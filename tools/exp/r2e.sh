set -x
export TC_POWER_REPEAT=1
export TC_POWER_VARIANTS='[["default",{"EVDR_TC_MODE":""}],["pair",{"EVDR_TC_MODE":"pair"}],["pair8",{"EVDR_TC_MODE":"pair8"}]]'
timeout 300 python -m pytest tests/test_gpu_0_rows.py -q -x -k "bench_flavour" > gpurun_out/r2e_t.log 2>&1; tail -3 gpurun_out/r2e_t.log
EVDR_TC_MODE=pair8 timeout 120 python tools/tc_diag.py one 64 2000 1030 >> gpurun_out/r2e_t.log 2>&1; tail -2 gpurun_out/r2e_t.log
for nq in 16 32 1024; do echo "== nq=$nq"; timeout 200 python tools/tc_power.py 20000 $nq 2.5; done > gpurun_out/r2e_pair8.log 2>&1
cat gpurun_out/r2e_pair8.log
EVDR_TC_MODE=pair python tools/tc_launch.py 20000 1024 3 > gpurun_out/r2e_plain_pair.log 2>&1 && EVDR_TC_MODE=pair ncu --set full --clock-control none --import-source on -k regex:maxsim_tc -s 1 -c 1 -o gpurun_out/r2e_prof_pair python tools/tc_launch.py 20000 1024 3 > gpurun_out/r2e_ncu_pair.log 2>&1
EVDR_TC_MODE=pair8 python tools/tc_launch.py 20000 1024 3 > gpurun_out/r2e_plain_pair8.log 2>&1 && EVDR_TC_MODE=pair8 ncu --set full --clock-control none --import-source on -k regex:maxsim_tc -s 1 -c 1 -o gpurun_out/r2e_prof_pair8 python tools/tc_launch.py 20000 1024 3 > gpurun_out/r2e_ncu_pair8.log 2>&1
cat gpurun_out/r2e_plain_pair.log gpurun_out/r2e_plain_pair8.log
