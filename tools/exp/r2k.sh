mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.log; tail -4 gpurun_out/r2k_tests.log
export TC_POWER_REPEAT=1
export TC_POWER_VARIANTS='[["1cta-family",{"EVDR_TC_MODE":"ew8"}],["1cta",{"EVDR_TC_MODE":"1cta"}],["pair8",{"EVDR_TC_MODE":"pair8"}]]'
for nq in 9 12 20 40 48 100; do echo "== nq=$nq"; timeout 200 python tools/tc_power.py 100000 $nq 1.5; done > gpurun_out/r2k_sweep.log 2>&1
cat gpurun_out/r2k_sweep.log
python tools/tc_launch.py 100000 1024 3 > gpurun_out/r2k_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:maxsim_tc -s 1 -c 1 -o gpurun_out/r2k_prof_pair8_1024x100k python tools/tc_launch.py 100000 1024 3 > gpurun_out/r2k_ncu.log 2>&1
cat gpurun_out/r2k_plain.log
