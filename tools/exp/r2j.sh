mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_0_rows.py tests/test_gpu_b_tc.py -q -x -k "bench_flavour or query_tile_configs" > gpurun_out/r2j_t.log 2>&1; tail -3 gpurun_out/r2j_t.log
export TC_POWER_REPEAT=1
export TC_POWER_VARIANTS='[["default",{"EVDR_TC_MODE":""}],["pair8",{"EVDR_TC_MODE":"pair8"}],["pair",{"EVDR_TC_MODE":"pair"}]]'
for nq in 16 32 64 256 1024; do echo "== nq=$nq"; timeout 200 python tools/tc_power.py 100000 $nq 2.5; done > gpurun_out/r2j_pair_dyn.log 2>&1
cat gpurun_out/r2j_pair_dyn.log
