mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2z_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_tests.log; tail -4 gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2z_bench_ref.json 2>/dev/null; echo ref rc=$?
cap() {  # name mode pages nq L
  EVDR_TC_MODE=$2 python tools/tc_launch.py $3 $4 3 $5 > gpurun_out/r2z_plain_$1.log 2>&1 && EVDR_TC_MODE=$2 ncu --set full --clock-control none --import-source on -k regex:maxsim_ -s 1 -c 1 -o gpurun_out/r2z_prof_$1 python tools/tc_launch.py $3 $4 3 $5 > gpurun_out/r2z_ncu_$1.log 2>&1
  cat gpurun_out/r2z_plain_$1.log
}
cap pair8_1024x100k "" 100000 1024 1040
cap scan_1x100k "" 100000 1 1040
cap sp4_1x1600k_64tok "" 1600000 1 64
cap pairsp_1024x400k_64tok "" 400000 1024 64
