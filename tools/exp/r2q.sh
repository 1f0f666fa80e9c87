mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-corpus-1m"
$B > gpurun_out/r2q_bench_plain.json 2> gpurun_out/r2q_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv $B > gpurun_out/r2q_ncu_bench.log 2>&1
cap() {  # name mode pages nq
  EVDR_TC_MODE=$2 python tools/tc_launch.py $3 $4 3 > gpurun_out/r2q_plain_$1.log 2>&1 && EVDR_TC_MODE=$2 ncu --set full --clock-control none --import-source on -k regex:maxsim_ -s 1 -c 1 -o gpurun_out/r2q_prof_$1 python tools/tc_launch.py $3 $4 3 > gpurun_out/r2q_ncu_$1.log 2>&1
  cat gpurun_out/r2q_plain_$1.log
}
cap pair8_1024x100k "" 100000 1024
cap ew8_1024x100k ew8 100000 1024
cap scan_1x100k "" 100000 1
cap pair8_64x100k "" 100000 64
cap pair8_32x100k "" 100000 32
