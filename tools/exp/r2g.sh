mkdir -p gpurun_out
for nq in 1 2; do
python tools/tc_launch.py 100000 $nq 5 > gpurun_out/r2g_plain_$nq.log 2>&1 && ncu --set full --import-source on -k regex:maxsim_scan -s 2 -c 1 -o gpurun_out/r2g_scan_q${nq}_baseclk python tools/tc_launch.py 100000 $nq 5 > gpurun_out/r2g_ncu_$nq.log 2>&1
cat gpurun_out/r2g_plain_$nq.log
done
