mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2ad_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ad_tests.log; tail -4 gpurun_out/r2ad_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2ad_bench_n1.json 2> gpurun_out/r2ad_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2ad_bench_ref.json 2>/dev/null; echo ref rc=$?
STRESS_SEED=31 timeout 200 python tools/stress_flavours.py 45 > gpurun_out/r2ad_stress.log 2>&1; echo stress rc=$?; tail -1 gpurun_out/r2ad_stress.log
