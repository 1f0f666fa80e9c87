mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log; tail -8 gpurun_out/r2m_tests.log
python tools/compare_torch_gpu.py > gpurun_out/r2m_compare.log 2>&1; cat gpurun_out/r2m_compare.log
