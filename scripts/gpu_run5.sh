cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=30
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -k "33_bit or small_cases" > gpurun_out/r2_pytest_dist33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist33.log
tail -15 gpurun_out/r2_pytest_dist33.log
