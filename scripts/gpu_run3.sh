cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=20
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist.log
tail -30 gpurun_out/r2_pytest_dist.log
