cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
python -m pytest tests -m gpu -x -q -s -k "balance or device_resident or shuffled_control" > gpurun_out/r2_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_new.log
grep -v "^$" gpurun_out/r2_pytest_new.log | tail -8
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu4.log
tail -4 gpurun_out/r2_pytest_gpu4.log
