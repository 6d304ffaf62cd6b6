set -x
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
python scripts/run_reference_tests.py > gpurun_out/r2_ref_suite.log 2>&1
tail -20 gpurun_out/r2_ref_suite.log
python scripts/c5_tail_demo.py > gpurun_out/r2_c5_tail.log 2>&1; echo "rc=$?" >> gpurun_out/r2_c5_tail.log
tail -3 gpurun_out/r2_c5_tail.log
nvidia-smi -L
