cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=30
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_pytest_dist_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist_2gpu.log
tail -5 gpurun_out/r2_pytest_dist_2gpu.log
timeout 300 python scripts/soak.py 60 7 > gpurun_out/r2_soak_2gpu.log 2>&1; echo "soak rc=$?" >> gpurun_out/r2_soak_2gpu.log
tail -3 gpurun_out/r2_soak_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2_dist2_n2.log 2>&1; echo "dist rc=$?" >> gpurun_out/r2_dist2_n2.log
grep -v "^W\|^\s*$" gpurun_out/r2_dist2_n2.log | tail -25
