cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
python -m pytest tests/test_gpu_dist.py tests/test_gpu_soak.py -x -q > gpurun_out/r2_pytest_dist6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_dist6.log
tail -3 gpurun_out/r2_pytest_dist6.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2d_dist2_n2.log 2>&1; echo "dist2 rc=$?" >> gpurun_out/r2d_dist2_n2.log
grep "dist_run\|rank 0 \|rc=" gpurun_out/r2d_dist2_n2.log | tail -5
