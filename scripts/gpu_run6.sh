cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=120
nvidia-smi -L | wc -l; free -g | head -2; df -h /dev/shm | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/c5_run.py 3100000000 2 > gpurun_out/r2_c5_n8.log 2>&1; echo "c5 rc=$?" >> gpurun_out/r2_c5_n8.log
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2_c5_n8.log | tail -30
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2_dist2_n8.log 2>&1; echo "dist8 rc=$?" >> gpurun_out/r2_dist2_n8.log
grep "dist_run\|rank 0 \|rc=" gpurun_out/r2_dist2_n8.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2_dist2_n4.log 2>&1; echo "dist4 rc=$?" >> gpurun_out/r2_dist2_n4.log
grep "dist_run\|rank 0 \|rc=" gpurun_out/r2_dist2_n4.log | tail -12
