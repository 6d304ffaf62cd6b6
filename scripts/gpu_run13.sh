cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2b_dist2_n8.log 2>&1; echo "dist8 rc=$?" >> gpurun_out/r2b_dist2_n8.log
grep "dist_run\|rank 0 \|rc=" gpurun_out/r2b_dist2_n8.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2b_dist2_n4.log 2>&1; echo "dist4 rc=$?" >> gpurun_out/r2b_dist2_n4.log
grep "dist_run\|rank 0 \|rc=" gpurun_out/r2b_dist2_n4.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/dist_run.py 250000000 rc nocheck > gpurun_out/r2b_dist2_n2.log 2>&1; echo "dist2 rc=$?" >> gpurun_out/r2b_dist2_n2.log
grep "dist_run\|rank 0 \|rc=" gpurun_out/r2b_dist2_n2.log | tail -8
