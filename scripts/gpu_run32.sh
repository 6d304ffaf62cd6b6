cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=120
free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 scripts/c5_run.py 3100000000 2 shuffled > gpurun_out/r2_c5_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_c5_full.log
grep "\[c5\].*rank 0\|\[c5\] [a-z]* [a-z]*:\|generated\|rc=\|Error\|error\|threshold" gpurun_out/r2_c5_full.log | cut -c1-600 | tail -14
