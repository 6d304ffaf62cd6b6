cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/c5_run.py 20000000 2 shuffled > gpurun_out/r2_c5_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2_c5_small.log
grep "\[c5\]\|rc=\|Error\|error" gpurun_out/r2_c5_small.log | cut -c1-400 | tail -12
