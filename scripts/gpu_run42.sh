cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
timeout 150 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu12.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err; echo "bench2 rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_2gpu_b.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'parity',l.get('parity',{}).get('sha256_matches_oracle'))
print({k:round(v['ms_per_step'],2) for k,v in l.get('kernel_classes',{}).items()})
PY
