cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench2 rc=$?"
tail -4 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_n2.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],l['e2e']['ms_per_step'],'scaling',l['scaling'])
print('parity',l['parity'])
print('roofline',l['roofline']['kernel'],l['roofline']['frac'])
print({k:round(v['ms_per_step'],2) for k,v in l['kernel_classes'].items()})
c1=l['configs1']; print('configs1',c1['value'],c1['ms_per_step'])
print('configs2',{k:v for k,v in l['configs2'].items() if k!='workload'})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/dist_run.py 5000000 rc nocheck 2>&1 | grep "rank 0 it\|kernel classes" | tail -3
