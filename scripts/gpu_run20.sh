cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=120
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"
tail -5 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_n8.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],l['e2e']['ms_per_step'],'scaling',l['scaling'])
print('parity',l['parity'])
print('roofline',l['roofline']['kernel'],l['roofline']['frac'])
print({k:round(v['ms_per_step'],2) for k,v in l['kernel_classes'].items()})
print('stages',{k:round(v,1) for k,v in l['stages_ms'].items()})
c1=l['configs1']; print('configs1',c1['value'],c1['ms_per_step'])
print('configs2',{k:v for k,v in l['configs2'].items() if k!='workload'})
c4=l.get('configs4'); 
if c4:
    print('configs4',{k:v for k,v in c4.items() if k not in ('workload','runs')})
    for r in c4.get('runs',[]): print({k:v for k,v in r.items() if k!='kernel_classes_rank0'}); print({k:(round(v['ms'],1),round(v['GBps'])) for k,v in r.get('kernel_classes_rank0',{}).items()})
PY
