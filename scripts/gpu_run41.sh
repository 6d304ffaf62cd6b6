cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err; echo "bench2 rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_2gpu_b.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'parity',l.get('parity'))
print({k:round(v['ms_per_step'],2) for k,v in l.get('kernel_classes',{}).items()})
for k in ('configs1','configs2'):
    if k in l: print(k, l[k].get('value'), l[k].get('ms_per_step'), l[k].get('parity'))
PY
tail -3 gpurun_out/r2_bench_2gpu_b.err
