cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu6.log
tail -3 gpurun_out/r2_pytest_gpu6.log
timeout 300 python scripts/soak.py 90 11 > gpurun_out/r2_soak_b.log 2>&1; echo "soak rc=$?" >> gpurun_out/r2_soak_b.log; tail -2 gpurun_out/r2_soak_b.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_reference.json 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_final.json'))
r=json.loads([x for x in open('gpurun_out/r2_bench_final_reference.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'ref',r['value'],'ratio e2e',l['e2e']['value']/r['value'])
print('roofline',l['roofline']['kernel'],l['roofline']['frac'],l['roofline']['traffic'])
print({k:round(v['ms_per_step'],2) for k,v in l['kernel_classes'].items()})
print('stages',{k:round(v,1) for k,v in l['stages_ms'].items()}, 'syncs', l['pipeline']['host_syncs'])
c1=l['configs1']; print('configs1',c1['value'],c1['ms_per_step'],c1['e2e']['value'],{k:round(v['ms_per_step'],3) for k,v in c1['kernel_classes'].items()})
print('configs2',l['configs2']['value'],l['configs2']['e2e'],l['configs2']['parity'])
PY
