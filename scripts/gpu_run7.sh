cd /root/repo
python bench.py --steps 5 --warmup 2 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_b.json'))
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e'])
print('parity',l['parity'])
print('roofline',l['roofline'])
print({k:round(v['ms_per_step'],2) for k,v in l['kernel_classes'].items()})
print('stages',l['stages_ms'])
c1=l['configs1']; print('configs1',c1['value'],c1['ms_per_step'],c1['e2e'],c1['python_list_api'],c1['roofline']['kernel'],c1['roofline']['frac'])
print('configs2',{k:v for k,v in l['configs2'].items() if k!='workload'})
print('cpu',l.get('cpu_baseline'))
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_b_reference.json 2>&1; echo "ref rc=$?"; cat gpurun_out/r2_bench_b_reference.json | cut -c1-600
