cd /root/repo
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_final.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_reference.json 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_final.json'))
r=json.loads([x for x in open('gpurun_out/r2_bench_final_reference.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'ref',r['value'],'ratio e2e',l['e2e']['value']/r['value'])
print('roofline',l['roofline'])
print('clocks',l['clocks'])
print('cpu',{k:v for k,v in l['cpu_baseline'].items() if k in ('value','seconds','triples_identical_to_gpu')}, l['cpu_baseline']['parallel_mode']['value'])
print(l['config']==r['config'], r['cpu_baseline']['sample'][:80])
PY
python -c "import __graft_entry__ as g; g.smoke()"
