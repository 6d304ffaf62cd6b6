cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=30
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu3.log
tail -6 gpurun_out/r2_pytest_gpu3.log
python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_c.json'))
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'])
print('parity',l['parity']['sha256_matches_oracle'])
print({k:round(v['ms_per_step'],2) for k,v in l['kernel_classes'].items()})
c1=l['configs1']; print('configs1',c1['value'],c1['ms_per_step'],{k:round(v['ms_per_step'],3) for k,v in c1['kernel_classes'].items()})
print('configs2',l['configs2']['value'],l['configs2']['parity'])
PY
