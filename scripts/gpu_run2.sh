cd /root/repo
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_dist.py --deselect tests/test_gpu_soak.py --ignore tests/test_gpu_dist.py --ignore tests/test_gpu_soak.py > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu2.log
tail -8 gpurun_out/r2_pytest_gpu2.log
python bench.py --steps 10 --warmup 3 --no-single-text --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2_bench_a.json'))
print(l['value'], l['ms_per_step'], l['e2e']['value'])
print({k:round(v['ms_per_step'],3) for k,v in l['kernel_classes'].items()})
"
