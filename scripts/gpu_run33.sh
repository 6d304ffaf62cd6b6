cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu7.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_gpu7.log
python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_new.log 2>&1; echo "ab new rc=$?"
NLZ_STAGE3_R1=rank python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_rank.log 2>&1; echo "ab rank rc=$?"
NLZ_STAGE3_R1=1 python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_r1.log 2>&1; echo "ab r1 rc=$?"
paste -d'|' <(cut -d' ' -f1-3 gpurun_out/r2_s3ab_new.log) <(cut -d' ' -f3 gpurun_out/r2_s3ab_rank.log) <(cut -d' ' -f3 gpurun_out/r2_s3ab_r1.log)
cut -d' ' -f1,4- gpurun_out/r2_s3ab_new.log | tr '\n' ';'; echo
cut -d' ' -f1,4- gpurun_out/r2_s3ab_r1.log | tr '\n' ';'; echo
python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_new.log 2>&1; tail -2 gpurun_out/r2_stage_new.log
NLZ_STAGE3_R1=rank python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_rank.log 2>&1; tail -2 gpurun_out/r2_stage_rank.log
NLZ_STAGE3_R1=1 python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_r1.log 2>&1; tail -2 gpurun_out/r2_stage_r1.log
python scripts/stage_times.py c2 > gpurun_out/r2_stage_c2_new.log 2>&1; tail -2 gpurun_out/r2_stage_c2_new.log
python bench.py --no-c5 > gpurun_out/r2_bench_s3.json 2> gpurun_out/r2_bench_s3.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_s3.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity"))
print(d.get("kernels_ms") or {k: v for k, v in d.items() if "kernel" in k})
PY
