cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu8.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_gpu8.log
python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_new.log 2>&1; echo "ab new rc=$?"
NLZ_NODES_SCALAR=1 python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_scalar.log 2>&1; echo "ab scalar rc=$?"
NLZ_STAGE3_R1=1 python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_r1.log 2>&1; echo "ab r1 rc=$?"
paste -d'|' <(cut -d' ' -f1-3 gpurun_out/r2_s3ab_new.log) <(cut -d' ' -f3 gpurun_out/r2_s3ab_scalar.log) <(cut -d' ' -f3 gpurun_out/r2_s3ab_r1.log)
cut -d' ' -f1,4- gpurun_out/r2_s3ab_new.log | tr '\n' ';'; echo
python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_new.log 2>&1; tail -2 gpurun_out/r2_stage_new.log
NLZ_NODES_SCALAR=1 python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_scalar.log 2>&1; tail -2 gpurun_out/r2_stage_scalar.log
python scripts/stage_times.py c2 > gpurun_out/r2_stage_c2_new.log 2>&1; tail -2 gpurun_out/r2_stage_c2_new.log
NLZ_NODES_SCALAR=1 python scripts/stage_times.py c2 > gpurun_out/r2_stage_c2_scalar.log 2>&1; tail -1 gpurun_out/r2_stage_c2_scalar.log
python bench.py --no-c5 > gpurun_out/r2_bench_s3.json 2> gpurun_out/r2_bench_s3.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_s3.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity", {}).get("sha256_matches_oracle"))
PY
