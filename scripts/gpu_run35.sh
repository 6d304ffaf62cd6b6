cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu9.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_gpu9.log
python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_new.log 2>&1; echo "ab new rc=$?"
NLZ_STAGE3_R1=1 python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_r1.log 2>&1; echo "ab r1 rc=$?"
paste -d'|' <(cut -d' ' -f1-3 gpurun_out/r2_s3ab_new.log) <(cut -d' ' -f3 gpurun_out/r2_s3ab_r1.log)
cut -d' ' -f1,4- gpurun_out/r2_s3ab_new.log | tr '\n' ';'; echo
python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_new.log 2>&1; tail -2 gpurun_out/r2_stage_new.log
NLZ_NODES_SCALAR=1 python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_scalar.log 2>&1; tail -1 gpurun_out/r2_stage_scalar.log
for b in 16 32 128; do NLZ_WALK_NODES=$b python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_b$b.log 2>&1; echo "budget $b"; tail -2 gpurun_out/r2_stage_b$b.log | cut -c1-400; done
python scripts/stage_times.py c2 > gpurun_out/r2_stage_c2_new.log 2>&1; tail -1 gpurun_out/r2_stage_c2_new.log
NLZ_NODES_SCALAR=1 python scripts/stage_times.py c2 > gpurun_out/r2_stage_c2_scalar.log 2>&1; tail -1 gpurun_out/r2_stage_c2_scalar.log
