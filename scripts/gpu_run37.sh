cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu10.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_gpu10.log
python scripts/dev/stage3_ab.py > gpurun_out/r2_s3ab_new.log 2>&1; echo "ab new rc=$?"
cut -d' ' -f1-3 gpurun_out/r2_s3ab_new.log | tr '\n' ';'; echo
cut -d' ' -f1,4- gpurun_out/r2_s3ab_new.log | tr '\n' ';'; echo
python scripts/stage_times.py 250000000 > gpurun_out/r2_stage_new.log 2>&1; tail -2 gpurun_out/r2_stage_new.log
python scripts/stage_times.py c2 > gpurun_out/r2_stage_c2_new.log 2>&1; tail -1 gpurun_out/r2_stage_c2_new.log
python scripts/stage_times.py c1 > gpurun_out/r2_stage_c1_new.log 2>&1; tail -1 gpurun_out/r2_stage_c1_new.log
