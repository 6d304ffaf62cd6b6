cd /root/repo
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_node_tables|k_lpnf_rank2|k_lpnf_hard" -c 4 -f -o /tmp/s3prof python scripts/profile_target_c4.py 60000000 > gpurun_out/r2_s3prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_s3prof.log
ncu -i /tmp/s3prof.ncu-rep --page raw --csv > gpurun_out/r2_s3prof_raw.csv 2>/dev/null; echo "raw rc=$?"
ncu -i /tmp/s3prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_s3prof_source.csv 2>gpurun_out/r2_s3prof_source.err; echo "source rc=$?"
ls -la gpurun_out/r2_s3prof_raw.csv gpurun_out/r2_s3prof_source.csv /tmp/s3prof.ncu-rep
