cd /root/repo
python scripts/profile_target_c4.py > gpurun_out/r2_prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_node_tables|k_lpnf_rank|k_lpnf_hard|k_rnear_apply|k_lcp_kasai|k_tree_level1" -c 7 -o /tmp/r2_prof_s3 python scripts/profile_target_c4.py > gpurun_out/r2_ncu_s3.log 2>&1
echo "full rc=$?"
ncu -i /tmp/r2_prof_s3.ncu-rep --page raw --csv > gpurun_out/r2_prof_s3_raw.csv 2>/dev/null
ncu -i /tmp/r2_prof_s3.ncu-rep --page source --csv -k regex:"k_node_tables" > gpurun_out/r2_prof_s3_src_node.csv 2>/dev/null
ncu -i /tmp/r2_prof_s3.ncu-rep --page source --csv -k regex:"k_lpnf_rank" > gpurun_out/r2_prof_s3_src_rank.csv 2>/dev/null
ls -la /tmp/r2_prof_s3.ncu-rep gpurun_out/r2_prof_s3_*.csv
cat gpurun_out/r2_prof_plain2.log
