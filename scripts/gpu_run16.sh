cd /root/repo
# (1) launch list of the bench command (one metric: no replay)
python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-legs --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
echo "launch-list rc=$?"
# (2) full capture of the heaviest kernels on a 60 Mbp text of the same recipe (replays save/restore the whole workspace)
python scripts/profile_target_c4.py > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_gather_rank|k_lpnf_rank|k_rs_scatter|k_group_stream|k_node_tables|k_regroup_apply|k_lcp_kasai|k_tile_sort" -s 6 -c 14 -o /tmp/r2_prof_c4 python scripts/profile_target_c4.py > gpurun_out/r2_ncu_c4.log 2>&1
echo "full rc=$?"
ncu -i /tmp/r2_prof_c4.ncu-rep --page raw --csv > gpurun_out/r2_prof_c4_raw.csv 2>/dev/null
ls -la /tmp/r2_prof_c4.ncu-rep gpurun_out/r2_prof_c4_raw.csv gpurun_out/r2_launches.csv
tail -2 gpurun_out/r2_ncu_c4.log
