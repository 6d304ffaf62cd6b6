cd /root/repo
python scripts/profile_target_c4.py > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_lcp_kasai|k_group_stream|k_lpnf_rank|k_node_tables|k_rs_scatter|k_gather_rank|k_regroup_apply|k_lpnf_hard|k_rnear_apply|k_chain_exit|k_tile_sort" -c 80 -o gpurun_out/r2_prof_c4 python scripts/profile_target_c4.py > gpurun_out/r2_ncu_c4.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_c4.log; cat gpurun_out/r2_prof_plain.log
