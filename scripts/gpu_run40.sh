cd /root/repo
export NLZ_BARRIER_TIMEOUT_S=60
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu11.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu11.log
tail -3 gpurun_out/r2_pytest_gpu11.log
timeout 200 python scripts/soak.py 60 12 > gpurun_out/r2_soak_c.log 2>&1; echo "soak rc=$?" >> gpurun_out/r2_soak_c.log; tail -2 gpurun_out/r2_soak_c.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final2_reference.json 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2_bench_final2.json'))
r=json.loads([x for x in open('gpurun_out/r2_bench_final2_reference.json') if x.startswith('{')][-1])
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],'ref',r['value'],'ratio e2e',l['e2e']['value']/r['value'])
print('roofline',l['roofline']['kernel'],l['roofline']['frac'],l['roofline']['traffic'])
print({k:round(v['ms_per_step'],2) for k,v in l['kernel_classes'].items()})
print('stages',{k:round(v,1) for k,v in l['stages_ms'].items()}, 'syncs', l['pipeline']['host_syncs'])
c1=l['configs1']; print('configs1',c1['value'],c1['ms_per_step'],c1['e2e']['value'],{k:round(v['ms_per_step'],3) for k,v in c1['kernel_classes'].items()})
print('configs2',l['configs2']['value'],l['configs2']['e2e'],l['configs2']['parity'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-legs --no-c5 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"k_build_keys|k_regroup_apply|k_scatter_pairs|k_lcp_kasai|k_tree_level1|k_node_tables|k_lpnf_rank|k_lpnf_hard" -c 14 -f -o /tmp/s3prof2 python scripts/profile_target_c4.py 60000000 > gpurun_out/r2b_s3prof.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/s3prof2.ncu-rep --page raw --csv > gpurun_out/r2b_s3prof_raw.csv 2>/dev/null; echo "raw rc=$?"
ls -la gpurun_out/r2b_*
