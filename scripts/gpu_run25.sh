cd /root/repo
for wn in 4 8 16 24; do
NLZ_WALK_NODES=$wn python bench.py --steps 2 --warmup 1 --no-legs --no-cpu-baseline > gpurun_out/r2_walk_$wn.json 2>/dev/null
python - <<PY
import json
l=json.load(open('gpurun_out/r2_walk_$wn.json'))
k=l['kernel_classes']
print($wn, 'total',round(l['ms_per_step'],1),'rank',round(k['lpnf_rank']['ms_per_step'],1),'hard',round(k['lpnf_hard']['ms_per_step'],1),'hardpos',l['pipeline']['hard_positions'],'parity',l['parity']['sha256_matches_oracle'])
PY
done
