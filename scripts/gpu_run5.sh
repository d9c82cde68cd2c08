set -x
cd $GRAFT_REPO_ROOT
for p in 1 2; do
B200ANI_PIPELINES=$p timeout 600 python bench.py --workload anim --genomes 200 --ncu --steps 2 --warmup 2 > gpurun_out/r02_anim200_q$p.json 2> gpurun_out/r02_anim200_q$p.err; tail -2 gpurun_out/r02_anim200_q$p.err
python - <<EOF
import json
d=json.load(open('gpurun_out/r02_anim200_q$p.json'))
print('PIPELINES=$p', d['value'], d['ms_per_step'], d['wall_ms_per_step'], {k:round(v,1) for k,v in d['stage_ms_per_step'].items() if v})
EOF
done
