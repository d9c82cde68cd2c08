set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for p in 1 2 3; do
B200ANI_PIPELINES=$p timeout 600 python bench.py --workload anim --genomes 200 --ncu --steps 2 --warmup 2 > gpurun_out/r02_anim200_p$p.json 2> gpurun_out/r02_anim200_p$p.err; tail -2 gpurun_out/r02_anim200_p$p.err
python - <<EOF
import json
d=json.load(open('gpurun_out/r02_anim200_p$p.json'))
print('PIPELINES=$p', d['value'], d['ms_per_step'], d['wall_ms_per_step'], {k:round(v,1) for k,v in d['stage_ms_per_step'].items() if v})
EOF
done
B200ANI_PIPELINES=2 timeout 600 python bench.py --workload anib --genomes 50 --ncu --steps 1 --warmup 1 > gpurun_out/r02_anib50_b.json 2> gpurun_out/r02_anib50_b.err; tail -2 gpurun_out/r02_anib50_b.err; head -c 400 gpurun_out/r02_anib50_b.json
