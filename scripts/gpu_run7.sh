set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload anim --genomes 200 --ncu --steps 2 --warmup 2 > gpurun_out/r02_anim200_r.json 2> gpurun_out/r02_anim200_r.err; tail -2 gpurun_out/r02_anim200_r.err
python - <<EOF
import json
d=json.load(open('gpurun_out/r02_anim200_r.json'))
print('anim200', d['value'], d['ms_per_step'], {k:round(v,1) for k,v in d['stage_ms_per_step'].items() if v})
EOF
timeout 900 python bench.py --workload anib --steps 1 --warmup 3 > gpurun_out/r02_anib200_a.json 2> gpurun_out/r02_anib200_a.err; tail -2 gpurun_out/r02_anib200_a.err; head -c 600 gpurun_out/r02_anib200_a.json
timeout 600 python bench.py --workload tetra --steps 5 --warmup 3 > gpurun_out/r02_tetra_b.json 2> gpurun_out/r02_tetra_b.err; tail -2 gpurun_out/r02_tetra_b.err; head -c 600 gpurun_out/r02_tetra_b.json
