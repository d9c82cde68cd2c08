set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_anim_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
for sh in 1 0; do
B200ANI_SHARE_INDEX=$sh timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$sh bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_n2_share$sh.json 2> gpurun_out/r02_n2_share$sh.err; tail -3 gpurun_out/r02_n2_share$sh.err
python - <<EOF
import json
for l in open('gpurun_out/r02_n2_share$sh.json'):
    if l.startswith('{'):
        d=json.loads(l); print('SHARE=$sh', d['value'], d['ms_per_step'], d['e2e']['value'])
        for r in d['per_rank']: print(r['rank'], round(r['ms_per_step'],1), {k:round(v,1) for k,v in r['stage_ms'].items() if k in ('index','index_share','total','seed','ext_dp')})
EOF
done
