set -x
cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -3 gpurun_out/r02_bench_n$N.err; grep -a '^{' gpurun_out/r02_bench_n$N.json | head -c 600
if [ "$MIXED" = "1" ]; then
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload mixed --steps 1 --warmup 3 > gpurun_out/r02_mixed_n$N.json 2> gpurun_out/r02_mixed_n$N.err; tail -3 gpurun_out/r02_mixed_n$N.err; grep -a '^{' gpurun_out/r02_mixed_n$N.json | head -c 800
fi
