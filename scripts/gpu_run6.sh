set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_anim500_n2.json 2> gpurun_out/r02_anim500_n2.err; tail -4 gpurun_out/r02_anim500_n2.err; head -c 1200 gpurun_out/r02_anim500_n2.json
