set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_anib_gpu.py -m gpu -q 2>&1 | tail -5
timeout 1200 python bench.py --workload anib --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_anib200_g.json 2> gpurun_out/r02_anib200_g.err; tail -3 gpurun_out/r02_anib200_g.err; head -c 300 gpurun_out/r02_anib200_g.json
