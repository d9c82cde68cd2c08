set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --workload anim50 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_anim50_b.json 2> gpurun_out/r02_anim50_b.err; tail -3 gpurun_out/r02_anim50_b.err; head -c 1500 gpurun_out/r02_anim50_b.json
