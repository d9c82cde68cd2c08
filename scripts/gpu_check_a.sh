set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --workload anib --steps 1 --warmup 3 > gpurun_out/r02_anib200_b.json 2> gpurun_out/r02_anib200_b.err; tail -3 gpurun_out/r02_anib200_b.err; head -c 1200 gpurun_out/r02_anib200_b.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err; tail -2 gpurun_out/r02_reference_n1.err; head -c 1500 gpurun_out/r02_reference_n1.json
